"""CPU restatement of the reference's prompt assembly — TEST INFRASTRUCTURE ONLY (see oracle/qwen3_tts_oracle.py).

What is restated and where it comes from
  * `_build_talker_inputs_local` ........ /root/reference/faster_qwen3_tts/model.py:331-553, statement by statement:
      instruct rows (:349-354), speaker row (:361-375), language / dialect id (:377-393), tts bos/eos/pad rows
      (:395-403), codec prefix ids (:405-417), pad/bos codec rows and optional speaker row (:418-428), role rows
      (:434-436), text-side pad..bos + codec[:-1] (:437-443), ICL branch (:447-460), first-text-token row (:462-471),
      non_streaming_mode rewrite (:472-504), trailing text (:505-514), left-pad by flip/pad/flip + mask (:519-535),
      trailing hiddens right-padded with the tts_pad vector (:537-551).
  * `generate_icl_prompt` ............... third-party `qwen-tts>=0.1.1` (pyproject.toml:29), NOT in /root/reference and not
      importable here.  Restated from its call site (model.py:452-459: arguments, return pair) and the layout recalled in
      SURVEY.md Appendix B.  PARITY UNPINNED for this function; everything else in this file follows in-tree reference
      lines.

The functions take an `OracleTTS` (embedding tables and text_projection on plain torch ops) and the config; ids are
CPU int64 tensors shaped like the reference's ([1, n]).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def generate_icl_prompt(orc, cfg, text_id, ref_id, ref_code, tts_pad_embed, tts_eos_embed, non_streaming_mode: bool):
    """Upstream helper called at model.py:452-459 (SURVEY.md Appendix B — recalled, unpinned).

    text  = [TP(ref_id ++ text_id), tts_eos]                      (T_t rows)
    codec = [CE(codec_bos), sum_g emb_g(ref_code[:, g])]          (T_ref + 1 rows)
    non-streaming: cat(text + CE(codec_pad), codec + tts_pad), trailing = tts_pad
    streaming:     position-wise overlay; the shorter side is padded with tts_pad, surplus text becomes the trailing rows
    """
    tc = cfg.talker
    text = torch.cat([orc.text_projection(torch.cat([ref_id, text_id], dim=1)), tts_eos_embed], dim=1)
    rows = orc.codec_embed(ref_code[:, 0])
    for g in range(orc.ncb):
        rows = rows + orc.pred_embed(g, ref_code[:, g + 1])
    bos = orc.codec_embed(torch.tensor([tc.codec_bos_id]))
    codec = torch.cat([bos, rows], dim=0).unsqueeze(0)
    pad_c = orc.codec_embed(torch.tensor([tc.codec_pad_id])).view(1, 1, -1)
    if non_streaming_mode:
        return torch.cat([text + pad_c, codec + tts_pad_embed], dim=1), tts_pad_embed
    Lt, Lc = text.shape[1], codec.shape[1]
    if Lt >= Lc:
        return text[:, :Lc] + codec, (text[:, Lc:] if Lt > Lc else tts_pad_embed)
    padded = torch.cat([text, tts_pad_embed.expand(-1, Lc - Lt, -1)], dim=1)
    return padded + codec, tts_pad_embed


def build_talker_inputs(orc, cfg, input_ids: Sequence[torch.Tensor], ref_ids: Sequence[Optional[torch.Tensor]],
                        voice_clone_prompt: Optional[dict], languages: Sequence[str], speakers: Optional[Sequence[Optional[str]]],
                        non_streaming_mode: bool, instruct_ids: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """model.py:331-553.  Returns (talker_input_embeds [B,T,H], attention_mask [B,T], trailing_text_hiddens [B,R,H],
    tts_pad_embed [1,1,H])."""
    tc = cfg.talker
    TP = orc.text_projection
    CE = lambda ids: orc.codec_embed(torch.tensor(ids, dtype=torch.long))
    per_item: List[List[torch.Tensor]] = [[] for _ in input_ids]
    spk_embeds = None
    if voice_clone_prompt is not None:  # :346-347 (m.generate_speaker_prompt: one row per item)
        spk_embeds = [e.reshape(-1).to(torch.bfloat16).cpu() for e in voice_clone_prompt["ref_spk_embedding"]]
    if instruct_ids is not None:  # :349-354
        for i, ins in enumerate(instruct_ids):
            if ins is not None:
                per_item[i].append(TP(ins))
    if speakers is None:
        speakers = [None] * len(input_ids)
    trailing = []
    pad_e = None
    for i, (ids, language, speaker) in enumerate(zip(input_ids, languages, speakers)):
        if spk_embeds is None:  # :361-372
            if speaker == "" or speaker is None:
                spk = None
            else:
                if speaker.lower() not in tc.spk_id:
                    raise NotImplementedError(f"Speaker {speaker} not implemented")
                spk = CE(tc.spk_id[speaker.lower()])
        else:  # :373-377
            spk = spk_embeds[i] if (voice_clone_prompt["x_vector_only_mode"][i] or voice_clone_prompt["icl_mode"][i]) else None
        assert language is not None
        if language.lower() == "auto":  # :379-385
            lang_id = None
        else:
            if language.lower() not in tc.codec_language_id:
                raise NotImplementedError(f"Language {language} not implemented")
            lang_id = tc.codec_language_id[language.lower()]
        if language.lower() in ["chinese", "auto"] and speaker not in ("", None) and tc.spk_is_dialect[speaker.lower()]:  # :387-393
            lang_id = tc.codec_language_id[tc.spk_is_dialect[speaker.lower()]]
        bos_e, eos_e, pad_e = TP(torch.tensor([[cfg.tts_bos_token_id, cfg.tts_eos_token_id, cfg.tts_pad_token_id]])).chunk(3, dim=1)
        if lang_id is None:  # :405-417
            prefix = [[tc.codec_nothink_id, tc.codec_think_bos_id, tc.codec_think_eos_id]]
        else:
            prefix = [[tc.codec_think_id, tc.codec_think_bos_id, lang_id, tc.codec_think_eos_id]]
        c0 = CE(prefix)
        c1 = CE([[tc.codec_pad_id, tc.codec_bos_id]])
        codec = torch.cat([c0, c1], dim=1) if spk is None else torch.cat([c0, spk.view(1, 1, -1), c1], dim=1)  # :425-428
        role = TP(ids[:, :3])  # :434-436
        body = torch.cat((pad_e.expand(-1, codec.shape[1] - 2, -1), bos_e), dim=1) + codec[:, :-1]  # :437-443
        x = torch.cat((role, body), dim=1)
        icl = (voice_clone_prompt is not None and voice_clone_prompt.get("ref_code", None) is not None
               and voice_clone_prompt["icl_mode"][i])
        if icl:  # :447-460
            icl_embed, trail = generate_icl_prompt(orc, cfg, ids[:, 3:-5], ref_ids[i][:, 3:-2],
                                                   voice_clone_prompt["ref_code"][i].cpu().clone(), pad_e, eos_e, non_streaming_mode)
            x = torch.cat([x, icl_embed], dim=1)
        else:
            x = torch.cat([x, TP(ids[:, 3:4]) + codec[:, -1:]], dim=1)  # :462-471
            if non_streaming_mode:  # :472-504
                x = x[:, :-1]
                n_text = ids[:, 3:-5].shape[1]
                x = torch.cat([
                    x,
                    torch.cat((TP(ids[:, 3:-5]), eos_e), dim=1) + CE([[tc.codec_pad_id] * (n_text + 1)]),
                    pad_e + CE([[tc.codec_bos_id]]),
                ], dim=1)
                trail = pad_e
            else:  # :505-514
                trail = torch.cat((TP(ids[:, 4:-5]), eos_e), dim=1)
        per_item[i].append(x)
        trailing.append(trail)
    seqs = [torch.cat([t for t in parts if t is not None], dim=1).squeeze(0) for parts in per_item]  # :516-517
    # :519-535 — left padding via flip / pad_sequence / flip, mask = idx >= num_pads
    lens = torch.tensor([s.shape[0] for s in seqs])
    rev = torch.nn.utils.rnn.pad_sequence([s.flip(dims=[0]) for s in seqs], batch_first=True, padding_value=0.0)
    tie = rev.flip(dims=[1])
    B, T = tie.shape[0], tie.shape[1]
    tam = (torch.arange(T).expand(B, -1) >= (T - lens).unsqueeze(1)).long()
    # :537-551 — trailing hiddens right-padded with the tts_pad vector
    tr = [t.squeeze(0) for t in trailing]
    tl = [t.shape[0] for t in tr]
    tth = torch.nn.utils.rnn.pad_sequence(tr, batch_first=True, padding_value=0.0)
    mask = torch.arange(max(tl)).expand(len(tl), -1) >= torch.tensor(tl).unsqueeze(1)
    tth[mask] = pad_e.squeeze()
    return tie, tam, tth, pad_e
