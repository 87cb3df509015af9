"""Stand-in for the un-vendored `qwen_tts.Qwen3TTSModel` that the reference wraps (`model.py:102-112`).

The reference dereferences a fixed set of attributes and helpers on that object (listed in SURVEY.md §8c:
`.model.talker`, `.model.config.talker_config`, `.model.speech_tokenizer`, `_build_assistant_text`,
`_tokenize_texts`, `create_voice_clone_prompt`, `generate_speaker_prompt`, `generate_icl_prompt`, ...).  This module
provides exactly that surface on top of the fq3 engine, so `FasterQwen3TTS` can keep the reference's call
structure.  Everything dimension- or id-dependent comes from `TTSConfig`.

Two ways in:
* a checkpoint directory (`from_pretrained("/path/to/Qwen3-TTS-12Hz-0.6B-Base")`): `checkpoint.py` reads config + safetensors
  into the arena and the codec decoder, `HFTokenizer` wraps the directory's BPE tokenizer files, `frontend.py` turns a
  reference wav into an x-vector (ECAPA speaker encoder) and reference codes (Mimi-style codec encoder);
* `synthetic://<preset>`: seeded random-init weights (BASELINE.json: no checkpoints offline), a hashing tokenizer, and
  `create_voice_clone_prompt` derives a deterministic pseudo x-vector and pseudo ref codes from the audio bytes so that every
  code path downstream (prompt layout, ICL prefill length, ref-code prepending and proportional trimming) runs with the
  right shapes.
"""
from __future__ import annotations

import hashlib
import logging
import os
import types
import wave
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .config import TTSConfig, preset
from .engine import Engine
from .weights import Arena, init_synthetic, pack_arena

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------------
# text side
# ------------------------------------------------------------------------------------------------
class SyntheticTokenizer:
    """Deterministic word/byte hashing into the text vocabulary (no BPE files offline).

    Templates match the slicing the reference applies to the ids (`model.py:435,454,466,480,509`):
    assistant text = 3 role ids + text + 5 trailer ids; ref text = 3 role ids + text + 2 trailer ids."""

    def __init__(self, cfg: TTSConfig):
        self.cfg = cfg
        self.lo, self.hi = 1000, max(2000, min(cfg.talker.text_vocab_size, cfg.tts_pad_token_id) - 8)
        if cfg.talker.text_vocab_size < 4096:  # tiny preset
            self.lo, self.hi = 16, cfg.talker.text_vocab_size - 16

    def encode(self, text: str) -> List[int]:
        ids = []
        for word in text.replace("\n", " \n ").split(" "):
            if not word:
                continue
            h = int.from_bytes(hashlib.blake2s(word.encode("utf-8"), digest_size=4).digest(), "little")
            ids.append(self.lo + h % (self.hi - self.lo))
        return ids or [self.lo]

    def role(self, who: str) -> List[int]:
        c = self.cfg
        return [c.im_start_token_id, c.assistant_token_id if who == "assistant" else c.user_token_id, c.newline_token_id]

    def assistant(self, text: str) -> List[int]:
        c = self.cfg
        return self.role("assistant") + self.encode(text) + [c.im_end_token_id, c.newline_token_id] + self.role("assistant")

    def ref(self, text: str) -> List[int]:
        c = self.cfg
        return self.role("assistant") + self.encode(text) + [c.im_end_token_id, c.newline_token_id]

    def instruct(self, text: str) -> List[int]:
        c = self.cfg
        return self.role("user") + self.encode(text) + [c.im_end_token_id, c.newline_token_id]


class HFTokenizer:
    """The checkpoint's own BPE tokenizer (`tokenizer.json`, or `vocab.json` + `merges.txt` through transformers).  The chat
    templates are the strings whose token layout the reference's slicing assumes (`model.py:435,454,466,480,509`; SURVEY.md
    §8c): the ids come out as 3 role ids + text + 5 (assistant) / 2 (ref, instruct) trailer ids."""

    ASSISTANT = "<|im_start|>assistant\n{}<|im_end|>\n<|im_start|>assistant\n"
    REF = "<|im_start|>assistant\n{}<|im_end|>\n"
    INSTRUCT = "<|im_start|>user\n{}<|im_end|>\n"

    def __init__(self, path: str):
        tj = os.path.join(path, "tokenizer.json")
        if os.path.exists(tj):
            from tokenizers import Tokenizer

            self._tok = Tokenizer.from_file(tj)
            self._encode = lambda t: self._tok.encode(t, add_special_tokens=False).ids
        elif os.path.exists(os.path.join(path, "vocab.json")):
            from transformers import AutoTokenizer

            self._tok = AutoTokenizer.from_pretrained(path)
            self._encode = lambda t: self._tok(t, add_special_tokens=False)["input_ids"]
        else:
            raise FileNotFoundError(f"{path}: neither tokenizer.json nor vocab.json + merges.txt")

    def encode(self, text: str) -> List[int]:
        return list(self._encode(text))

    def assistant(self, text: str) -> List[int]:
        return self.encode(self.ASSISTANT.format(text))

    def ref(self, text: str) -> List[int]:
        return self.encode(self.REF.format(text))

    def instruct(self, text: str) -> List[int]:
        return self.encode(self.INSTRUCT.format(text))


@dataclass
class VoiceClonePromptItem:
    ref_code: Optional[torch.Tensor]
    ref_spk_embedding: torch.Tensor
    x_vector_only_mode: bool
    icl_mode: bool
    ref_text: Optional[str]


# ------------------------------------------------------------------------------------------------
# talker-side modules the reference reaches through `m.talker`
# ------------------------------------------------------------------------------------------------
class _Embedding:
    """nn.Embedding look-alike over an arena view (row gather is pure data movement)."""

    def __init__(self, weight: torch.Tensor):
        self.weight = weight

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.embedding(ids.to(self.weight.device), self.weight)


class _CodecHead:
    """`talker.codec_head` (generate.py:182): GEMV through the streaming kernel."""

    def __init__(self, engine: Engine, weight: torch.Tensor):
        self.engine, self.weight = engine, weight

    def __call__(self, hidden: torch.Tensor) -> torch.Tensor:
        x = hidden.reshape(-1, hidden.shape[-1]).to(torch.bfloat16).contiguous()
        return self.engine.linear(self.weight, x).reshape(*hidden.shape[:-1], -1)


class _TextProjection:
    """`talker.text_projection`: Linear(+bias) -> SiLU -> Linear(+bias).  All rows of a call in two tensor-core GEMMs of
    libfq3codec.so reading the arena weights in place (a prompt has 3-60 text rows: this is a small dense contraction, and
    twelve 8-row launches of the streaming kernel were half of the host time of the prompt build); the streaming kernel's
    `fq3_linear` stays the path for shapes the GEMM does not take (K < 64) and when FQ3_TEXTPROJ_GEMM=0."""

    def __init__(self, engine: Engine, arena: Arena):
        self.engine = engine
        self.w1 = arena.view("talker.text_projection.linear_fc1.weight")
        self.b1 = arena.view("talker.text_projection.linear_fc1.bias")
        self.w2 = arena.view("talker.text_projection.linear_fc2.weight")
        self.b2 = arena.view("talker.text_projection.linear_fc2.bias")
        self._gemm = None
        self._plans = {}
        if os.environ.get("FQ3_TEXTPROJ_GEMM", "1") != "0" and self.w1.shape[1] % 64 == 0 and self.w2.shape[1] % 64 == 0:
            from . import codec as _codec
            self._gemm = _codec
            self._lib = _codec.load_lib()
            self._b1f, self._b2f = self.b1.float().contiguous(), self.b2.float().contiguous()  # the GEMM epilogue adds an fp32 bias

    def _plan(self, M: int):
        c = self._gemm
        dev = self.w1.device
        with torch.inference_mode(False):  # cached staging buffers are written in place from either mode
            x = torch.empty(M, self.w1.shape[1], dtype=torch.bfloat16, device=dev)
            h = torch.empty(M, self.w1.shape[0], dtype=torch.bfloat16, device=dev)
            y = torch.empty(M, self.w2.shape[0], dtype=torch.bfloat16, device=dev)
        ops = []
        for A, W, b, out, fl in ((x, self.w1, self._b1f, h, c.F_BIAS | c.F_SILU), (h, self.w2, self._b2f, y, c.F_BIAS)):
            o = c.Op()
            o.kind, o.flags = c.K_GEMM, fl
            o.M, o.N, o.K, o.taps, o.cin = M, W.shape[0], W.shape[1], 1, W.shape[1]
            o.a_rows, o.lda, o.col_mod, o.ldc = M, A.shape[1], W.shape[0], out.shape[1]
            o.A, o.B, o.C, o.bias = A.data_ptr(), W.data_ptr(), out.data_ptr(), b.data_ptr()
            ops.append(o)
        keep = [x, h, y]
        with torch.inference_mode(False):
            c.attach_splitk_workspace(ops, dev, keep)
        return (c.Op * 2)(*ops), x, y, keep

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        rows = x.reshape(-1, shape[-1]).to(torch.bfloat16).contiguous()
        M = rows.shape[0]
        if M == 0:  # a one-word text leaves no trailing rows (model.py:505-514 projects an empty slice)
            return torch.empty(*shape[:-1], self.w2.shape[0], dtype=torch.bfloat16, device=rows.device)
        if self._gemm is not None and 0 < M <= 512:
            plan = self._plans.get(M)
            if plan is None:
                if len(self._plans) >= 32:
                    self._plans.clear()
                plan = self._plans[M] = self._plan(M)
            arr, xin, yout, _ = plan
            xin.copy_(rows)
            if self._lib.fq3c_run(arr, 2, torch.cuda.current_stream().cuda_stream) != 0:
                raise self._gemm.CodecError(self._lib.fq3c_last_error().decode())
            return yout.clone().reshape(*shape[:-1], -1)
        outs = []
        for i in range(0, M, 8):
            h = self.engine.linear(self.w1, rows[i:i + 8].contiguous(), bias=self.b1, silu=True)
            outs.append(self.engine.linear(self.w2, h, bias=self.b2))
        return torch.cat(outs, 0).reshape(*shape[:-1], -1)


class _CodePredictor:
    def __init__(self, embeds: List[_Embedding]):
        self._embeds = embeds

    def get_input_embeddings(self):
        return self._embeds


class Talker:
    """Surface of `Qwen3TTSTalkerForConditionalGeneration` used by generate.py:99-124 and model.py:353-514."""

    def __init__(self, engine: Engine, arena: Arena, cfg: TTSConfig, stream_idx: int = 0):
        if stream_idx != 0:
            raise ValueError("the Talker operator surface addresses stream 0 only (the reference's seam is bs = 1)")
        self.engine, self.cfg, self.stream_idx = engine, cfg, stream_idx
        self.device = engine.device
        self.config = cfg.talker
        self.rope_deltas = None
        self._codec_embed = _Embedding(arena.view("talker.codec_embedding"))
        self._text_embed = _Embedding(arena.view("talker.model.text_embedding.weight"))
        self.text_projection = _TextProjection(engine, arena)
        self.codec_head = _CodecHead(engine, arena.view("talker.codec_head"))
        self.code_predictor = _CodePredictor(
            [_Embedding(arena.view(f"talker.code_predictor.codec_embedding.{i}")) for i in range(cfg.predictor.num_codebooks)]
        )

    def get_input_embeddings(self):
        return self._codec_embed

    def get_text_embeddings(self):
        return self._text_embed

    def forward(self, inputs_embeds, attention_mask=None, trailing_text_hidden=None, tts_pad_embed=None, **_):
        """Prefill as the reference's operator-at-a-time loop calls it (generate.py:107-118).  The engine writes
        K/V straight into the static cache, so `past_key_values` is a token the TalkerGraph recognises."""
        from .engine import SamplingPolicy

        n_pad = 0 if attention_mask is None else int((attention_mask[0] == 0).sum().item())
        if trailing_text_hidden is not None and tts_pad_embed is not None:
            self.engine.set_text_conditioning(self.stream_idx, trailing_text_hidden[0], tts_pad_embed)
        logits = self.engine.prefill(self.stream_idx, inputs_embeds[0], n_pad, SamplingPolicy(do_sample=False), want_logits=True)
        hidden = self.engine.last_hidden(0)  # fq3_prefill leaves the last row's hidden state in row 0 (include/fq3.h)
        self.rope_deltas = torch.tensor([[-n_pad]], dtype=torch.float32, device=self.device)
        T = inputs_embeds.shape[1]
        return types.SimpleNamespace(
            logits=logits.to(torch.bfloat16).view(1, 1, -1), past_hidden=hidden.view(1, 1, -1), generation_step=0,
            past_key_values=EnginePrefilledKV(self.engine, self.stream_idx, T),
        )


class EnginePrefilledKV:
    """Marker returned by Talker.forward: the prefix KV already lives in the engine's static cache."""

    def __init__(self, engine: Engine, stream_idx: int, length: int):
        self.engine, self.stream_idx, self.length = engine, stream_idx, length


# ------------------------------------------------------------------------------------------------
# the base model object
# ------------------------------------------------------------------------------------------------
class InnerModel:
    """`base_model.model` (Qwen3TTSForConditionalGeneration surface)."""

    def __init__(self, cfg: TTSConfig, talker: Talker, speech_tokenizer):
        self.talker = talker
        self.speech_tokenizer = speech_tokenizer
        self.tts_model_type = cfg.tts_model_type
        self.tts_model_size = cfg.tts_model_size
        self.config = types.SimpleNamespace(
            talker_config=cfg.talker, tts_bos_token_id=cfg.tts_bos_token_id, tts_eos_token_id=cfg.tts_eos_token_id,
            tts_pad_token_id=cfg.tts_pad_token_id,
        )
        self._cfg = cfg

    def generate_speaker_prompt(self, voice_clone_prompt) -> List[torch.Tensor]:
        """model.py:347 — one speaker row [H] per item."""
        return [e.to(self.talker.device, torch.bfloat16).reshape(-1) for e in voice_clone_prompt["ref_spk_embedding"]]

    def generate_icl_prompt(self, text_id, ref_id, ref_code, tts_pad_embed, tts_eos_embed, non_streaming_mode):
        """Upstream ICL prompt as recalled in SURVEY.md Appendix B: text = [TP(ref ++ text), eos];
        codec = [CE(bos), sum_g emb_g(ref_code[:, g])]; non-streaming concatenates the two blocks, streaming
        overlays them position-wise and returns the surplus text as the trailing hiddens."""
        t = self.talker
        tc = self._cfg.talker
        dev = t.device
        text = torch.cat([t.text_projection(t.get_text_embeddings()(torch.cat([ref_id, text_id], dim=1))), tts_eos_embed], dim=1)
        ref_code = ref_code.to(dev)
        rows = t.get_input_embeddings()(ref_code[:, 0])
        for g, emb in enumerate(t.code_predictor.get_input_embeddings()):
            rows = rows + emb(ref_code[:, g + 1])
        bos = t.get_input_embeddings()(torch.tensor([tc.codec_bos_id], device=dev))
        codec = torch.cat([bos, rows], dim=0).unsqueeze(0)
        pad_c = t.get_input_embeddings()(torch.tensor([tc.codec_pad_id], device=dev)).view(1, 1, -1)
        if non_streaming_mode:
            out = torch.cat([text + pad_c, codec + tts_pad_embed], dim=1)
            return out, tts_pad_embed
        Lt, Lc = text.shape[1], codec.shape[1]
        if Lt >= Lc:
            out = text[:, :Lc] + codec
            trailing = text[:, Lc:] if Lt > Lc else tts_pad_embed
        else:
            padded = torch.cat([text, tts_pad_embed.expand(-1, Lc - Lt, -1)], dim=1)
            out = padded + codec
            trailing = tts_pad_embed
        return out, trailing


class Qwen3TTSBaseModel:
    """`base_model` of the reference (qwen_tts.Qwen3TTSModel surface)."""

    def __init__(self, cfg: TTSConfig, engine: Engine, arena: Arena, speech_tokenizer, tokenizer=None, frontend=None):
        self.cfg = cfg
        self.engine = engine
        self.arena = arena
        self.tokenizer = tokenizer if tokenizer is not None else SyntheticTokenizer(cfg)
        self.frontend = frontend  # frontend.VoiceFrontEnd of a real checkpoint; None = pseudo voice encoders (synthetic presets)
        self.device = engine.device
        self.model = InnerModel(cfg, Talker(engine, arena, cfg), speech_tokenizer)

    @property
    def speech_tokenizer(self):
        """The wrapper level also hands out the codec (`model.model.speech_tokenizer.decode(...)` in
        examples/generate_with_embedding.py:98); the reference's own code reaches it one level down (model.py:56-58, :642)."""
        return self.model.speech_tokenizer

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_name: str, device_map="cuda", torch_dtype=torch.bfloat16, attn_implementation="sdpa",
                        max_seq_len: int = 2048, max_streams: int = 1, seed: int = 0, weights=None, cfg: Optional[TTSConfig] = None,
                        allow_synthetic: bool = False):
        """A checkpoint DIRECTORY (what a user of the reference passes, model.py:107) is loaded for real: config.json +
        safetensors -> arena / codec decoder (`checkpoint.py`), the directory's BPE tokenizer, and — when the files carry them —
        the speaker encoder and the codec encoder for reference audio (`frontend.py`).  A hub id ("Qwen/Qwen3-TTS-12Hz-0.6B-Base")
        is resolved through the local Hugging Face cache only (no network here) and raises when it is not cached.

        "synthetic://0.6B-Base", "synthetic://1.7B-CustomVoice", or the bare preset names ("0.6B-Base", "tiny", ...) build
        seeded RANDOM-INIT weights of that architecture (BASELINE.json: no checkpoints offline), a hashing tokenizer and pseudo
        speaker / reference-code encoders; `allow_synthetic=True` asks for that stand-in under a real name instead of an
        error.  The audio is noise then: only shapes and timing are meaningful."""
        from .codec import SpeechTokenizer

        dev = torch.device(device_map if isinstance(device_map, str) else "cuda")
        if dev.type != "cuda":
            raise ValueError("the fq3 engine needs a CUDA device")
        if torch_dtype not in (torch.bfloat16, "bfloat16"):
            raise ValueError("the fq3 engine computes in bf16 (fp32 accumulate); other dtypes are not implemented")
        looks_real = os.path.isdir(model_name) or model_name.lower().startswith("qwen/") or model_name.lower().endswith(".safetensors")
        if looks_real and not allow_synthetic and weights is None:
            path = model_name if os.path.isdir(model_name) else cls._resolve_hub_id(model_name)
            return cls._from_checkpoint(path, dev, max_seq_len, max_streams)
        if weights is None:
            logger.warning("fq3: %r -> seeded random-init weights, hashing tokenizer, pseudo voice encoders (synthetic stand-in)", model_name)
        cfg = cfg or preset(os.path.basename(os.path.normpath(model_name)) if os.path.isdir(model_name) else model_name)
        w = weights if weights is not None else init_synthetic(cfg, seed=seed)
        arena = pack_arena(cfg, w, max_seq_len, dev)
        engine = Engine(cfg, arena, max_seq_len=max_seq_len, max_streams=max_streams, max_frames=max(4096, max_seq_len))
        tok = SpeechTokenizer.synthetic(cfg.codec, dev, seed=seed + 1)
        return cls(cfg, engine, arena, tok)

    @staticmethod
    def _resolve_hub_id(model_name: str) -> str:
        try:
            from huggingface_hub import snapshot_download

            return snapshot_download(model_name, local_files_only=True)
        except Exception as e:  # not cached / hub library absent: there is no network to fall back to
            raise FileNotFoundError(
                f"{model_name!r} is not in the local Hugging Face cache and this build never downloads: pass the checkpoint "
                "directory, or 'synthetic://<preset>' for random-init weights of the same architecture"
            ) from e

    @classmethod
    def _from_checkpoint(cls, path: str, dev, max_seq_len: int, max_streams: int):
        """model.py:107-119 for a directory in the layout of checkpoint.py's docstring."""
        import json

        from . import checkpoint as ck
        from . import frontend as fe
        from .codec import CodecDecoder, SpeechTokenizer

        cfg = ck.read_config(path)
        w = ck.load_lm_weights(path, cfg)
        arena = pack_arena(cfg, w, max_seq_len, dev)
        del w
        engine = Engine(cfg, arena, max_seq_len=max_seq_len, max_streams=max_streams, max_frames=max(4096, max_seq_len))
        sdir = os.path.join(path, "speech_tokenizer")
        if not os.path.isdir(sdir):
            raise ck.CheckpointError(f"{path}: no speech_tokenizer/ directory (the 12 Hz codec decoder lives there)")
        tok = SpeechTokenizer(CodecDecoder(cfg.codec, ck.load_codec_weights(sdir, cfg.codec), dev))
        with open(os.path.join(path, "config.json")) as f:
            raw = json.load(f)
        frontend = None
        spk_w = ck.load_prefixed(path, ("speaker_encoder.", "model.speaker_encoder."))
        if spk_w:
            scfg = dict(raw.get("speaker_encoder_config") or {})
            scfg.setdefault("enc_dim", cfg.speaker_embed_dim or cfg.talker.hidden_size)
            codec_enc = None
            enc_w = ck.load_prefixed(sdir, ("encoder.", "model.encoder."))
            if enc_w:
                with open(os.path.join(sdir, "config.json")) as f:
                    craw = json.load(f)
                codec_enc = fe.CodecEncoder(craw.get("encoder_config") or {}, enc_w, cfg.talker.num_code_groups, dev)
            frontend = fe.VoiceFrontEnd(fe.SpeakerEncoder(scfg, spk_w, dev), codec_enc)
        elif cfg.tts_model_type == "base":
            logger.warning("fq3: %s carries no speaker_encoder.* tensors; voice cloning falls back to pseudo voice encoders", path)
        return cls(cfg, engine, arena, tok, tokenizer=HFTokenizer(path), frontend=frontend)

    # ---- text helpers (model.py:223-261) ---------------------------------------------------------
    def _build_assistant_text(self, text: str) -> str:
        return text

    def _build_ref_text(self, text: str) -> str:
        return "\x00ref\x00" + text

    def _build_instruct_text(self, text: str) -> str:
        return "\x00instruct\x00" + text

    def _tokenize_texts(self, texts: Sequence[str]) -> List[torch.Tensor]:
        out = []
        for t in texts:
            if t.startswith("\x00ref\x00"):
                ids = self.tokenizer.ref(t[5:])
            elif t.startswith("\x00instruct\x00"):
                ids = self.tokenizer.instruct(t[10:])
            else:
                ids = self.tokenizer.assistant(t)
            t_ = torch.tensor([ids], dtype=torch.long, device=self.device)
            t_._host_ids = list(ids)  # the prompt builder lays the prompt out on the host: spare it a device read-back
            out.append(t_)
        return out

    # ---- validation (model.py:846-847) ---------------------------------------------------------
    def get_supported_speakers(self) -> List[str]:
        return sorted(self.cfg.talker.spk_id)

    def get_supported_languages(self) -> List[str]:
        return ["auto"] + sorted(self.cfg.talker.codec_language_id)

    def _validate_languages(self, languages):
        for l in languages:
            if l is not None and l.lower() != "auto" and l.lower() not in self.cfg.talker.codec_language_id:
                raise ValueError(f"Unsupported language {l!r}; supported: {self.get_supported_languages()}")

    def _validate_speakers(self, speakers):
        for s in speakers:
            if s is None or s.lower() not in self.cfg.talker.spk_id:
                raise ValueError(f"Unsupported speaker {s!r}; supported: {self.get_supported_speakers()}")

    # ---- voice prompts (model.py:230-265) --------------------------------------------------------
    def _audio_fingerprint(self, ref_audio) -> Tuple[int, float]:
        """(seed, seconds) of a reference clip: path -> bytes hash + duration; (array, sr) -> content hash."""
        if isinstance(ref_audio, (tuple, list)):
            a, sr = np.asarray(ref_audio[0], dtype=np.float32), int(ref_audio[1])
            h = hashlib.blake2s(a.tobytes(), digest_size=8).digest()
            return int.from_bytes(h, "little"), len(a) / max(sr, 1)
        path = str(ref_audio)
        secs = 3.0
        try:
            with wave.open(path, "rb") as wf:
                secs = wf.getnframes() / float(wf.getframerate())
            with open(path, "rb") as f:
                h = hashlib.blake2s(f.read(1 << 20), digest_size=8).digest()
        except (FileNotFoundError, wave.Error, EOFError):
            h = hashlib.blake2s(path.encode(), digest_size=8).digest()
        return int.from_bytes(h, "little"), secs

    def create_voice_clone_prompt(self, ref_audio, ref_text: str = "", x_vector_only_mode: bool = False):
        if self.frontend is not None:  # real checkpoint: speaker encoder + codec encoder (frontend.py)
            spk = self.frontend.x_vector(ref_audio).to(torch.bfloat16).to(self.device)
            code = None if x_vector_only_mode else self.frontend.ref_codes(ref_audio).to(self.device)
            return [VoiceClonePromptItem(code, spk, x_vector_only_mode, not x_vector_only_mode, ref_text or None)]
        seed, secs = self._audio_fingerprint(ref_audio)
        g = torch.Generator().manual_seed(seed % (2**63 - 1))
        H = self.cfg.talker.hidden_size
        spk = (0.05 * torch.randn(H, generator=g)).to(torch.bfloat16).to(self.device)
        code = None
        if not x_vector_only_mode:
            n = max(1, int(round(secs * 12.5)))
            code = torch.randint(0, self.cfg.codec.codebook_size, (n, self.cfg.talker.num_code_groups), generator=g).to(self.device)
        return [VoiceClonePromptItem(code, spk, x_vector_only_mode, not x_vector_only_mode, ref_text or None)]

    def _prompt_items_to_voice_clone_prompt(self, items):
        return dict(
            ref_code=[i.ref_code for i in items], ref_spk_embedding=[i.ref_spk_embedding for i in items],
            x_vector_only_mode=[i.x_vector_only_mode for i in items], icl_mode=[i.icl_mode for i in items],
        )
