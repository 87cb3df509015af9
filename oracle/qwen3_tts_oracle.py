"""CPU/PyTorch restatement of the reference hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module; the product package never does (it fails loudly when libfq3.so is missing).

What is restated and where it comes from
  * frame loop, EOS / min_new_tokens / suppress rules ........ faster_qwen3_tts/generate.py:41-50, :99-199
  * streaming chunk/yield contract and timing dict ............. faster_qwen3_tts/streaming.py:44-188
  * sampling and repetition penalty ............................ faster_qwen3_tts/sampling.py:10-66
  * predictor 15-step order (small_to_mtp on every input) ...... faster_qwen3_tts/predictor_graph.py:115-167
  * decode position = cache_pos + rope_delta on all mrope axes . faster_qwen3_tts/talker_graph.py:198-214
  * layer arithmetic (third-party, NOT in /root/reference): `qwen-tts>=0.1.1` (pyproject.toml:29, unpinned
    upper bound) whose talker/predictor are dense Qwen3 decoder stacks.  Restated from the in-container
    sibling transformers 5.5.0 `models/qwen3/modeling_qwen3.py` (RMSNorm, per-head q/k norm before RoPE,
    GQA eager attention, SwiGLU) — see SURVEY.md §8(c).

PARITY PINNING.  The reference's own implementation of the transformer arithmetic cannot be imported here
(`qwen_tts` absent, no weights, no network), and the reference holds no golden vectors for it: that part is
"parity unpinned" against upstream and is instead pinned against transformers' Qwen3Model run live
(tests/test_oracle_cpu.py) and against committed fixtures (tests/golden/).  Sampling IS pinned against the
reference's own `sampling.py`, loaded by file path when the fixtures were generated
(tests/golden/make_golden.py), and against the known-answer test tests/test_sampling.py:10-21.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# sampling.py restated
# ------------------------------------------------------------------------------------------------
def apply_repetition_penalty(logits: torch.Tensor, token_history: torch.Tensor, repetition_penalty: float) -> torch.Tensor:
    """sampling.py:10-29 — HF-style penalty over the *unique* history, in place."""
    if repetition_penalty == 1.0 or token_history.numel() == 0:
        return logits
    seen = torch.unique(token_history)
    picked = logits[..., seen]
    logits[..., seen] = torch.where(picked > 0, picked / repetition_penalty, picked * repetition_penalty)
    return logits


def sample_logits(
    logits: torch.Tensor,
    *,
    temperature: float,
    top_k: int,
    top_p: float,
    do_sample: bool,
    suppress_mask: Optional[torch.Tensor] = None,
    suppress_tokens: Optional[Iterable[int]] = None,
    generator: Optional[torch.Generator] = None,
) -> torch.Tensor:
    """sampling.py:32-66 — suppress -> (argmax | temperature -> top-k -> top-p -> multinomial)."""
    x = logits.clone()
    neg = float("-inf")
    if suppress_mask is not None:
        x[..., suppress_mask] = neg
    if suppress_tokens:
        x[..., list(suppress_tokens)] = neg
    if not do_sample:
        return x.argmax(dim=-1)
    x = x / temperature
    if top_k > 0:
        kth = torch.topk(x, min(top_k, x.size(-1))).values[..., -1:]
        x = torch.where(x < kth, torch.full_like(x, neg), x)
    if top_p < 1.0:
        srt, idx = torch.sort(x, descending=True)
        cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
        drop = cum > top_p
        drop[..., 0] = False
        srt[drop] = neg
        x = torch.full_like(x, neg).scatter_(-1, idx, srt)
    return torch.multinomial(F.softmax(x, dim=-1), 1, generator=generator).squeeze(-1)


def candidate_set(logits: torch.Tensor, *, temperature: float, top_k: int, top_p: float,
                  suppress_mask=None, suppress_tokens=None) -> torch.Tensor:
    """Boolean mask of the tokens the sampler can still draw (for set-exact parity of top-k / top-p)."""
    x = logits.clone().float()
    neg = float("-inf")
    if suppress_mask is not None:
        x[..., suppress_mask] = neg
    if suppress_tokens:
        x[..., list(suppress_tokens)] = neg
    x = x / temperature
    if top_k > 0:
        kth = torch.topk(x, min(top_k, x.size(-1))).values[..., -1:]
        x = torch.where(x < kth, torch.full_like(x, neg), x)
    if top_p < 1.0:
        srt, idx = torch.sort(x, descending=True, stable=True)
        cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
        drop = cum > top_p
        drop[..., 0] = False
        srt[drop] = neg
        x = torch.full_like(x, neg).scatter_(-1, idx, srt)
    return torch.isfinite(x)


# ------------------------------------------------------------------------------------------------
# dense Qwen3 stack (transformers models/qwen3/modeling_qwen3.py restated)
# ------------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    h = x.to(torch.float32)
    h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps)
    return weight * h.to(dt)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.outer(positions.to(torch.float32).cpu(), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).to(positions.device), emb.sin().to(dtype).to(positions.device)


class OracleStack:
    """One decoder stack over a growing KV cache; batch 1."""

    def __init__(self, cfg, weights: Dict[str, torch.Tensor], prefix: str, attn: str = "eager"):
        self.cfg, self.w, self.prefix, self.attn = cfg, weights, prefix, attn
        self.cache: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * cfg.num_hidden_layers

    def reset(self):
        self.cache = [None] * self.cfg.num_hidden_layers

    def kv_len(self) -> int:
        return 0 if self.cache[0] is None else self.cache[0][0].shape[2]

    def _layer(self, l: int, x: torch.Tensor, cos, sin, mask: torch.Tensor) -> torch.Tensor:
        c, w, p = self.cfg, self.w, f"{self.prefix}.layers.{l}"
        T = x.shape[1]
        h = rms_norm(x, w[f"{p}.input_layernorm.weight"], c.rms_norm_eps)
        q = F.linear(h, w[f"{p}.self_attn.q_proj.weight"]).view(1, T, c.num_attention_heads, c.head_dim)
        k = F.linear(h, w[f"{p}.self_attn.k_proj.weight"]).view(1, T, c.num_key_value_heads, c.head_dim)
        v = F.linear(h, w[f"{p}.self_attn.v_proj.weight"]).view(1, T, c.num_key_value_heads, c.head_dim)
        q = rms_norm(q, w[f"{p}.self_attn.q_norm.weight"], c.rms_norm_eps).transpose(1, 2)
        k = rms_norm(k, w[f"{p}.self_attn.k_norm.weight"], c.rms_norm_eps).transpose(1, 2)
        v = v.transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        if self.cache[l] is not None:
            k = torch.cat([self.cache[l][0], k], dim=2)
            v = torch.cat([self.cache[l][1], v], dim=2)
        self.cache[l] = (k, v)
        g = c.num_attention_heads // c.num_key_value_heads
        kk = k.repeat_interleave(g, dim=1)
        vv = v.repeat_interleave(g, dim=1)
        scale = c.head_dim ** -0.5
        if self.attn == "eager":  # transformers eager_attention_forward
            s = torch.matmul(q, kk.transpose(2, 3)) * scale
            s = s + mask
            pr = F.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)
            o = torch.matmul(pr, vv)
        else:  # fp32 attention (what an sdpa math backend computes)
            s = torch.matmul(q.float(), kk.float().transpose(2, 3)) * scale + mask.float()
            o = torch.matmul(F.softmax(s, dim=-1), vv.float()).to(q.dtype)
        o = o.transpose(1, 2).reshape(1, T, c.num_attention_heads * c.head_dim)
        x = x + F.linear(o, w[f"{p}.self_attn.o_proj.weight"])
        h = rms_norm(x, w[f"{p}.post_attention_layernorm.weight"], c.rms_norm_eps)
        m = F.silu(F.linear(h, w[f"{p}.mlp.gate_proj.weight"])) * F.linear(h, w[f"{p}.mlp.up_proj.weight"])
        return x + F.linear(m, w[f"{p}.mlp.down_proj.weight"])

    def forward(self, x: torch.Tensor, rope_positions: torch.Tensor, n_pad: int = 0, final_norm: bool = True) -> torch.Tensor:
        """x [1,T,H]; rope_positions [T]; keys at cache index < n_pad are masked (left padding)."""
        c = self.cfg
        T, past = x.shape[1], self.kv_len()
        cos, sin = rope_cos_sin(rope_positions, c.head_dim, c.rope_theta, x.dtype)
        qi = torch.arange(past, past + T, device=x.device).view(T, 1)
        ki = torch.arange(0, past + T, device=x.device).view(1, past + T)
        allowed = (ki <= qi) & (ki >= n_pad)
        mask = torch.zeros(T, past + T, dtype=x.dtype, device=x.device).masked_fill(~allowed, torch.finfo(x.dtype).min)
        mask = mask.view(1, 1, T, past + T)
        for l in range(c.num_hidden_layers):
            x = self._layer(l, x, cos, sin, mask)
        if final_norm:
            x = rms_norm(x, self.w[f"{self.prefix}.norm.weight"], c.rms_norm_eps)
        return x


@dataclass
class OraclePolicy:
    do_sample: bool = True
    top_k: int = 50
    top_p: float = 1.0
    temperature: float = 0.9


class OracleTTS:
    """Talker + code predictor + frame loop on plain torch ops (any device, default CPU)."""

    def __init__(self, cfg, weights: Dict[str, torch.Tensor], attn: str = "eager", device="cpu"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.w = {k: v.to(self.device) for k, v in weights.items()}
        self.talker = OracleStack(cfg.talker, self.w, "talker.model", attn)
        self.predictor = OracleStack(cfg.predictor, self.w, "talker.code_predictor.model", attn)
        self.ncb = cfg.predictor.num_codebooks
        self.sub = OraclePolicy()  # predictor policy frozen at "capture" (model.py:124-133)
        self.rope_delta = 0
        self.n_pad = 0

    # -- modules the reference reaches through qwen_tts ------------------------------------------
    def codec_embed(self, ids: torch.Tensor) -> torch.Tensor:
        return F.embedding(ids, self.w["talker.model.codec_embedding.weight"])

    def pred_embed(self, i: int, ids: torch.Tensor) -> torch.Tensor:
        return F.embedding(ids, self.w[f"talker.code_predictor.model.codec_embedding.{i}.weight"])

    def codec_head(self, h: torch.Tensor) -> torch.Tensor:
        return F.linear(h, self.w["talker.codec_head.weight"])

    def small_to_mtp(self, x: torch.Tensor) -> torch.Tensor:
        k = "talker.code_predictor.small_to_mtp_projection.weight"
        if k not in self.w:
            return x
        return F.linear(x, self.w[k], self.w["talker.code_predictor.small_to_mtp_projection.bias"])

    def text_projection(self, ids: torch.Tensor) -> torch.Tensor:
        e = F.embedding(ids, self.w["talker.model.text_embedding.weight"])
        h = F.silu(F.linear(e, self.w["talker.text_projection.linear_fc1.weight"], self.w["talker.text_projection.linear_fc1.bias"]))
        return F.linear(h, self.w["talker.text_projection.linear_fc2.weight"], self.w["talker.text_projection.linear_fc2.bias"])

    # -- operators ------------------------------------------------------------------------------
    def talker_prefill(self, tie: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        """generate.py:107-124 — returns (last-row logits [1,V], past_hidden [1,1,H], prefill_len)."""
        self.talker.reset()
        T = tie.shape[1]
        n_pad = 0 if attention_mask is None else int((attention_mask[0] == 0).sum())
        self.n_pad, self.rope_delta = n_pad, -n_pad
        pos = torch.clamp(torch.arange(T, device=tie.device) - n_pad, min=0)
        h = self.talker.forward(tie, pos, n_pad=n_pad)
        return self.codec_head(h[:, -1, :]), h[:, -1:, :].clone(), T

    def talker_step(self, x: torch.Tensor, position: int) -> torch.Tensor:
        """talker_graph.py:198-214 — one decode step, returns post-norm hidden [1,1,H]."""
        assert self.talker.kv_len() == position, (self.talker.kv_len(), position)
        pos = torch.tensor([position + self.rope_delta], device=x.device)
        return self.talker.forward(x, pos, n_pad=self.n_pad)

    def predictor_loop(self, pred_input: torch.Tensor, generator=None, forced: Optional[torch.Tensor] = None):
        """predictor_graph.py:115-167 — returns (int64 [15], list of 15 logits rows [V_p]).
        `forced` teacher-forces the codes (logits parity without divergence)."""
        s = self.sub
        self.predictor.reset()
        h = self.small_to_mtp(pred_input)
        h = self.predictor.forward(h, torch.arange(2, device=h.device))
        toks, all_logits = [], []
        for i in range(self.ncb):
            logits = F.linear(h[:, -1:, :], self.w[f"talker.code_predictor.lm_head.{i}.weight"])[:, 0, :]
            all_logits.append(logits[0].float())
            tok = sample_logits(logits, temperature=s.temperature, top_k=s.top_k, top_p=s.top_p, do_sample=s.do_sample,
                                generator=generator)
            if forced is not None:
                tok = forced[i].view(1).to(tok.device)
            toks.append(tok[0])
            if i + 1 < self.ncb:
                emb = self.small_to_mtp(self.pred_embed(i, tok.unsqueeze(0)))
                h = self.predictor.forward(emb, torch.tensor([2 + i], device=emb.device))
        return torch.stack(toks), all_logits

    # -- frame loop (generate.py:99-215) -----------------------------------------------------------
    def _suppress_mask(self) -> torch.Tensor:
        V, eos = self.cfg.talker.vocab_size, self.cfg.talker.codec_eos_token_id
        m = torch.zeros(V, dtype=torch.bool, device=self.device)
        m[max(0, V - 1024):] = True
        if 0 <= eos < V:
            m[eos] = False
        return m

    def generate_frames(self, tie, tam, tth, tpe, *, max_new_tokens=2048, min_new_tokens=2, temperature=0.9, top_k=50,
                        top_p=1.0, do_sample=True, repetition_penalty=1.05, max_seq_len=2048, generator=None, trace=None,
                        forced: Optional[torch.Tensor] = None):
        """Generator yielding one int64[16] frame per step; `trace` (dict) collects per-step logits.
        `forced` int64 [N,16] teacher-forces every sampled id (row i = frame i), so a CUDA run can be checked
        step by step without bf16 near-ties making the two runs diverge."""
        eos = self.cfg.talker.codec_eos_token_id
        smask = self._suppress_mask()
        kw = dict(temperature=temperature, top_k=top_k, top_p=top_p, do_sample=do_sample, suppress_mask=smask, generator=generator)
        logits, past_hidden, prefill_len = self.talker_prefill(tie, tam)
        if trace is not None:
            trace["prefill_logits"] = logits[0].float().clone()
            trace["talker_logits"], trace["pred_logits"] = [], []
        token = sample_logits(logits, suppress_tokens=[eos] if min_new_tokens > 0 else None, **kw)
        if trace is not None:
            trace["prefill_token"] = token.clone()
            trace["talker_final"] = []
        gen_step, history = 0, []
        for step_idx in range(max_new_tokens):
            if forced is not None:
                if step_idx >= forced.shape[0]:
                    break
                token = forced[step_idx, 0].view(1).to(self.device)
            if token.item() == eos:
                break
            last_id_hidden = self.codec_embed(token.unsqueeze(1))
            codes, plog = self.predictor_loop(torch.cat((past_hidden, last_id_hidden), dim=1), generator=generator,
                                              forced=None if forced is None else forced[step_idx, 1:].to(self.device))
            frame = torch.cat([token.view(1), codes])
            history.append(token.view(()))
            hs = [last_id_hidden] + [self.pred_embed(i, codes[i].view(1, 1)) for i in range(self.ncb)]
            x = torch.cat(hs, dim=1).sum(1, keepdim=True)
            x = x + (tth[:, gen_step].unsqueeze(1) if gen_step < tth.shape[1] else tpe)
            if trace is not None:
                trace["pred_logits"].append(torch.stack(plog))
            yield frame
            pos = prefill_len + step_idx
            if pos >= max_seq_len - 1:
                break
            hidden = self.talker_step(x, pos)
            logits = self.codec_head(hidden[:, -1, :]).unsqueeze(0)
            if trace is not None:
                trace["talker_logits"].append(logits[0, 0].float().clone())
            if repetition_penalty != 1.0:
                logits = apply_repetition_penalty(logits, torch.stack(history), repetition_penalty)
            if trace is not None:
                fin = logits[0, 0].float().clone()
                fin[smask] = float("-inf")
                if len(history) < min_new_tokens:
                    fin[eos] = float("-inf")
                trace["talker_final"].append(fin)
            token = sample_logits(logits.squeeze(0), suppress_tokens=[eos] if len(history) < min_new_tokens else None, **kw)
            past_hidden = hidden[:, -1:, :].clone()
            gen_step += 1

    def fast_generate(self, tie, tam, tth, tpe, **kw):
        """generate.py:16-215 shape: (int64 [T,16] | None, timing dict)."""
        t0 = time.time()
        frames = list(self.generate_frames(tie, tam, tth, tpe, **kw))
        dt = time.time() - t0
        n = len(frames)
        timing = {"prefill_ms": 0.0, "decode_s": dt, "steps": n, "ms_per_step": dt / n * 1000 if n else 0,
                  "steps_per_s": n / dt if dt > 0 else 0}
        return (torch.stack(frames) if frames else None), timing

    def fast_generate_streaming(self, tie, tam, tth, tpe, *, chunk_size=12, **kw):
        """streaming.py:19-188 shape: yields (int64 [<=chunk,16], timing dict)."""
        buf, total, idx, t0 = [], 0, 0, time.time()
        for frame in self.generate_frames(tie, tam, tth, tpe, **kw):
            buf.append(frame)
            if len(buf) >= chunk_size:
                total += len(buf)
                yield torch.stack(buf), {"chunk_index": idx, "chunk_steps": len(buf), "prefill_ms": 0.0,
                                         "decode_ms": (time.time() - t0) * 1000, "total_steps_so_far": total, "is_final": False}
                buf, idx, t0 = [], idx + 1, time.time()
        if buf:
            total += len(buf)
            yield torch.stack(buf), {"chunk_index": idx, "chunk_steps": len(buf), "prefill_ms": 0.0,
                                     "decode_ms": (time.time() - t0) * 1000, "total_steps_so_far": total, "is_final": True}
