"""Same-box GPU anchor — TEST / BENCH INFRASTRUCTURE ONLY (see oracle/qwen3_tts_oracle.py for who may import oracle/).

The reference's fast path is "HF modules + StaticCache + torch.cuda.CUDAGraph replay" (talker_graph.py:109-147,
predictor_graph.py:115-202).  Its arithmetic lives in the un-vendored `qwen_tts`, so the reference itself cannot run on
this box; this file restates that *execution style* with plain torch ops on the same synthetic weights, so that bench.py
can time "what the reference does" next to the B200-native kernel on the same GPU ("otherwise the
oracle-in-CUDA-graph restatement"):

  * static K/V caches [layers][1, kv_heads, max_seq, d], updated with index_copy_ at a device-resident position
    (talker_graph.py:43-69, :169, :206-211);
  * attention over ALL max_seq slots with an additive mask row selected by position (talker_graph.py:71-92), GQA through
    repeat_interleave like transformers' sdpa path;
  * one CUDA graph for the talker decode step, one for the predictor's 15-step loop with argmax sampling inside
    (predictor_graph.py:115-167), replayed per frame; the 16-row embedding sum, codec_head and the first-codebook argmax are
    eager torch ops between the replays like generate.py:149-199 (no repetition penalty, no EOS test: timing only).

Numerics are the oracle's (bf16 weights, HF rounding points); tests/test_anchor_gpu.py checks a replayed frame against
OracleTTS so the timed work is the real computation.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from .qwen3_tts_oracle import rms_norm, rope_cos_sin, rotate_half


class _StaticStack:
    def __init__(self, cfg, w: Dict[str, torch.Tensor], prefix: str, max_seq: int, device):
        self.cfg, self.w, self.p, self.S = cfg, w, prefix, max_seq
        c = cfg
        self.k = [torch.zeros(1, c.num_key_value_heads, max_seq, c.head_dim, dtype=torch.bfloat16, device=device)
                  for _ in range(c.num_hidden_layers)]
        self.v = [torch.zeros_like(t) for t in self.k]
        pos = torch.arange(max_seq, device=device)
        self.cos, self.sin = rope_cos_sin(pos, c.head_dim, c.rope_theta, torch.bfloat16)  # [S, d]
        # additive mask rows: row p lets a query at position p see slots [0, p] (talker_graph.py:71-92)
        m = torch.full((max_seq, max_seq), float("-inf"), device=device)
        self.mask = torch.triu(m, diagonal=1).to(torch.bfloat16)

    def zero(self):
        for t in self.k + self.v:
            t.zero_()

    def step(self, x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        """x [1, T, H] at positions pos int64 [T] (device tensor) -> post-final-norm hidden [1, T, H]."""
        c, w = self.cfg, self.w
        T = x.shape[1]
        cos, sin = self.cos.index_select(0, pos), self.sin.index_select(0, pos)
        mask = self.mask.index_select(0, pos).view(1, 1, T, self.S)
        rep = c.num_attention_heads // c.num_key_value_heads
        for l in range(c.num_hidden_layers):
            p = f"{self.p}.layers.{l}"
            h = rms_norm(x, w[f"{p}.input_layernorm.weight"], c.rms_norm_eps)
            q = F.linear(h, w[f"{p}.self_attn.q_proj.weight"]).view(1, T, c.num_attention_heads, c.head_dim)
            k = F.linear(h, w[f"{p}.self_attn.k_proj.weight"]).view(1, T, c.num_key_value_heads, c.head_dim)
            v = F.linear(h, w[f"{p}.self_attn.v_proj.weight"]).view(1, T, c.num_key_value_heads, c.head_dim)
            q = rms_norm(q, w[f"{p}.self_attn.q_norm.weight"], c.rms_norm_eps).transpose(1, 2)
            k = rms_norm(k, w[f"{p}.self_attn.k_norm.weight"], c.rms_norm_eps).transpose(1, 2)
            v = v.transpose(1, 2)
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
            self.k[l].index_copy_(2, pos, k)
            self.v[l].index_copy_(2, pos, v)
            kk = self.k[l].repeat_interleave(rep, dim=1)
            vv = self.v[l].repeat_interleave(rep, dim=1)
            o = F.scaled_dot_product_attention(q, kk, vv, attn_mask=mask)
            o = o.transpose(1, 2).reshape(1, T, c.num_attention_heads * c.head_dim)
            x = x + F.linear(o, w[f"{p}.self_attn.o_proj.weight"])
            h = rms_norm(x, w[f"{p}.post_attention_layernorm.weight"], c.rms_norm_eps)
            m = F.silu(F.linear(h, w[f"{p}.mlp.gate_proj.weight"])) * F.linear(h, w[f"{p}.mlp.up_proj.weight"])
            x = x + F.linear(m, w[f"{p}.mlp.down_proj.weight"])
        return rms_norm(x, w[f"{self.p}.norm.weight"], c.rms_norm_eps)


class GraphAnchor:
    """Greedy frame loop as two CUDA-graph replays + eager glue per frame, batch 1."""

    def __init__(self, cfg, weights: Dict[str, torch.Tensor], max_seq_len: int = 2048, device="cuda"):
        self.cfg = cfg
        dev = torch.device(device)
        self.dev = dev
        self.w = {k: v.to(dev) for k, v in weights.items()}
        self.ncb = cfg.predictor.num_codebooks
        self.talker = _StaticStack(cfg.talker, self.w, "talker.model", max_seq_len, dev)
        self.pred = _StaticStack(cfg.predictor, self.w, "talker.code_predictor.model", self.ncb + 2, dev)
        H = cfg.talker.hidden_size
        self.t_in = torch.zeros(1, 1, H, dtype=torch.bfloat16, device=dev)
        self.t_pos = torch.zeros(1, dtype=torch.long, device=dev)
        self.t_out = torch.zeros(1, 1, H, dtype=torch.bfloat16, device=dev)
        self.p_in = torch.zeros(1, 2, H, dtype=torch.bfloat16, device=dev)
        self.p_out = torch.zeros(self.ncb, dtype=torch.long, device=dev)
        self.p_pos2 = torch.arange(2, device=dev)
        self.p_pos1 = [torch.tensor([2 + i], device=dev) for i in range(self.ncb)]
        self.g_talker = self.g_pred = None

    def _s2m(self, x):
        k = "talker.code_predictor.small_to_mtp_projection.weight"
        return x if k not in self.w else F.linear(x, self.w[k], self.w["talker.code_predictor.small_to_mtp_projection.bias"])

    def _talker_body(self):
        self.t_out.copy_(self.talker.step(self.t_in, self.t_pos))

    def _pred_body(self):
        """predictor_graph.py:115-167 with greedy sampling in the graph."""
        self.pred.zero()  # predictor_graph.py:212
        h = self.pred.step(self._s2m(self.p_in), self.p_pos2)
        for i in range(self.ncb):
            logits = F.linear(h[:, -1:, :], self.w[f"talker.code_predictor.lm_head.{i}.weight"])[:, 0, :]
            tok = logits.argmax(dim=-1)
            self.p_out[i] = tok[0]
            if i + 1 < self.ncb:
                e = F.embedding(tok.view(1, 1), self.w[f"talker.code_predictor.model.codec_embedding.{i}.weight"])
                h = self.pred.step(self._s2m(e), self.p_pos1[i])

    @torch.inference_mode()
    def capture(self, warmup: int = 3):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._talker_body()
                self._pred_body()
        torch.cuda.current_stream().wait_stream(s)
        self.g_talker, self.g_pred = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_talker):
            self._talker_body()
        with torch.cuda.graph(self.g_pred):
            self._pred_body()

    @torch.inference_mode()
    def prefill(self, tie: torch.Tensor):
        """Eager full-sequence pass (generate.py:107-118): fills the static cache, returns (first token, past_hidden)."""
        self.talker.zero()
        T = tie.shape[1]
        h = self.talker.step(tie.to(self.dev), torch.arange(T, device=self.dev))
        logits = F.linear(h[:, -1, :], self.w["talker.codec_head.weight"])
        return logits.argmax(dim=-1), h[:, -1:, :].clone(), T

    @torch.inference_mode()
    def frame(self, token: torch.Tensor, past_hidden: torch.Tensor, position: int, text_row: torch.Tensor, use_graphs: bool = True):
        """One iteration of generate.py:149-199 (greedy, no penalty): returns (codes int64[16], next token, past_hidden)."""
        cur = F.embedding(token.view(1, 1), self.w["talker.model.codec_embedding.weight"])
        self.p_in.copy_(torch.cat((past_hidden, cur), dim=1))
        self.g_pred.replay() if use_graphs else self._pred_body()
        codes = self.p_out.clone()
        rows: List[torch.Tensor] = [cur] + [
            F.embedding(codes[i].view(1, 1), self.w[f"talker.code_predictor.model.codec_embedding.{i}.weight"]) for i in range(self.ncb)]
        x = torch.cat(rows, dim=1).sum(1, keepdim=True) + text_row
        self.t_in.copy_(x)
        self.t_pos.fill_(position)
        self.g_talker.replay() if use_graphs else self._talker_body()
        hidden = self.t_out.clone()
        nxt = F.linear(hidden[:, -1, :], self.w["talker.codec_head.weight"]).argmax(dim=-1)
        return torch.cat([token.view(1), codes]), nxt, hidden

    @torch.inference_mode()
    def time_frames(self, tie: torch.Tensor, tpe: torch.Tensor, n_frames: int) -> float:
        """ms per frame of the replayed loop (CUDA events, after prefill)."""
        tok, ph, T = self.prefill(tie)
        pad = tpe.to(self.dev).view(1, 1, -1)
        for i in range(3):
            _, tok, ph = self.frame(tok, ph, T + i, pad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_frames):
            _, tok, ph = self.frame(tok, ph, T + 3 + i, pad)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_frames
