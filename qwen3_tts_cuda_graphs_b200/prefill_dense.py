"""Dense (tensor-core) prefill of the talker prompt — the prefill `talker.forward` of generate.py:107-118.

The persistent decode kernel is a weight-streaming GEMV machine: pushing a T-row prompt through it costs one weight
sweep per 8 rows.  A prompt is a dense contraction instead (2 · 440 M · T FLOP against ONE weight sweep), so rows
[0, T-1) go through the tcgen05 implicit-GEMM kernel of libfq3codec.so (include/fq3_codec.h) with the same fused
epilogues and rounding points the decode kernel uses, plus one fused kernel for the per-head q/k RMSNorm, the rotary
embedding and the K/V scatter straight into the engine's static cache (`fq3_kv_cache_ptr`).  The LAST row then takes
the ordinary decode-kernel path (`fq3_prefill_tail`), which yields logits, the first-token sample, `past_hidden` and the
stream state exactly as `fq3_prefill` does.

Weights are read in place from the engine's arena (row-major [N, K] bf16 is what the GEMM's B operand wants); only the
RMSNorm gammas are copied once to fp32.  Left-padded (batched) prompts keep the chunked path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import codec as _codec
from .codec import F_RESID, F_SWIGLU, K_ATTN, K_GEMM, K_RMSNORM, Op

K_QKNORM_ROPE_KV = 8


class DensePrefill:
    """Op list for rows [0, R) of a prompt, built once for a row capacity and re-used with M patched per call."""

    MIN_ROWS = int(os.environ.get("FQ3_DENSE_MIN_ROWS", "16"))  # below this the chunked path (<= 2 launches of the decode kernel) is at least as fast
    MAX_GRAPHS = 8  # captured (rows, stream) plans kept

    def __init__(self, engine):
        self.eng = engine
        self.lib = _codec.load_lib()
        self.cfg = engine.cfg.talker
        self.cap = 0
        self.ops = None
        self.arr = None
        self.keep = []
        self.graphs = {}
        self.seen = {}
        self.graph_failed = set()
        self.use_graphs = os.environ.get("FQ3C_GRAPH", "1") != "0"
        # full mode (default): ALL T rows go through ALL layers here and only the final norm + codec_head + first-token sample
        # run in the decode kernel (fq3_prefill_head); FQ3_DENSE_FULL=0: rows [0, T-1) here, the last row through the decode
        # kernel's 141 phases (fq3_prefill_tail)
        self.full = os.environ.get("FQ3_DENSE_FULL", "1") != "0"
        a, t = engine.arena, self.cfg
        dev = engine.device
        L = t.num_hidden_layers
        # fp32 copies of the RMSNorm weights (the codec RMSNorm kernel multiplies by an fp32 scale; bf16 -> fp32 is exact)
        self.ln = [(a.view(f"talker.model.layers.{l}.ln1").float().contiguous(), a.view(f"talker.model.layers.{l}.ln2").float().contiguous())
                   for l in range(L)]
        self.base = a.buf.data_ptr()
        self.off = lambda name: self.base + a.offsets[name]
        self.dev = dev

    def _build(self, cap: int):
        t, eng = self.cfg, self.eng
        H, I = t.hidden_size, t.intermediate_size
        nq, nkv, d = t.num_attention_heads, t.num_key_value_heads, t.head_dim
        Q = (nq + 2 * nkv) * d
        dev = self.dev
        bf = lambda r, c: torch.empty(r, c, dtype=torch.bfloat16, device=dev)
        x = [bf(cap, H), bf(cap, H)]
        h, qkv, att, mid = bf(cap, H), bf(cap, Q), bf(cap, nq * d), bf(cap, I)
        self.keep = x + [h, qkv, att, mid]
        self.x0 = x[0]
        ops = []

        def gemm(A, W, N, K, out, flags=0, res=None):
            o = Op()
            o.kind, o.flags = K_GEMM, flags
            o.M, o.N, o.K, o.taps, o.cin = cap, N, K, 1, K
            o.a_rows, o.lda, o.col_mod = cap, A.shape[1], N
            o.ldc = out.shape[1]
            o.A, o.B, o.C = A.data_ptr(), W, out.data_ptr()
            if res is not None:
                o.res, o.ldr = res.data_ptr(), res.shape[1]
            ops.append(o)

        def rmsnorm(A, scale, out):
            o = Op()
            o.kind, o.M, o.N = K_RMSNORM, cap, H
            o.A, o.lda, o.a_rows = A.data_ptr(), A.shape[1], cap
            o.C, o.ldc, o.col_mod = out.data_ptr(), out.shape[1], H
            o.scale, o.f0 = scale.data_ptr(), t.rms_norm_eps
            ops.append(o)

        cur = 0
        L = t.num_hidden_layers
        for l in range(L):
            p = f"talker.model.layers.{l}"
            rmsnorm(x[cur], self.ln[l][0], h)
            gemm(h, self.off(f"{p}.wqkv"), Q, H, qkv)
            o = Op()
            o.kind, o.M, o.N = K_QKNORM_ROPE_KV, cap, Q
            o.A, o.lda, o.a_rows = qkv.data_ptr(), Q, cap
            o.i0, o.i1, o.i2, o.f0 = nq, nkv, 0, t.rms_norm_eps
            o.p0, o.p1 = self.off(f"{p}.qn"), self.off(f"{p}.kn")
            o.B, o.bias = self.off("rope.talker.cos"), self.off("rope.talker.sin")
            o.C = eng.lib.fq3_kv_cache_ptr(eng.h, 0, l, 0)   # patched per stream in run()
            o.C2 = eng.lib.fq3_kv_cache_ptr(eng.h, 0, l, 1)
            o.ldc, o.col_mod = eng.max_seq_len, Q
            ops.append(o)
            if l == L - 1 and not self.full:
                break  # the last row (decode-kernel path) only needs this layer's K/V of the earlier rows
            o = Op()
            o.kind, o.M, o.N = K_ATTN, cap, nq * d
            o.flags = 1  # fq3_codec.h: the prompt-prefill form of FQ3C_ATTN (two-phase kernel; no history rows, head_dim 128)
            o.A, o.lda, o.a_rows = qkv.data_ptr(), Q, cap
            o.C, o.ldc, o.col_mod = att.data_ptr(), nq * d, nq * d
            o.i0, o.i1, o.i2, o.K = nq, nkv, d, eng.max_seq_len  # window >= T: plain causal attention
            ops.append(o)
            gemm(att, self.off(f"{p}.wo"), H, nq * d, x[cur ^ 1], flags=F_RESID, res=x[cur])
            cur ^= 1
            rmsnorm(x[cur], self.ln[l][1], h)
            gemm(h, self.off(f"{p}.wgu"), 2 * I, H, mid, flags=F_SWIGLU)
            gemm(mid, self.off(f"{p}.wdown"), H, I, x[cur ^ 1], flags=F_RESID, res=x[cur])
            cur ^= 1
        for h in self.graphs.values():
            self.lib.fq3c_graph_destroy(h)
        self.graphs, self.seen = {}, {}
        ops = _codec.fuse_row_norms(ops)
        if os.environ.get("FQ3C_SPLITK", "1") != "0":
            _codec.attach_splitk_workspace(ops, dev, self.keep)
        self.x_last = x[cur]  # full mode: the residual stream after the last layer
        self.ops = ops
        self.kv_ops = [(i, o) for i, o in enumerate(ops) if o.kind == K_QKNORM_ROPE_KV]
        self.arr = (Op * len(ops))(*ops)
        self.cap = cap
        self.cur_stream = 0

    def close(self):
        for h in self.graphs.values():
            self.lib.fq3c_graph_destroy(h)
        self.graphs = {}

    def run(self, stream_idx: int, embeds: torch.Tensor, R: int):
        """rows [0, R) of `embeds` (bf16 [T, H], device): fills the K/V cache of every layer for positions [0, R)."""
        if R > self.cap:
            self._build(max(64, (R + 63) // 64 * 64))
        arr = self.arr
        if stream_idx != self.cur_stream:
            for layer, (i, _) in enumerate(self.kv_ops):
                arr[i].C = self.eng.lib.fq3_kv_cache_ptr(self.eng.h, stream_idx, layer, 0)
                arr[i].C2 = self.eng.lib.fq3_kv_cache_ptr(self.eng.h, stream_idx, layer, 1)
            self.cur_stream = stream_idx
        if arr[0].M != R:
            for i in range(len(self.ops)):
                arr[i].M = R
        self.x0[:R].copy_(embeds[:R])
        stream = torch.cuda.current_stream().cuda_stream
        # ~250 small launches: the second prompt of the same length on the same stream replays them as one CUDA graph
        key = (R, stream_idx)
        g = self.graphs.get(key)
        if g is None and self.use_graphs and self.seen.get(key, 0) >= 1 and key not in self.graph_failed:
            h = C.c_void_p()
            if self.lib.fq3c_graph_create(arr, len(self.ops), C.byref(h)) == 0:
                if len(self.graphs) >= self.MAX_GRAPHS:
                    old = self.graphs.pop(next(iter(self.graphs)))  # the oldest graph (dicts keep insertion order)
                    self.lib.fq3c_graph_destroy(old)
                self.graphs[key] = g = h
            else:
                self.graph_failed.add(key)
        self.seen[key] = self.seen.get(key, 0) + 1
        rc = self.lib.fq3c_graph_launch(g, stream) if g is not None else self.lib.fq3c_run(arr, len(self.ops), stream)
        if rc != 0:
            raise _codec.CodecError(self.lib.fq3c_last_error().decode())


def make(engine) -> Optional[DensePrefill]:
    if os.environ.get("FQ3_DENSE_PREFILL", "1") == "0":
        return None
    if engine.cfg.talker.head_dim != 128:
        return None
    return DensePrefill(engine)
