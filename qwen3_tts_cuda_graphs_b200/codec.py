"""12 Hz speech-tokenizer DECODER on libfq3codec.so — `speech_tokenizer.decode` of the reference's base model.

The reference calls `m.speech_tokenizer.decode({"audio_codes": codes[1,T,16]}) -> ([wav], sr)` eagerly through
cuDNN/ATen/cuBLAS (faster_qwen3_tts/model.py:642,782,811,884,971,988,1054,1136,1153; module body in the un-vendored
`qwen_tts`, architecture per SURVEY.md §8c: split-RVQ dequantiser -> causal conv -> 8-layer sliding-window transformer
-> 2x(transposed conv + ConvNeXt) -> conv7 -> 4x{SnakeBeta, transposed conv, 3 residual units} -> SnakeBeta -> conv7).

Here the host lowers the decoder for a given frame count T to a flat list of `fq3c_op` records over channels-last
bf16 activations (include/fq3_codec.h) and hands the list to ONE C call, `fq3c_run`.  Every dense contraction is the
same tensor-core implicit-GEMM kernel:
  * causal conv k, dilation d  -> taps k, row offsets -(k-1-j)*d, weights [C_out, k*C_in]
  * transposed conv k=s (ConvNeXt stage) -> one tap, N = s*C_out, output rows [T, s*C_out] == [T*s, C_out]
  * transposed conv k=2s, stride s (vocoder) -> two taps, N = s*C_out: trimmed s on the right only (causal; config
    trans_conv_trim = "right", T*s rows, taps (0, -1)) or s on both sides like the sibling ("both", (T-1)*s rows, taps (+1, 0))
SnakeBeta is fused into the epilogue of the GEMM that produces its input; bias, GELU, layer-scale and residual adds
are epilogue flags as well.  Plans (op list + activation buffers) are cached per T, so steady-state streaming
decode issues no allocation.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import shutil
from typing import Dict, List, Optional, Tuple

import torch

from . import build as _build
from .config import CodecDecoderConfig

K_GEMM, K_RVQ, K_RMSNORM, K_ROPE, K_ATTN, K_DWCONV, K_LAYERNORM, K_SNAKE = range(8)
K_COPY, K_ADVANCE, K_ROLL = 9, 10, 11
F_BIAS, F_GELU, F_RESID, F_SCALE, F_SWIGLU, F_CLAMP, F_OUT_F32, F_SNAKE2, F_SILU = 1, 2, 4, 8, 16, 32, 64, 128, 256


class Op(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("taps", C.c_int32), ("cin", C.c_int32), ("a_rows", C.c_int32), ("col_mod", C.c_int32),
        ("tap_off", C.c_int32 * 8), ("lda", C.c_int32), ("ldc", C.c_int32), ("ldr", C.c_int32),
        ("i0", C.c_int32), ("i1", C.c_int32), ("i2", C.c_int32), ("f0", C.c_float), ("f1", C.c_float),
        ("A", C.c_void_p), ("B", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p), ("scale", C.c_void_p),
        ("p0", C.c_void_p), ("p1", C.c_void_p), ("C", C.c_void_p), ("C2", C.c_void_p),
        ("ws", C.c_void_p), ("ws_bytes", C.c_int64), ("m_begin", C.c_int32), ("reserved", C.c_int32),
        ("norm_out", C.c_void_p), ("norm_w", C.c_void_p), ("norm_ld", C.c_int32), ("norm_eps", C.c_float),
    ]


SPLITK_MAX_ROWS = 512          # GEMMs up to this many rows may be split over K (include/fq3_codec.h: fq3c_op.ws)
SPLITK_WS_CAP = 64 << 20       # bytes of fp32 workspace one plan may pin


VOCODER_DILATIONS = (1, 3, 9)  # the three residual units of a decoder block (7-tap causal convs)


def vocoder_tail_starts(rows0: int, upsample_rates, skip: int, trim: str = "both"):
    """First output row every vocoder op must produce so that waveform samples [skip, n) come out exactly as in a full
    decode.  Layers, in order: dec0 (7-tap causal conv over `rows0` rows); per block i: transposed conv (kernel 2r, stride r;
    trim "both": output row m*r + k reads input rows m and m+1, (rows-1)*r output rows; trim "right": it reads rows m-1 and m,
    rows*r output rows), then three residual units (7-tap causal conv with dilation d -> reads back 6*d rows; 1x1 conv +
    residual -> same row); final 7-tap causal conv.
    Returns (dec0_begin, tconv_begin[i] in the transposed conv's own (input) rows, unit_begin[i][j], final_begin)."""
    nblk = len(upsample_rates)
    right = trim == "right"
    lvl_rows = [rows0]
    for r in upsample_rates:
        lvl_rows.append(lvl_rows[-1] * r if right else (lvl_rows[-1] - 1) * r)
    s_need = max(0, min(int(skip), max(lvl_rows[-1] - 1, 0)))
    fin_begin = s_need
    s_need -= 6
    unit_begin = [[0] * 3 for _ in range(nblk)]
    tconv_begin = [0] * nblk
    for i in range(nblk - 1, -1, -1):
        for j in (2, 1, 0):
            unit_begin[i][j] = max(0, s_need)
            s_need -= 6 * VOCODER_DILATIONS[j]
        tconv_begin[i] = max(0, s_need) // upsample_rates[i]
        s_need = tconv_begin[i] - (1 if right else 0)
    return max(0, s_need), tconv_begin, unit_begin, fin_begin


def fuse_row_norms(ops: list) -> list:
    """An RMSNORM op that reads exactly the rows the GEMM before it wrote becomes that GEMM's fused norm (fq3c_op.norm_out):
    inside the split-K reduction when the GEMM is split, one small launch otherwise — same arithmetic, fewer launches (two per
    transformer layer of the prompt prefill and of the codec).  FQ3C_FUSE_NORM=0 keeps the ops apart."""
    if os.environ.get("FQ3C_FUSE_NORM", "1") == "0":
        return ops
    out = []
    for o in ops:
        g = out[-1] if out else None
        if (o.kind == K_RMSNORM and g is not None and g.kind == K_GEMM and not (g.flags & (F_SWIGLU | F_OUT_F32)) and not g.norm_out
                and g.m_begin == 0 and o.A == g.C and o.lda == g.ldc and o.N == g.N and o.M == g.M and o.a_rows == g.M):
            g.norm_out, g.norm_w, g.norm_ld, g.norm_eps = o.C, o.scale, o.ldc, o.f0
            continue
        out.append(o)
    return out


def attach_splitk_workspace(ops, device, keep) -> None:
    """One fp32 workspace per stream-ordered op list: every short GEMM of the list may use it for split-K partial tiles."""
    need = 0
    for o in ops:
        if o.kind == K_GEMM and 0 < o.M <= SPLITK_MAX_ROWS and o.K >= 512:
            need = max(need, 8 * o.M * ((o.N + 7) // 8 * 8) * 4)
    if need == 0:
        return
    ws = torch.empty(min(need, SPLITK_WS_CAP) // 4, dtype=torch.float32, device=device)
    keep.append(ws)
    for o in ops:
        if o.kind == K_GEMM and 0 < o.M <= SPLITK_MAX_ROWS and o.K >= 512:
            o.ws, o.ws_bytes = ws.data_ptr(), ws.numel() * 4


_lib = None


class CodecError(RuntimeError):
    pass


def load_lib():
    """dlopen libfq3codec.so (built in-tree by build.py).  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.lib_path("libfq3codec.so")
    if (not os.path.exists(path) or _build.is_stale("libfq3codec.so")) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        _build.build_lib("libfq3codec.so")
    if not os.path.exists(path):
        raise CodecError(f"{path} is missing: run `python -m qwen3_tts_cuda_graphs_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.fq3c_abi_version.restype = C.c_int
    lib.fq3c_last_error.restype = C.c_char_p
    lib.fq3c_run.restype = C.c_int
    lib.fq3c_run.argtypes = [C.POINTER(Op), C.c_int, C.c_void_p]
    lib.fq3c_launch_count.restype = C.c_int64
    lib.fq3c_graph_create.restype = C.c_int
    lib.fq3c_graph_create.argtypes = [C.POINTER(Op), C.c_int, C.POINTER(C.c_void_p)]
    lib.fq3c_graph_launch.restype = C.c_int
    lib.fq3c_graph_launch.argtypes = [C.c_void_p, C.c_void_p]
    lib.fq3c_graph_destroy.restype = C.c_int
    lib.fq3c_graph_destroy.argtypes = [C.c_void_p]
    if lib.fq3c_abi_version() != 5:
        raise CodecError("libfq3codec.so ABI version mismatch")
    _lib = lib
    return lib


# ------------------------------------------------------------------------------------------------
# synthetic weights (no checkpoint offline): names follow the sibling transformers module tree
# ------------------------------------------------------------------------------------------------
def codec_tensor_specs(c: CodecDecoderConfig, pre_conv_kernel: int = 3) -> List[tuple]:
    H, I, d = c.hidden_size, c.intermediate_size, c.head_dim
    s: List[tuple] = []
    for g in range(c.num_quantizers):
        s.append((f"quantizer.codebook.{g}", (c.codebook_size, c.codebook_dim), "emb"))
    s += [("quantizer.rvq_first.output_proj.weight", (c.latent_dim, c.codebook_dim), "lin"),
          ("quantizer.rvq_rest.output_proj.weight", (c.latent_dim, c.codebook_dim), "lin"),
          ("pre_conv.conv.weight", (H, c.latent_dim, pre_conv_kernel), "conv"), ("pre_conv.conv.bias", (H,), "bias")]
    for l in range(c.num_hidden_layers):
        p = f"pre_transformer.layers.{l}"
        s += [(f"{p}.input_layernorm.weight", (H,), "norm"),
              (f"{p}.self_attn.q_proj.weight", (c.num_attention_heads * d, H), "lin"),
              (f"{p}.self_attn.k_proj.weight", (c.num_key_value_heads * d, H), "lin"),
              (f"{p}.self_attn.v_proj.weight", (c.num_key_value_heads * d, H), "lin"),
              (f"{p}.self_attn.o_proj.weight", (H, c.num_attention_heads * d), "lin"),
              (f"{p}.self_attn_layer_scale.scale", (H,), "lscale"),
              (f"{p}.post_attention_layernorm.weight", (H,), "norm"),
              (f"{p}.mlp.gate_proj.weight", (I, H), "lin"), (f"{p}.mlp.up_proj.weight", (I, H), "lin"),
              (f"{p}.mlp.down_proj.weight", (H, I), "lin"), (f"{p}.mlp_layer_scale.scale", (H,), "lscale")]
    s.append(("pre_transformer.norm.weight", (H,), "norm"))
    for i, f in enumerate(c.upsampling_ratios):
        p = f"upsample.{i}"
        s += [(f"{p}.0.conv.weight", (H, H, f), "tconv"), (f"{p}.0.conv.bias", (H,), "bias"),
              (f"{p}.1.dwconv.conv.weight", (H, 1, 7), "dw"), (f"{p}.1.dwconv.conv.bias", (H,), "bias"),
              (f"{p}.1.norm.weight", (H,), "norm"), (f"{p}.1.norm.bias", (H,), "bias"),
              (f"{p}.1.pwconv1.weight", (4 * H, H), "lin"), (f"{p}.1.pwconv1.bias", (4 * H,), "bias"),
              (f"{p}.1.pwconv2.weight", (H, 4 * H), "lin"), (f"{p}.1.pwconv2.bias", (H,), "bias"),
              (f"{p}.1.gamma", (H,), "gamma")]
    D = c.decoder_dim
    s += [("decoder.0.conv.weight", (D, H, 7), "conv"), ("decoder.0.conv.bias", (D,), "bias")]
    for i, r in enumerate(c.upsample_rates):
        cin, cout = D // 2 ** i, D // 2 ** (i + 1)
        p = f"decoder.{i + 1}.block"
        s += [(f"{p}.0.alpha", (cin,), "snake"), (f"{p}.0.beta", (cin,), "snake"),
              (f"{p}.1.conv.weight", (cin, cout, 2 * r), "tconv"), (f"{p}.1.conv.bias", (cout,), "bias")]
        for j in range(3):
            q = f"{p}.{j + 2}"
            s += [(f"{q}.act1.alpha", (cout,), "snake"), (f"{q}.act1.beta", (cout,), "snake"),
                  (f"{q}.conv1.conv.weight", (cout, cout, 7), "conv"), (f"{q}.conv1.conv.bias", (cout,), "bias"),
                  (f"{q}.act2.alpha", (cout,), "snake"), (f"{q}.act2.beta", (cout,), "snake"),
                  (f"{q}.conv2.conv.weight", (cout, cout, 1), "conv"), (f"{q}.conv2.conv.bias", (cout,), "bias")]
    n = len(c.upsample_rates) + 1
    out_dim = D // 2 ** len(c.upsample_rates)
    s += [(f"decoder.{n}.alpha", (out_dim,), "snake"), (f"decoder.{n}.beta", (out_dim,), "snake"),
          (f"decoder.{n + 1}.conv.weight", (1, out_dim, 7), "conv"), (f"decoder.{n + 1}.conv.bias", (1,), "bias")]
    return s


def init_codec_synthetic(c: CodecDecoderConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random init with fan-in scaling (activations stay O(1) through ~60 layers, so the waveform is a
    non-degenerate signal in [-1, 1] and parity tolerances mean something)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in codec_tensor_specs(c):
        if kind == "norm":
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "bias":
            w = 0.02 * torch.randn(shape, generator=g)
        elif kind == "lscale":
            w = torch.full(shape, 0.3) + 0.05 * torch.randn(shape, generator=g)
        elif kind == "gamma":
            w = torch.full(shape, 0.3) + 0.05 * torch.randn(shape, generator=g)
        elif kind == "snake":
            w = 0.3 * torch.randn(shape, generator=g)
        elif kind == "emb":
            w = torch.randn(shape, generator=g) / math.sqrt(c.num_quantizers)
        elif kind == "dw":
            w = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif kind == "tconv":  # [cin, cout, k]; each output sample sums k/stride taps of cin channels
            taps = 1 if shape[2] <= 2 else 2
            w = torch.randn(shape, generator=g) / math.sqrt(shape[0] * taps)
        elif kind == "conv":  # [cout, cin, k]
            w = torch.randn(shape, generator=g) / math.sqrt(shape[1] * shape[2])
            if name.endswith("conv2.conv.weight"):
                w = 0.3 * w  # residual branches stay small: the stack neither explodes nor saturates the clamp
            if shape[0] == 1:
                w = 0.25 * w  # output conv: waveform rms ~0.3
        else:  # lin [out, in]
            w = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        # the CUDA path holds matrices in bf16: round here so oracle and engine see identical parameters
        out[name] = w.to(torch.bfloat16).to(dtype) if kind in ("lin", "conv", "tconv", "emb") else w.to(dtype)
    return out


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
class _Plan:
    def __init__(self):
        self.ops: List[Op] = []
        self.keep: List[torch.Tensor] = []
        self.codes: Optional[torch.Tensor] = None
        self.wav: Optional[torch.Tensor] = None
        self.arr = None
        self.runs = 0
        self.graph = None          # fq3c_graph handle once the plan has been captured
        self.graph_failed = False


class CodecDecoder:
    """Weights repacked for the implicit-GEMM kernels + per-T launch plans."""

    MAX_LONG_PLANS = 6

    def __init__(self, cfg: CodecDecoderConfig, weights: Dict[str, torch.Tensor], device):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the fq3 codec decoder runs on CUDA only (no CPU fallback)")
        self.lib = load_lib()
        self.use_graphs = os.environ.get("FQ3C_GRAPH", "1") != "0"
        self._plans: Dict[int, _Plan] = {}
        self.g: Dict[str, torch.Tensor] = {}
        self._pack(weights)

    # ---- weight packing -------------------------------------------------------------------------
    def _bf(self, t):
        return t.to(self.device, torch.bfloat16).contiguous()

    def _f32(self, t):
        return t.to(self.device, torch.float32).contiguous()

    def _conv_w(self, w):  # [cout, cin, k] -> [cout, k*cin]
        return self._bf(w.permute(0, 2, 1).reshape(w.shape[0], -1))

    def _pack(self, w):
        c, g = self.cfg, self.g
        g["cb"] = self._bf(torch.stack([w[f"quantizer.codebook.{q}"] for q in range(c.num_quantizers)]))
        g["rvq_proj"] = self._bf(torch.cat([w["quantizer.rvq_first.output_proj.weight"],
                                            w["quantizer.rvq_rest.output_proj.weight"]], dim=1))
        g["pre_conv.w"] = self._conv_w(w["pre_conv.conv.weight"])
        g["pre_conv.b"] = self._f32(w["pre_conv.conv.bias"])
        for l in range(c.num_hidden_layers):
            p = f"pre_transformer.layers.{l}"
            g[f"{p}.ln1"] = self._f32(w[f"{p}.input_layernorm.weight"])
            g[f"{p}.wqkv"] = self._bf(torch.cat([w[f"{p}.self_attn.q_proj.weight"], w[f"{p}.self_attn.k_proj.weight"],
                                                 w[f"{p}.self_attn.v_proj.weight"]], 0))
            g[f"{p}.wo"] = self._bf(w[f"{p}.self_attn.o_proj.weight"])
            g[f"{p}.ls1"] = self._f32(w[f"{p}.self_attn_layer_scale.scale"])
            g[f"{p}.ln2"] = self._f32(w[f"{p}.post_attention_layernorm.weight"])
            gate, up = w[f"{p}.mlp.gate_proj.weight"], w[f"{p}.mlp.up_proj.weight"]
            g[f"{p}.wgu"] = self._bf(torch.stack([gate, up], dim=1).reshape(2 * gate.shape[0], gate.shape[1]))
            g[f"{p}.wdown"] = self._bf(w[f"{p}.mlp.down_proj.weight"])
            g[f"{p}.ls2"] = self._f32(w[f"{p}.mlp_layer_scale.scale"])
        g["norm"] = self._f32(w["pre_transformer.norm.weight"])
        for i, f in enumerate(c.upsampling_ratios):
            p = f"upsample.{i}"
            tw = w[f"{p}.0.conv.weight"]  # [cin, cout, f] -> rows (phase, cout), cols cin
            g[f"{p}.t.w"] = self._bf(tw.permute(2, 1, 0).reshape(f * tw.shape[1], tw.shape[0]))
            g[f"{p}.t.b"] = self._f32(w[f"{p}.0.conv.bias"])
            g[f"{p}.dw.w"] = self._f32(w[f"{p}.1.dwconv.conv.weight"].reshape(-1, 7))
            g[f"{p}.dw.b"] = self._f32(w[f"{p}.1.dwconv.conv.bias"])
            g[f"{p}.ln.w"] = self._f32(w[f"{p}.1.norm.weight"])
            g[f"{p}.ln.b"] = self._f32(w[f"{p}.1.norm.bias"])
            g[f"{p}.pw1.w"] = self._bf(w[f"{p}.1.pwconv1.weight"])
            g[f"{p}.pw1.b"] = self._f32(w[f"{p}.1.pwconv1.bias"])
            g[f"{p}.pw2.w"] = self._bf(w[f"{p}.1.pwconv2.weight"])
            g[f"{p}.pw2.b"] = self._f32(w[f"{p}.1.pwconv2.bias"])
            g[f"{p}.gamma"] = self._f32(w[f"{p}.1.gamma"])
        g["dec0.w"] = self._conv_w(w["decoder.0.conv.weight"])
        g["dec0.b"] = self._f32(w["decoder.0.conv.bias"])

        def snake(prefix, name):
            g[f"{name}.ea"] = self._f32(torch.exp(w[f"{prefix}.alpha"].float()))
            g[f"{name}.ib"] = self._f32(1.0 / (torch.exp(w[f"{prefix}.beta"].float()) + 1e-9))

        for i, r in enumerate(c.upsample_rates):
            p = f"decoder.{i + 1}.block"
            snake(f"{p}.0", f"{p}.0")
            tw = w[f"{p}.1.conv.weight"]  # [cin, cout, 2r]
            cin, cout = tw.shape[0], tw.shape[1]
            a = tw[:, :, :r].permute(2, 1, 0).reshape(r * cout, cin)   # tap 0: x[m+1]
            b = tw[:, :, r:].permute(2, 1, 0).reshape(r * cout, cin)   # tap 1: x[m]
            g[f"{p}.1.w"] = self._bf(torch.cat([a, b], dim=1))
            g[f"{p}.1.b"] = self._f32(w[f"{p}.1.conv.bias"])
            for j in range(3):
                q = f"{p}.{j + 2}"
                snake(f"{q}.act1", f"{q}.act1")
                snake(f"{q}.act2", f"{q}.act2")
                g[f"{q}.c1.w"] = self._conv_w(w[f"{q}.conv1.conv.weight"])
                g[f"{q}.c1.b"] = self._f32(w[f"{q}.conv1.conv.bias"])
                g[f"{q}.c2.w"] = self._conv_w(w[f"{q}.conv2.conv.weight"])
                g[f"{q}.c2.b"] = self._f32(w[f"{q}.conv2.conv.bias"])
        n = len(c.upsample_rates) + 1
        snake(f"decoder.{n}", "final")
        g["final.w"] = self._conv_w(w[f"decoder.{n + 1}.conv.weight"])
        g["final.b"] = self._f32(w[f"decoder.{n + 1}.conv.bias"])

    # ---- plan construction ----------------------------------------------------------------------
    def _buf(self, plan: _Plan, rows: int, cols: int, dtype=torch.bfloat16) -> torch.Tensor:
        t = torch.empty(max(rows, 1), cols, dtype=dtype, device=self.device)
        plan.keep.append(t)
        return t

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return None if t is None else t.data_ptr()

    def _gemm(self, plan, A, W, M, N, cin, taps=1, tap_off=(0,), flags=0, bias=None, res=None, scale=None, snake=None,
              col_mod=None, out=None, out2=None, out_f32=False, m_begin=0):
        o = Op()
        o.m_begin = max(0, min(int(m_begin), M - 1))
        o.kind, o.flags = K_GEMM, flags | (F_OUT_F32 if out_f32 else 0)
        o.M, o.N, o.K, o.taps, o.cin = M, N, taps * cin, taps, cin
        o.a_rows, o.lda = A.shape[0], A.shape[1]
        o.col_mod = col_mod or N
        for i, t in enumerate(tap_off):
            o.tap_off[i] = t
        width = N // 2 if flags & F_SWIGLU else N
        if out is None:
            out = self._buf(plan, M, width, torch.float32 if out_f32 else torch.bfloat16)
        o.ldc = out.shape[1]
        o.A, o.B, o.C = A.data_ptr(), W.data_ptr(), out.data_ptr()
        plan.keep += [t for t in (A, W, bias, res, scale, out) if t is not None]  # ops hold raw pointers
        if bias is not None:
            o.flags |= F_BIAS
            o.bias = bias.data_ptr()
        if res is not None:
            o.flags |= F_RESID
            o.res, o.ldr = res.data_ptr(), res.shape[1]
        if scale is not None:
            o.flags |= F_SCALE
            o.scale = scale.data_ptr()
        if snake is not None:
            o.flags |= F_SNAKE2
            if out2 is None:
                out2 = self._buf(plan, M, width)
            o.p0, o.p1, o.C2 = self.g[snake + ".ea"].data_ptr(), self.g[snake + ".ib"].data_ptr(), out2.data_ptr()
        plan.ops.append(o)
        return out, out2

    def _simple(self, plan, kind, A, M, N, out=None, **kw):
        o = Op()
        o.kind, o.M, o.N = kind, M, N
        o.A, o.lda, o.a_rows = A.data_ptr(), A.shape[1], A.shape[0]
        if out is None:
            out = self._buf(plan, M, N)
        o.C, o.ldc = out.data_ptr(), out.shape[1]
        o.col_mod = N
        for k, v in kw.items():
            setattr(o, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        plan.keep += [A, out] + [v for v in kw.values() if isinstance(v, torch.Tensor)]
        plan.ops.append(o)
        return out

    def _build(self, T: int, skip: int = 0) -> _Plan:
        c, g = self.cfg, self.g
        plan = _Plan()
        H, d, nh, nkv = c.hidden_size, c.head_dim, c.num_attention_heads, c.num_key_value_heads
        plan.codes = torch.zeros(T, c.num_quantizers, dtype=torch.int64, device=self.device)
        # 1. split-RVQ dequantiser: gather+sum, then both output projections as one GEMM over [first | rest]
        q = self._simple(plan, K_RVQ, plan.codes, T, 2 * c.codebook_dim, B=g["cb"], K=c.codebook_size,
                         i0=c.num_semantic_quantizers, i1=c.num_quantizers, i2=c.codebook_dim)
        plan.ops[-1].lda = c.num_quantizers
        x, _ = self._gemm(plan, q, g["rvq_proj"], T, c.latent_dim, 2 * c.codebook_dim)
        # 2. causal pre-conv
        k = g["pre_conv.w"].shape[1] // c.latent_dim
        x, _ = self._gemm(plan, x, g["pre_conv.w"], T, H, c.latent_dim, taps=k, tap_off=[j - (k - 1) for j in range(k)],
                          bias=g["pre_conv.b"])
        # 3. sliding-window transformer
        qkv_dim = (nh + 2 * nkv) * d
        for l in range(c.num_hidden_layers):
            p = f"pre_transformer.layers.{l}"
            h = self._simple(plan, K_RMSNORM, x, T, H, scale=g[f"{p}.ln1"], f0=c.rms_norm_eps)
            qkv, _ = self._gemm(plan, h, g[f"{p}.wqkv"], T, qkv_dim, H)
            self._simple(plan, K_ROPE, qkv, T, qkv_dim, out=qkv, i0=nh + nkv, i1=d, i2=0, f0=c.rope_theta)
            a = self._simple(plan, K_ATTN, qkv, T, nh * d, i0=nh, i1=nkv, i2=d, K=c.sliding_window)
            x, _ = self._gemm(plan, a, g[f"{p}.wo"], T, H, nh * d, scale=g[f"{p}.ls1"], res=x)
            h = self._simple(plan, K_RMSNORM, x, T, H, scale=g[f"{p}.ln2"], f0=c.rms_norm_eps)
            m, _ = self._gemm(plan, h, g[f"{p}.wgu"], T, 2 * c.intermediate_size, H, flags=F_SWIGLU)
            x, _ = self._gemm(plan, m, g[f"{p}.wdown"], T, H, c.intermediate_size, scale=g[f"{p}.ls2"], res=x)
        x = self._simple(plan, K_RMSNORM, x, T, H, scale=g["norm"], f0=c.rms_norm_eps)
        # 4. ConvNeXt upsampling stages
        rows = T
        for i, f in enumerate(c.upsampling_ratios):
            p = f"upsample.{i}"
            y, _ = self._gemm(plan, x, g[f"{p}.t.w"], rows, f * H, H, bias=g[f"{p}.t.b"], col_mod=H)
            rows *= f
            x = y.view(rows, H)
            h = self._simple(plan, K_DWCONV, x, rows, H, B=g[f"{p}.dw.w"], bias=g[f"{p}.dw.b"], taps=7)
            h = self._simple(plan, K_LAYERNORM, h, rows, H, scale=g[f"{p}.ln.w"], bias=g[f"{p}.ln.b"], f0=1e-6)
            h, _ = self._gemm(plan, h, g[f"{p}.pw1.w"], rows, 4 * H, H, bias=g[f"{p}.pw1.b"], flags=F_GELU)
            x, _ = self._gemm(plan, h, g[f"{p}.pw2.w"], rows, H, 4 * H, bias=g[f"{p}.pw2.b"], scale=g[f"{p}.gamma"], res=x)
        # 5. vocoder.  Every layer is causal with a finite reach, so when the caller only wants the samples from `skip` on
        #    (streaming: the 25 context frames of a window are decoded for their state, not for their audio) each layer only
        #    has to produce the rows the wanted samples can see: walk the receptive field back from the output to get the
        #    first row of every op (exactly the same samples come out; the rows in front are never computed nor read).
        D = c.decoder_dim
        nblk = len(c.upsample_rates)
        dils = VOCODER_DILATIONS
        right = c.trans_conv_trim == "right"
        dec0_begin, tconv_begin, unit_begin, fin_begin = vocoder_tail_starts(rows, c.upsample_rates, skip, c.trans_conv_trim)
        _, xs = self._gemm(plan, x, g["dec0.w"], rows, D, H, taps=7, tap_off=[j - 6 for j in range(7)], bias=g["dec0.b"],
                           snake="decoder.1.block.0", m_begin=dec0_begin)
        for i, r in enumerate(c.upsample_rates):
            cin, cout = D // 2 ** i, D // 2 ** (i + 1)
            p = f"decoder.{i + 1}.block"
            # weights [W[:, :, :r] | W[:, :, r:]]: trimmed on both sides, output block m = x[m+1] W_lo + x[m] W_hi (m < rows - 1);
            # trimmed on the right only (causal), block m = x[m] W_lo + x[m-1] W_hi with x[-1] = 0 (m < rows)
            m_out = rows if right else rows - 1
            y, ys = self._gemm(plan, xs, g[f"{p}.1.w"], m_out, r * cout, cin, taps=2, tap_off=[0, -1] if right else [1, 0], bias=g[f"{p}.1.b"],
                               col_mod=cout, snake=f"{p}.2.act1", m_begin=tconv_begin[i])
            rows = m_out * r
            x, xs = y.view(-1, cout)[:max(rows, 1)], ys.view(-1, cout)[:max(rows, 1)]
            for j, dil in enumerate(dils):
                q_ = f"{p}.{j + 2}"
                _, hs = self._gemm(plan, xs, g[f"{q_}.c1.w"], rows, cout, cout, taps=7, tap_off=[(t - 6) * dil for t in range(7)],
                                   bias=g[f"{q_}.c1.b"], snake=f"{q_}.act2", m_begin=unit_begin[i][j])
                nxt = f"{p}.{j + 3}.act1" if j < 2 else (f"decoder.{i + 2}.block.0" if i + 1 < nblk else "final")
                x, xs = self._gemm(plan, hs, g[f"{q_}.c2.w"], rows, cout, cout, bias=g[f"{q_}.c2.b"], res=x, snake=nxt,
                                   m_begin=unit_begin[i][j])
        out_dim = D // 2 ** nblk
        plan.wav = torch.zeros(max(rows, 1), 1, dtype=torch.float32, device=self.device)
        self._gemm(plan, xs, g["final.w"], rows, 1, out_dim, taps=7, tap_off=[j - 6 for j in range(7)], bias=g["final.b"],
                   flags=F_CLAMP, out=plan.wav, out_f32=True, m_begin=fin_begin)
        plan.n_samples = rows
        plan.ops = fuse_row_norms(plan.ops)
        if os.environ.get("FQ3C_SPLITK", "1") != "0":
            attach_splitk_workspace(plan.ops, self.device, plan.keep)
        plan.arr = (Op * len(plan.ops))(*plan.ops)
        return plan

    def run_plan(self, plan: _Plan):
        """Launch a plan's op list on the current stream (also the hook the per-op parity tests use)."""
        if plan.arr is None:
            plan.arr = (Op * len(plan.ops))(*plan.ops)
        stream = torch.cuda.current_stream().cuda_stream
        if self.use_graphs and plan.runs >= 1 and getattr(plan, "graph", None) is None and not plan.graph_failed:
            # second use of a cached plan: its launches are static, replay them as one CUDA graph from now on
            h = C.c_void_p()
            if self.lib.fq3c_graph_create(plan.arr, len(plan.ops), C.byref(h)) == 0:
                plan.graph = h
            else:
                plan.graph_failed = True
        plan.runs += 1
        if getattr(plan, "graph", None) is not None:
            rc = self.lib.fq3c_graph_launch(plan.graph, stream)
        else:
            rc = self.lib.fq3c_run(plan.arr, len(plan.ops), stream)
        if rc != 0:
            raise CodecError(self.lib.fq3c_last_error().decode())

    def n_samples(self, T: int) -> int:
        return self.cfg.n_samples(T)

    # ---- run ------------------------------------------------------------------------------------
    @torch.inference_mode()
    def decode(self, codes: torch.Tensor, skip_samples: int = 0) -> torch.Tensor:
        """codes int64 [T, Q] (any device) -> float32 waveform [n_samples] on the decoder's device.  With `skip_samples` = k
        the caller promises to ignore samples [0, k) (the context part of a streaming window): they come back as zeros and
        the vocoder only computes what samples [k, n) can see — those are bit-identical to a full decode."""
        T = int(codes.shape[0])
        if T == 0 or self.n_samples(T) <= 0:
            return torch.zeros(0, dtype=torch.float32, device=self.device)
        skip = int(skip_samples) if (skip_samples and os.environ.get("FQ3C_TAIL_ONLY", "1") != "0") else 0
        key = T if skip == 0 else (T, skip)
        plan = self._plans.get(key)
        if plan is None:
            # streaming revisits a handful of sizes (chunk multiples, then chunk+25; with an ICL voice also reference + 8 / 16 / 24
            # frames, the same for every request of that voice).  Plans are kept least-recently-used: at most MAX_LONG_PLANS of more
            # than 96 frames (a long plan pins ~1 MB of activations per frame), 24 in all.
            def frames(k):
                return k[0] if isinstance(k, tuple) else k
            while True:
                long_keys = [k for k in self._plans if frames(k) > 96]
                if len(long_keys) >= self.MAX_LONG_PLANS and T > 96:
                    victim = long_keys[0]
                elif len(self._plans) >= 24:
                    victim = next(iter(self._plans))
                else:
                    break
                old = self._plans.pop(victim)
                if old.graph is not None:
                    self.lib.fq3c_graph_destroy(old.graph)
            plan = self._plans[key] = self._build(T, skip)
        else:
            self._plans[key] = self._plans.pop(key)  # most recently used goes last
        plan.codes.copy_(codes.to(torch.int64), non_blocking=True)
        self.run_plan(plan)
        return plan.wav.view(-1)[: plan.n_samples].clone()

    def open_stream(self, max_chunk_frames: int = 8, split_k: Optional[bool] = None) -> "CodecStream":
        """Stateful incremental decode of one utterance (CodecStream): every chunk costs its own frames only."""
        return CodecStream(self, max_chunk_frames, split_k)

    @property
    def launch_count(self) -> int:
        return int(self.lib.fq3c_launch_count())


class CodecStream:
    """Stateful incremental decode of one utterance (replaces the re-decoding policy of
    faster_qwen3_tts/model.py:737-826: 25 context frames with every chunk of 8).

    With the causal transposed convolutions (config trans_conv_trim = "right") every layer of the decoder looks backwards only,
    so a chunk needs nothing but the rows earlier chunks left behind: the last k-1 (dilated: 6*d) input rows of every causal
    convolution, one row per vocoder transposed conv, and the rotated K / V rows of the last window-1 positions of every
    transformer layer.  Each such layer input is a buffer [history rows | new rows]: the implicit GEMM reads its taps at
    (offset + history) >= 0, so rows of earlier chunks take the place of the causal zero padding, and the last ops of a
    chunk's list roll every buffer's tail into its history rows.  A chunk of n frames costs n frames and returns exactly
    n * total_upsample samples; the concatenation over chunks equals the full, non-streaming decode of all frames bit for
    bit when neither side splits its GEMMs over K (FQ3C_SPLITK=0, or split_k=False here), because every output element then
    accumulates the same products in the same order; with split-K (the default, faster) the two agree up to fp32 summation order.
    """

    def __init__(self, dec: "CodecDecoder", max_chunk_frames: int = 8, split_k: Optional[bool] = None):
        if dec.cfg.trans_conv_trim != "right":
            raise ValueError('stateful decode needs causal transposed convolutions (trans_conv_trim = "right"): the sibling trim '
                             '("both") makes every vocoder block look one input row ahead')
        self.dec, self.cfg, self.device = dec, dec.cfg, dec.device
        self.max_frames = int(max_chunk_frames)
        # short GEMMs with a long reduction may be split over K like the full decode's (faster: 8 rows cover few tiles); the chunks
        # then equal the full decode up to fp32 summation order instead of bit for bit.  FQ3C_SPLITK=0 / split_k=False: exact.
        self.split_k = (os.environ.get("FQ3C_SPLITK", "1") != "0") if split_k is None else bool(split_k)
        self.hist: List[tuple] = []          # (buffer [hist + rows_max, cols], hist rows, name)
        self.rows_of: Dict[str, int] = {}    # new rows per frame of every history buffer
        self._plans: Dict[int, _Plan] = {}
        self.pos = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.codes = torch.zeros(self.max_frames, self.cfg.num_quantizers, dtype=torch.int64, device=self.device)
        self.n_frames = 0
        c = self.cfg
        H, W = c.hidden_size, c.sliding_window
        qkv_dim = (c.num_attention_heads + 2 * c.num_key_value_heads) * c.head_dim
        self.k_pre = dec.g["pre_conv.w"].shape[1] // c.latent_dim
        self.b: Dict[str, torch.Tensor] = {}
        self._mk("pre_in", self.k_pre - 1, 1, c.latent_dim)
        for l in range(c.num_hidden_layers):
            self._mk(f"qkv{l}", W - 1, 1, qkv_dim)
        per = 1
        for i, f in enumerate(c.upsampling_ratios):
            per *= f
            self._mk(f"cnx{i}", 6, per, H)
        self._mk("dec0_in", 6, per, H)
        D = c.decoder_dim
        for i, r in enumerate(c.upsample_rates):
            cin, cout = D // 2 ** i, D // 2 ** (i + 1)
            self._mk(f"tc{i}", 1, per, cin)
            per *= r
            for j, dil in enumerate(VOCODER_DILATIONS):
                self._mk(f"u{i}.{j}", 6 * dil, per, cout)
        self._mk("fin_in", 6, per, D // 2 ** len(c.upsample_rates))
        self.tmp = torch.zeros(max(h * t.shape[1] for t, h, _ in self.hist), dtype=torch.bfloat16, device=self.device)

    def _mk(self, name, hist, rows_per_frame, cols):
        t = torch.zeros(hist + rows_per_frame * self.max_frames, cols, dtype=torch.bfloat16, device=self.device)
        self.b[name] = t
        self.hist.append((t, hist, name))
        self.rows_of[name] = rows_per_frame

    def reset(self):
        """Start a new utterance: history rows back to the causal zero padding, position 0."""
        for t, h, _ in self.hist:
            t[:h].zero_()
        self.pos.zero_()
        self.n_frames = 0

    def _build(self, T: int) -> _Plan:
        d_, c, g, b = self.dec, self.cfg, self.dec.g, self.b
        plan = _Plan()
        H, d, nh, nkv, W = c.hidden_size, c.head_dim, c.num_attention_heads, c.num_key_value_heads, c.sliding_window
        plan.codes = self.codes
        q = d_._simple(plan, K_RVQ, self.codes[:T], T, 2 * c.codebook_dim, B=g["cb"], K=c.codebook_size,
                       i0=c.num_semantic_quantizers, i1=c.num_quantizers, i2=c.codebook_dim)
        plan.ops[-1].lda = c.num_quantizers
        k = self.k_pre
        d_._gemm(plan, q, g["rvq_proj"], T, c.latent_dim, 2 * c.codebook_dim, out=b["pre_in"][k - 1:k - 1 + T])
        x, _ = d_._gemm(plan, b["pre_in"][:k - 1 + T], g["pre_conv.w"], T, H, c.latent_dim, taps=k, tap_off=list(range(k)),
                        bias=g["pre_conv.b"])
        qkv_dim = (nh + 2 * nkv) * d
        for l in range(c.num_hidden_layers):
            p = f"pre_transformer.layers.{l}"
            buf = b[f"qkv{l}"]
            new = buf[W - 1:W - 1 + T]
            h = d_._simple(plan, K_RMSNORM, x, T, H, scale=g[f"{p}.ln1"], f0=c.rms_norm_eps)
            d_._gemm(plan, h, g[f"{p}.wqkv"], T, qkv_dim, H, out=new)
            d_._simple(plan, K_ROPE, new, T, qkv_dim, out=new, i0=nh + nkv, i1=d, i2=0, f0=c.rope_theta, p0=self.pos)
            a = d_._simple(plan, K_ATTN, buf[:W - 1 + T], T, nh * d, i0=nh, i1=nkv, i2=d, K=W, taps=W - 1, p0=self.pos)
            x, _ = d_._gemm(plan, a, g[f"{p}.wo"], T, H, nh * d, scale=g[f"{p}.ls1"], res=x)
            h = d_._simple(plan, K_RMSNORM, x, T, H, scale=g[f"{p}.ln2"], f0=c.rms_norm_eps)
            m, _ = d_._gemm(plan, h, g[f"{p}.wgu"], T, 2 * c.intermediate_size, H, flags=F_SWIGLU)
            x, _ = d_._gemm(plan, m, g[f"{p}.wdown"], T, H, c.intermediate_size, scale=g[f"{p}.ls2"], res=x)
        x = d_._simple(plan, K_RMSNORM, x, T, H, scale=g["norm"], f0=c.rms_norm_eps)
        rows = T
        n_up = len(c.upsampling_ratios)
        for i, f in enumerate(c.upsampling_ratios):
            p = f"upsample.{i}"
            cb = b[f"cnx{i}"]
            xin = cb[6:6 + rows * f]
            d_._gemm(plan, x, g[f"{p}.t.w"], rows, f * H, H, bias=g[f"{p}.t.b"], col_mod=H, out=xin.view(rows, f * H))
            rows *= f
            h = d_._simple(plan, K_DWCONV, xin, rows, H, B=g[f"{p}.dw.w"], bias=g[f"{p}.dw.b"], taps=7, i0=6)
            h = d_._simple(plan, K_LAYERNORM, h, rows, H, scale=g[f"{p}.ln.w"], bias=g[f"{p}.ln.b"], f0=1e-6)
            h, _ = d_._gemm(plan, h, g[f"{p}.pw1.w"], rows, 4 * H, H, bias=g[f"{p}.pw1.b"], flags=F_GELU)
            out = b["dec0_in"][6:6 + rows] if i == n_up - 1 else None
            x, _ = d_._gemm(plan, h, g[f"{p}.pw2.w"], rows, H, 4 * H, bias=g[f"{p}.pw2.b"], scale=g[f"{p}.gamma"], res=xin, out=out)
        if n_up == 0:
            d_._simple(plan, K_COPY, x, rows, H, out=b["dec0_in"][6:6 + rows])
        D = c.decoder_dim
        nblk = len(c.upsample_rates)
        d_._gemm(plan, b["dec0_in"][:6 + rows], g["dec0.w"], rows, D, H, taps=7, tap_off=list(range(7)), bias=g["dec0.b"],
                 snake="decoder.1.block.0", out2=b["tc0"][1:1 + rows])
        for i, r in enumerate(c.upsample_rates):
            cin, cout = D // 2 ** i, D // 2 ** (i + 1)
            p = f"decoder.{i + 1}.block"
            u0 = b[f"u{i}.0"]
            y, _ = d_._gemm(plan, b[f"tc{i}"][:1 + rows], g[f"{p}.1.w"], rows, r * cout, cin, taps=2, tap_off=[1, 0], bias=g[f"{p}.1.b"],
                            col_mod=cout, snake=f"{p}.2.act1", out2=u0[6:6 + rows * r].view(rows, r * cout))
            rows *= r
            x = y.view(rows, cout)
            for j, dil in enumerate(VOCODER_DILATIONS):
                q_ = f"{p}.{j + 2}"
                uin = b[f"u{i}.{j}"]
                _, hs = d_._gemm(plan, uin[:6 * dil + rows], g[f"{q_}.c1.w"], rows, cout, cout, taps=7, tap_off=[t * dil for t in range(7)],
                                 bias=g[f"{q_}.c1.b"], snake=f"{q_}.act2")
                if j < 2:
                    nxt, dn = f"{p}.{j + 3}.act1", 6 * VOCODER_DILATIONS[j + 1]
                    o2 = b[f"u{i}.{j + 1}"][dn:dn + rows]
                elif i + 1 < nblk:
                    nxt, o2 = f"decoder.{i + 2}.block.0", b[f"tc{i + 1}"][1:1 + rows]
                else:
                    nxt, o2 = "final", b["fin_in"][6:6 + rows]
                x, _ = d_._gemm(plan, hs, g[f"{q_}.c2.w"], rows, cout, cout, bias=g[f"{q_}.c2.b"], res=x, snake=nxt, out2=o2)
        out_dim = D // 2 ** nblk
        plan.wav = torch.zeros(max(rows, 1), 1, dtype=torch.float32, device=self.device)
        d_._gemm(plan, b["fin_in"][:6 + rows], g["final.w"], rows, 1, out_dim, taps=7, tap_off=list(range(7)), bias=g["final.b"],
                 flags=F_CLAMP, out=plan.wav, out_f32=True)
        plan.n_samples = rows
        # roll: the last `hist` rows of [history | new] become the history of the next chunk (through a scratch buffer when
        # the two ranges overlap), then the position counter moves on
        if os.environ.get("FQ3C_ROLL", "1") != "0":
            # one launch for all of it (FQ3C_ROLL): 37 copies + the counter were 38 of a chunk's 129 launches
            rows = []
            for t, hist, name in self.hist:
                assert t.is_contiguous() and t.shape[1] % 8 == 0
                rows.append([t.data_ptr(), hist, self.rows_of[name] * T, t.shape[1], t.shape[1]])
            desc = torch.tensor(rows, dtype=torch.int64, device=self.device)
            plan.keep.append(desc)
            roll = Op()
            roll.kind, roll.M, roll.N, roll.i0 = K_ROLL, len(rows), 1, T
            roll.A, roll.C = desc.data_ptr(), self.pos.data_ptr()
            plan.ops.append(roll)
        else:
            for t, hist, name in self.hist:
                n_new = self.rows_of[name] * T
                cols = t.shape[1]
                src = t[n_new:n_new + hist]
                if n_new >= hist:
                    d_._simple(plan, K_COPY, src, hist, cols, out=t[:hist])
                else:
                    tmp = self.tmp[:hist * cols].view(hist, cols)
                    d_._simple(plan, K_COPY, src, hist, cols, out=tmp)
                    d_._simple(plan, K_COPY, tmp, hist, cols, out=t[:hist])
            adv = Op()
            adv.kind, adv.M, adv.N, adv.i0, adv.C = K_ADVANCE, 1, 1, T, self.pos.data_ptr()
            plan.ops.append(adv)
        plan.ops = fuse_row_norms(plan.ops)
        if self.split_k:
            attach_splitk_workspace(plan.ops, self.device, plan.keep)
        plan.arr = (Op * len(plan.ops))(*plan.ops)
        return plan

    @torch.inference_mode()
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes int64 [n, Q], the next n <= max_chunk_frames frames of the utterance -> float32 [n * total_upsample] on the device."""
        T = int(codes.shape[0])
        if T == 0:
            return torch.zeros(0, dtype=torch.float32, device=self.device)
        if T > self.max_frames:
            return torch.cat([self.decode(codes[i:i + self.max_frames]) for i in range(0, T, self.max_frames)])
        plan = self._plans.get(T)
        if plan is None:
            plan = self._plans[T] = self._build(T)
        self.codes[:T].copy_(codes.to(torch.int64), non_blocking=True)
        self.dec.run_plan(plan)
        self.n_frames += T
        return plan.wav.view(-1)[: plan.n_samples].clone()

    def close(self):
        for plan in self._plans.values():
            if plan.graph is not None:
                self.dec.lib.fq3c_graph_destroy(plan.graph)
                plan.graph = None
        self._plans.clear()


class SpeechTokenizer:
    """`model.speech_tokenizer` surface the reference uses: `.decode({"audio_codes": [B,T,Q]}) -> ([wav], sr)` (model.py:642), the
    list form `.decode([{"audio_codes": [T,Q]}, ...])` of its examples, and `.sample_rate` (model.py:56-58)."""

    def __init__(self, decoder: CodecDecoder):
        self.decoder = decoder
        self.sample_rate = decoder.cfg.sample_rate

    @classmethod
    def synthetic(cls, cfg: CodecDecoderConfig, device, seed: int = 1):
        return cls(CodecDecoder(cfg, init_codec_synthetic(cfg, seed=seed), device))

    def decode(self, inputs) -> Tuple[List[torch.Tensor], int]:
        if isinstance(inputs, (list, tuple)):  # [{"audio_codes": [T, Q]}, ...]: the form of the reference's examples/generate_with_embedding.py:98
            out: List[torch.Tensor] = []
            for item in inputs:
                out.extend(self.decode(item)[0])
            return out, self.sample_rate
        codes = inputs["audio_codes"] if isinstance(inputs, dict) else inputs
        skip = int(inputs.get("skip_samples", 0)) if isinstance(inputs, dict) else 0  # extension: see CodecDecoder.decode
        if codes.dim() == 2:
            codes = codes.unsqueeze(0)
        return [self.decoder.decode(codes[b], skip) for b in range(codes.shape[0])], self.sample_rate
