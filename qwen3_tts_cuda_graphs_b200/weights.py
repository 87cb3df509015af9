"""Synthetic (random-init) weights of the named Qwen3-TTS architectures and the arena packer.

There is no network and no checkpoint in this environment, so BASELINE.json's configs are run on
random-init weights of the named architecture (seeded, N(0, 0.02^2), RMSNorm weight 1 — SURVEY.md §8d).
Tensor names follow the upstream `qwen_tts` module tree that the reference dereferences
(`predictor_graph.py:52-57`, `generate.py:99-102`, `model.py:353,395-403`), so a real state dict with the same
keys can be packed by the same code.

`pack_arena` lays every matrix the kernels touch into ONE contiguous device buffer in streaming order
(fused [q;k;v] rows, interleaved [gate_j; up_j] rows, 256-byte aligned) — the layout `include/fq3.h`
documents for `fq3_stack_desc.layer_offs`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .config import StackConfig, TTSConfig

_ALIGN = 256


def _stack_names(prefix: str, cfg: StackConfig) -> List[tuple]:
    h, i, d = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    out = []
    for l in range(cfg.num_hidden_layers):
        p = f"{prefix}.layers.{l}"
        out += [
            (f"{p}.input_layernorm.weight", (h,), "norm"),
            (f"{p}.self_attn.q_proj.weight", (cfg.q_dim, h), "lin"),
            (f"{p}.self_attn.k_proj.weight", (cfg.kv_dim, h), "lin"),
            (f"{p}.self_attn.v_proj.weight", (cfg.kv_dim, h), "lin"),
            (f"{p}.self_attn.q_norm.weight", (d,), "norm"),
            (f"{p}.self_attn.k_norm.weight", (d,), "norm"),
            (f"{p}.self_attn.o_proj.weight", (h, cfg.q_dim), "lin"),
            (f"{p}.post_attention_layernorm.weight", (h,), "norm"),
            (f"{p}.mlp.gate_proj.weight", (i, h), "lin"),
            (f"{p}.mlp.up_proj.weight", (i, h), "lin"),
            (f"{p}.mlp.down_proj.weight", (h, i), "lin"),
        ]
    out.append((f"{prefix}.norm.weight", (h,), "norm"))
    return out


def tensor_specs(cfg: TTSConfig) -> List[tuple]:
    """(name, shape, kind) of every LM-side tensor; kind in {lin, emb, norm, bias, head}."""
    t, p = cfg.talker, cfg.predictor
    ncb = p.num_codebooks
    specs = [
        ("talker.model.codec_embedding.weight", (t.vocab_size, t.hidden_size), "emb"),
        ("talker.model.text_embedding.weight", (t.text_vocab_size, t.text_hidden_size), "emb"),
        ("talker.text_projection.linear_fc1.weight", (t.text_hidden_size, t.text_hidden_size), "lin"),
        ("talker.text_projection.linear_fc1.bias", (t.text_hidden_size,), "bias"),
        ("talker.text_projection.linear_fc2.weight", (t.hidden_size, t.text_hidden_size), "lin"),
        ("talker.text_projection.linear_fc2.bias", (t.hidden_size,), "bias"),
    ]
    specs += _stack_names("talker.model", t)
    specs.append(("talker.codec_head.weight", (t.vocab_size, t.hidden_size), "head"))
    if t.hidden_size != p.hidden_size:
        specs += [
            ("talker.code_predictor.small_to_mtp_projection.weight", (p.hidden_size, t.hidden_size), "lin"),
            ("talker.code_predictor.small_to_mtp_projection.bias", (p.hidden_size,), "bias"),
        ]
    specs += _stack_names("talker.code_predictor.model", p)
    for i in range(ncb):
        specs.append((f"talker.code_predictor.model.codec_embedding.{i}.weight", (p.vocab_size, t.hidden_size), "emb"))
    for i in range(ncb):
        specs.append((f"talker.code_predictor.lm_head.{i}.weight", (p.vocab_size, p.hidden_size), "head"))
    return specs


def init_synthetic(
    cfg: TTSConfig,
    seed: int = 0,
    std: float = 0.02,
    norm_jitter: float = 0.0,
    head_scale: float = 1.0,
    dtype: torch.dtype = torch.bfloat16,
    skip_text_embedding: bool = False,
) -> Dict[str, torch.Tensor]:
    """Seeded random init on the CPU (the same tensors feed the oracle and the CUDA engine).

    head_scale > 1 amplifies codec_head / lm_heads so greedy margins are far above one bf16 ulp — the
    "amplified-head" variant SURVEY.md §7 prescribes for free-running token-exactness tests.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in tensor_specs(cfg):
        if skip_text_embedding and name == "talker.model.text_embedding.weight":
            continue
        if kind == "norm":
            w = torch.ones(shape, dtype=torch.float32)
            if norm_jitter:
                w = w + norm_jitter * torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "bias":
            w = std * torch.randn(shape, generator=g, dtype=torch.float32)
        else:
            w = torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)
            if kind == "head":
                w *= head_scale
        out[name] = w.to(dtype)
    return out


def rope_tables(head_dim: int, theta: float, length: int, dtype=torch.bfloat16):
    """cos/sin exactly as HF Qwen3RotaryEmbedding computes them (fp32 outer product, cat(freqs, freqs),
    cast to the activation dtype).  With all three mrope axes equal (talker_graph.py:209-211) the
    interleaved multimodal rope reduces to this 1-D table."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(length, dtype=torch.float32)
    freqs = torch.outer(pos, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


@dataclass
class Arena:
    """Packed device weights + the byte offsets the C ABI consumes."""

    buf: torch.Tensor  # uint8 [bytes] on the device
    offsets: Dict[str, int] = field(default_factory=dict)
    shapes: Dict[str, tuple] = field(default_factory=dict)

    def view(self, name: str) -> torch.Tensor:
        off, shape = self.offsets[name], self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return self.buf[off : off + 2 * n].view(torch.bfloat16).view(shape)


def _packed_stack(prefix: str, cfg: StackConfig, w: Dict[str, torch.Tensor]):
    """Yield (arena_name, tensor) in streaming order for one stack."""
    for l in range(cfg.num_hidden_layers):
        p = f"{prefix}.layers.{l}"
        yield f"{p}.ln1", w[f"{p}.input_layernorm.weight"]
        yield f"{p}.wqkv", torch.cat(
            [w[f"{p}.self_attn.q_proj.weight"], w[f"{p}.self_attn.k_proj.weight"], w[f"{p}.self_attn.v_proj.weight"]], 0
        )
        yield f"{p}.qn", w[f"{p}.self_attn.q_norm.weight"]
        yield f"{p}.kn", w[f"{p}.self_attn.k_norm.weight"]
        yield f"{p}.wo", w[f"{p}.self_attn.o_proj.weight"]
        yield f"{p}.ln2", w[f"{p}.post_attention_layernorm.weight"]
        gate, up = w[f"{p}.mlp.gate_proj.weight"], w[f"{p}.mlp.up_proj.weight"]
        yield f"{p}.wgu", torch.stack([gate, up], dim=1).reshape(2 * gate.shape[0], gate.shape[1])
        yield f"{p}.wdown", w[f"{p}.mlp.down_proj.weight"]
    yield f"{prefix}.norm", w[f"{prefix}.norm.weight"]


LAYER_FIELDS = ("ln1", "wqkv", "qn", "kn", "wo", "ln2", "wgu", "wdown")


def pack_arena(cfg: TTSConfig, w: Dict[str, torch.Tensor], max_seq_len: int, device) -> Arena:
    """Copy the weights into one device buffer in the order the persistent kernel streams them."""
    t, p = cfg.talker, cfg.predictor
    items = []
    # predictor first, then talker: the order one frame touches them
    items += list(_packed_stack("talker.code_predictor.model", p, w))
    for i in range(p.num_codebooks):
        items.append((f"talker.code_predictor.lm_head.{i}", w[f"talker.code_predictor.lm_head.{i}.weight"]))
    if t.hidden_size != p.hidden_size:
        items.append(("talker.code_predictor.s2m.weight", w["talker.code_predictor.small_to_mtp_projection.weight"]))
        items.append(("talker.code_predictor.s2m.bias", w["talker.code_predictor.small_to_mtp_projection.bias"]))
    items += list(_packed_stack("talker.model", t, w))
    items.append(("talker.codec_head", w["talker.codec_head.weight"]))
    items.append(("talker.codec_embedding", w["talker.model.codec_embedding.weight"]))
    for i in range(p.num_codebooks):
        items.append(
            (f"talker.code_predictor.codec_embedding.{i}", w[f"talker.code_predictor.model.codec_embedding.{i}.weight"])
        )
    for name in (
        "talker.text_projection.linear_fc1.weight", "talker.text_projection.linear_fc1.bias",
        "talker.text_projection.linear_fc2.weight", "talker.text_projection.linear_fc2.bias",
        "talker.model.text_embedding.weight",
    ):
        if name in w:
            items.append((name, w[name]))
    cos_t, sin_t = rope_tables(t.head_dim, t.rope_theta, max_seq_len + 8)
    cos_p, sin_p = rope_tables(p.head_dim, p.rope_theta, p.num_code_groups + 2)
    items += [("rope.talker.cos", cos_t), ("rope.talker.sin", sin_t), ("rope.pred.cos", cos_p), ("rope.pred.sin", sin_p)]

    total = 0
    offs, shapes = {}, {}
    for name, ten in items:
        assert ten.dtype == torch.bfloat16, name
        offs[name] = total
        shapes[name] = tuple(ten.shape)
        total += (ten.numel() * 2 + _ALIGN - 1) // _ALIGN * _ALIGN
    buf = torch.empty(total, dtype=torch.uint8, device=device)
    for name, ten in items:
        n = ten.numel() * 2
        buf[offs[name] : offs[name] + n].view(torch.bfloat16).copy_(ten.contiguous().view(-1), non_blocking=False)
    return Arena(buf=buf, offsets=offs, shapes=shapes)


def param_bytes(cfg: TTSConfig) -> Dict[str, int]:
    """Algorithmic weight bytes per frame for the roofline line."""
    t, p = cfg.talker, cfg.predictor
    talker_layers = t.num_hidden_layers * t.layer_params() + t.hidden_size
    head = t.vocab_size * t.hidden_size
    pred_pass = p.num_hidden_layers * p.layer_params() + p.hidden_size
    s2m = (p.hidden_size * t.hidden_size + p.hidden_size) if t.hidden_size != p.hidden_size else 0
    lm_heads = p.num_codebooks * p.vocab_size * p.hidden_size
    return {
        "talker_step": 2 * (talker_layers + head),
        "predictor_pass": 2 * (pred_pass + s2m),
        "predictor_heads": 2 * lm_heads,
        "frame_streaming": 2 * (talker_layers + head + p.num_codebooks * (pred_pass + s2m) + lm_heads),
        "frame_read_once": 2 * (talker_layers + head + pred_pass + s2m + lm_heads),
    }
