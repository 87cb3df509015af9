"""Shared builders for the parity tests (oracle + engine from the same seeded weights)."""
from __future__ import annotations

import torch

from qwen3_tts_cuda_graphs_b200.config import preset, with_layers
from qwen3_tts_cuda_graphs_b200.weights import init_synthetic, pack_arena


def make_cfg(name="tiny", talker_layers=None, predictor_layers=None):
    cfg = preset(name)
    if talker_layers is not None:
        cfg = with_layers(cfg, talker_layers, predictor_layers)
    return cfg


def make_weights(cfg, seed=0, head_scale=1.0, norm_jitter=0.1):
    return init_synthetic(cfg, seed=seed, head_scale=head_scale, norm_jitter=norm_jitter, skip_text_embedding=True)


def make_engine(cfg, w, max_seq_len=256, max_streams=1, max_frames=512):
    from qwen3_tts_cuda_graphs_b200.engine import Engine

    arena = pack_arena(cfg, w, max_seq_len, torch.device("cuda"))
    return Engine(cfg, arena, max_seq_len=max_seq_len, max_streams=max_streams, max_frames=max_frames)


def make_oracle(cfg, w, attn="eager", device="cpu"):
    from oracle.qwen3_tts_oracle import OracleTTS

    return OracleTTS(cfg, w, attn=attn, device=device)


def synth_prompt(cfg, T=14, R=1, seed=1, scale=0.05):
    """Random prompt embeddings of the shapes _build_talker_inputs_local returns (model.py:553)."""
    g = torch.Generator().manual_seed(seed)
    H = cfg.talker.hidden_size
    tie = (scale * torch.randn(1, T, H, generator=g)).to(torch.bfloat16)
    tam = torch.ones(1, T, dtype=torch.long)
    tth = (scale * torch.randn(1, R, H, generator=g)).to(torch.bfloat16)
    tpe = (scale * torch.randn(1, 1, H, generator=g)).to(torch.bfloat16)
    return tie, tam, tth, tpe


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu().reshape(-1), b.float().cpu().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def margin_argmax_agree(a: torch.Tensor, b: torch.Tensor, tol: float):
    """argmax(a) == argmax(b) unless b's top-2 margin is below tol (bf16 near-tie, SURVEY.md §7 hard parts)."""
    a, b = a.float().cpu().reshape(-1), b.float().cpu().reshape(-1)
    top2 = torch.topk(b, 2).values
    if float(top2[0] - top2[1]) <= tol:
        return True
    return int(a.argmax()) == int(b.argmax())
