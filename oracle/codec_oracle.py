"""CPU/PyTorch restatement of the 12 Hz speech-tokenizer DECODER — TEST INFRASTRUCTURE ONLY.

The decoder lives in the un-vendored `qwen-tts` package (`speech_tokenizer.decode`, called at
faster_qwen3_tts/model.py:642,782,811,884,971,988,1054,1136,1153); it is restated here from the in-container
sibling `transformers/models/qwen3_omni_moe/modeling_qwen3_omni_moe.py:3283-3790` (Code2Wav: causal convs,
ConvNeXt upsampling, sliding-window transformer with layer scale, SnakeBeta vocoder blocks) with a split-RVQ
dequantiser (1 semantic + 15 acoustic codebooks, shared output projections) in place of the sibling's
summed embedding, as SURVEY.md §8(c) records.  Pinned against the sibling modules run live
(tests/test_codec_cpu.py); "parity unpinned" against upstream qwen_tts itself.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def causal_conv1d(x, w, b, dilation=1):
    """Qwen3OmniMoeCausalConvNet (stride 1): left pad (k-1)*dilation.  x [B,C,T]."""
    k = w.shape[-1]
    pad = (k - 1) * dilation
    return F.conv1d(F.pad(x, (pad, 0)), w, b, dilation=dilation)


def causal_trans_conv1d(x, w, b, stride, trim="both"):
    """ConvTranspose1d, then trim (k - stride).  "both": on BOTH sides, as Qwen3OmniMoeCausalTransConvNet does (the sibling this
    file is pinned to); "right": at the end only — the causal variant whose length law (stride * T, no lookahead) is the one the
    reference's own sample WAVs show (every file is k x 1920 samples)."""
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    pad = k - stride
    if pad <= 0:
        return y
    return y[..., : y.shape[-1] - pad] if trim == "right" else y[..., pad: y.shape[-1] - pad]


def snake_beta(x, alpha, beta):
    a = torch.exp(alpha).view(1, -1, 1)
    bb = torch.exp(beta).view(1, -1, 1)
    return x + (1.0 / (bb + 1e-9)) * torch.pow(torch.sin(x * a), 2)


def rms_norm(x, w, eps):
    dt = x.dtype
    h = x.float()
    h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps)
    return w * h.to(dt)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class CodecOracle:
    def __init__(self, cfg, weights: Dict[str, torch.Tensor], device="cpu"):
        self.cfg = cfg
        self.w = {k: v.to(device) for k, v in weights.items()}
        self.device = torch.device(device)

    # ---- stages ---------------------------------------------------------------------------------
    def dequant(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [T, Q] -> [1, latent, T]"""
        c, w = self.cfg, self.w
        ns = c.num_semantic_quantizers
        first = sum(F.embedding(codes[:, g], w[f"quantizer.codebook.{g}"]) for g in range(ns))
        rest = sum(F.embedding(codes[:, g], w[f"quantizer.codebook.{g}"]) for g in range(ns, c.num_quantizers))
        h = F.linear(first, w["quantizer.rvq_first.output_proj.weight"]) + F.linear(rest, w["quantizer.rvq_rest.output_proj.weight"])
        return h.t().unsqueeze(0)

    def transformer(self, x: torch.Tensor) -> torch.Tensor:
        """x [1, T, H] -> [1, T, H]; sliding-window causal attention, layer scale, final RMSNorm."""
        c, w = self.cfg, self.w
        T = x.shape[1]
        d, nh, nkv = c.head_dim, c.num_attention_heads, c.num_key_value_heads
        inv = 1.0 / (c.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
        fr = torch.outer(torch.arange(T, dtype=torch.float32), inv)
        emb = torch.cat((fr, fr), -1)
        cos, sin = emb.cos().to(x.dtype).to(x.device), emb.sin().to(x.dtype).to(x.device)
        qi = torch.arange(T).view(T, 1)
        ki = torch.arange(T).view(1, T)
        allowed = (ki <= qi) & (ki > qi - c.sliding_window)
        mask = torch.zeros(T, T, dtype=x.dtype).masked_fill(~allowed, torch.finfo(x.dtype).min).to(x.device)
        for l in range(c.num_hidden_layers):
            p = f"pre_transformer.layers.{l}"
            h = rms_norm(x, w[f"{p}.input_layernorm.weight"], c.rms_norm_eps)
            q = F.linear(h, w[f"{p}.self_attn.q_proj.weight"]).view(1, T, nh, d).transpose(1, 2)
            k = F.linear(h, w[f"{p}.self_attn.k_proj.weight"]).view(1, T, nkv, d).transpose(1, 2)
            v = F.linear(h, w[f"{p}.self_attn.v_proj.weight"]).view(1, T, nkv, d).transpose(1, 2)
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
            g = nh // nkv
            k, v = k.repeat_interleave(g, 1), v.repeat_interleave(g, 1)
            s = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5) + mask
            a = torch.matmul(F.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype), v)
            a = a.transpose(1, 2).reshape(1, T, nh * d)
            x = x + w[f"{p}.self_attn_layer_scale.scale"] * F.linear(a, w[f"{p}.self_attn.o_proj.weight"])
            h = rms_norm(x, w[f"{p}.post_attention_layernorm.weight"], c.rms_norm_eps)
            m = F.silu(F.linear(h, w[f"{p}.mlp.gate_proj.weight"])) * F.linear(h, w[f"{p}.mlp.up_proj.weight"])
            x = x + w[f"{p}.mlp_layer_scale.scale"] * F.linear(m, w[f"{p}.mlp.down_proj.weight"])
        return rms_norm(x, w["pre_transformer.norm.weight"], c.rms_norm_eps)

    def convnext(self, x, p):
        w = self.w
        h = F.conv1d(F.pad(x, (6, 0)), w[f"{p}.dwconv.conv.weight"], w[f"{p}.dwconv.conv.bias"], groups=x.shape[1])
        h = h.permute(0, 2, 1)
        h = F.layer_norm(h, (h.shape[-1],), w[f"{p}.norm.weight"], w[f"{p}.norm.bias"], 1e-6)
        h = F.linear(h, w[f"{p}.pwconv1.weight"], w[f"{p}.pwconv1.bias"])
        h = F.gelu(h)
        h = F.linear(h, w[f"{p}.pwconv2.weight"], w[f"{p}.pwconv2.bias"])
        h = w[f"{p}.gamma"] * h
        return x + h.permute(0, 2, 1)

    def res_unit(self, x, p, dilation):
        w = self.w
        h = snake_beta(x, w[f"{p}.act1.alpha"], w[f"{p}.act1.beta"])
        h = causal_conv1d(h, w[f"{p}.conv1.conv.weight"], w[f"{p}.conv1.conv.bias"], dilation)
        h = snake_beta(h, w[f"{p}.act2.alpha"], w[f"{p}.act2.beta"])
        h = causal_conv1d(h, w[f"{p}.conv2.conv.weight"], w[f"{p}.conv2.conv.bias"])
        return h + x

    # ---- full decode ----------------------------------------------------------------------------
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """codes int64 [T, Q] -> waveform float [n_samples] in [-1, 1]."""
        c, w = self.cfg, self.w
        x = self.dequant(codes.to(self.device))
        x = causal_conv1d(x, w["pre_conv.conv.weight"], w["pre_conv.conv.bias"])
        x = self.transformer(x.permute(0, 2, 1)).permute(0, 2, 1)
        return self.upsample_and_vocode(x)

    def upsample_and_vocode(self, x: torch.Tensor) -> torch.Tensor:
        """[1, hidden, T] -> waveform; sibling Code2Wav.forward after pre_transformer (:3768-3778)."""
        c, w = self.cfg, self.w
        for i, f in enumerate(c.upsampling_ratios):
            x = causal_trans_conv1d(x, w[f"upsample.{i}.0.conv.weight"], w[f"upsample.{i}.0.conv.bias"], f)
            x = self.convnext(x, f"upsample.{i}.1")
        x = causal_conv1d(x, w["decoder.0.conv.weight"], w["decoder.0.conv.bias"])
        for i, r in enumerate(c.upsample_rates):
            p = f"decoder.{i + 1}.block"
            x = snake_beta(x, w[f"{p}.0.alpha"], w[f"{p}.0.beta"])
            x = causal_trans_conv1d(x, w[f"{p}.1.conv.weight"], w[f"{p}.1.conv.bias"], r, getattr(c, "trans_conv_trim", "both"))
            for j, dil in enumerate((1, 3, 9)):
                x = self.res_unit(x, f"{p}.{j + 2}", dil)
        n = len(c.upsample_rates) + 1
        x = snake_beta(x, w[f"decoder.{n}.alpha"], w[f"decoder.{n}.beta"])
        x = causal_conv1d(x, w[f"decoder.{n + 1}.conv.weight"], w[f"decoder.{n + 1}.conv.bias"])
        return x.clamp(-1, 1).reshape(-1).float()
