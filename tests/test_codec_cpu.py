"""Pins the codec-decoder oracle (test infrastructure) against the sibling transformers modules run live.

The 12 Hz decoder lives in the un-vendored `qwen_tts`; the in-container sibling is transformers'
Qwen3OmniMoeCode2Wav.  The oracle's transformer and its upsample+vocoder stack must reproduce the
sibling module on the sibling's own random weights; the split-RVQ dequantiser has no sibling and is checked against
a direct restatement.
"""
import pytest
import torch

from qwen3_tts_cuda_graphs_b200.codec import codec_tensor_specs, init_codec_synthetic
from qwen3_tts_cuda_graphs_b200.config import preset


def _sibling(cfg):
    mod = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    conf = pytest.importorskip("transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe")
    c = conf.Qwen3OmniMoeCode2WavConfig(
        codebook_size=cfg.codebook_size, hidden_size=cfg.hidden_size, max_position_embeddings=8000,
        rope_theta=cfg.rope_theta, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, attention_bias=False, sliding_window=cfg.sliding_window,
        intermediate_size=cfg.intermediate_size, hidden_act="silu", layer_scale_initial_scale=0.3,
        rms_norm_eps=cfg.rms_norm_eps, num_hidden_layers=cfg.num_hidden_layers, num_quantizers=cfg.num_quantizers,
        upsample_rates=list(cfg.upsample_rates), upsampling_ratios=list(cfg.upsampling_ratios),
        decoder_dim=cfg.decoder_dim, attention_dropout=0.0,
    )
    c._attn_implementation = "eager"
    torch.manual_seed(0)
    m = mod.Qwen3OmniMoeCode2Wav(c).eval()
    # default init leaves SnakeBeta / layer-scale / gamma at constants: randomise so every term is exercised
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.copy_(0.3 * torch.randn_like(p))
            elif n.endswith("gamma") or n.endswith("scale"):
                p.copy_(0.3 + 0.05 * torch.randn_like(p))
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.copy_(0.02 * torch.randn_like(p))
    return m


def _oracle_from_sibling(cfg, m):
    from oracle.codec_oracle import CodecOracle

    sd = m.state_dict()
    w = init_codec_synthetic(cfg, seed=5)  # quantizer / pre_conv entries (no sibling counterpart)
    for name, shape, _ in codec_tensor_specs(cfg):
        if name.startswith("quantizer.") or name.startswith("pre_conv."):
            continue
        assert name in sd, name
        assert tuple(sd[name].shape) == tuple(shape), (name, tuple(sd[name].shape), shape)
        w[name] = sd[name].clone().float()
    return CodecOracle(cfg, w)


@pytest.fixture(scope="module")
def pair():
    from dataclasses import replace
    cfg = replace(preset("tiny").codec, trans_conv_trim="both")  # the sibling trims its transposed convs on both sides
    m = _sibling(cfg)
    return cfg, m, _oracle_from_sibling(cfg, m)


def test_head_dim_matches_sibling(pair):
    cfg, m, _ = pair
    assert m.pre_transformer.layers[0].self_attn.head_dim == cfg.head_dim


@pytest.mark.parametrize("T", [1, 5, 19])
def test_transformer_equals_sibling(pair, T):
    cfg, m, orc = pair
    x = torch.randn(1, T, cfg.hidden_size, generator=torch.Generator().manual_seed(T))
    with torch.no_grad():
        ref = m.pre_transformer(inputs_embeds=x).last_hidden_state
        got = orc.transformer(x)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), float((got - ref).abs().max())


@pytest.mark.parametrize("T", [1, 2, 9])
def test_upsample_and_vocoder_equal_sibling(pair, T):
    cfg, m, orc = pair
    h = torch.randn(1, cfg.hidden_size, T, generator=torch.Generator().manual_seed(10 + T))
    with torch.no_grad():
        x = h
        for blocks in m.upsample:
            for b in blocks:
                x = b(x)
        for b in m.decoder:
            x = b(x)
        ref = x.clamp(min=-1, max=1).reshape(-1)
        got = orc.upsample_and_vocode(h)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), float((got - ref).abs().max())
    # length law of the trimmed transposed convolutions (1920*T - 555 for the full-size rates)
    n = T
    for f in cfg.upsampling_ratios:
        n *= f
    for r in cfg.upsample_rates:
        n = (n - 1) * r
    assert got.numel() == n


def test_causal_trim_length_law_and_causality():
    """trans_conv_trim = "right" (the default): exactly total_upsample * T samples (1920 * T at the full-size rates, the length
    of every sample WAV the reference ships) and a strictly causal decoder — appending frames never changes earlier samples,
    and the variant is the sibling-trimmed decode shifted by one block per transposed conv (same weights, same taps)."""
    from dataclasses import replace
    from oracle.codec_oracle import CodecOracle, causal_trans_conv1d
    from qwen3_tts_cuda_graphs_b200.codec import init_codec_synthetic
    from qwen3_tts_cuda_graphs_b200.config import CodecDecoderConfig
    assert CodecDecoderConfig().trans_conv_trim == "right" and CodecDecoderConfig().n_samples(7) == 1920 * 7
    assert replace(CodecDecoderConfig(), trans_conv_trim="both").n_samples(7) == 1920 * 7 - 555
    cfg = preset("tiny").codec
    assert cfg.trans_conv_trim == "right"
    orc = CodecOracle(cfg, init_codec_synthetic(cfg, seed=3))
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, cfg.codebook_size, (7, cfg.num_quantizers), generator=g)
    full = orc.decode(codes)
    assert full.numel() == cfg.total_upsample * 7 == cfg.n_samples(7)
    for t in (1, 3, 6):
        part = orc.decode(codes[:t])
        assert part.numel() == cfg.total_upsample * t
        err = float((part - full[: part.numel()]).abs().max())
        assert err <= 2e-4, (t, err)  # fp32 matmuls of different shapes
    # one transposed conv: "right" = "both" with one more block in front
    x = torch.randn(1, 6, 9, generator=g)
    w = torch.randn(6, 4, 10, generator=g)
    b = torch.randn(4, generator=g)
    yr, yb = causal_trans_conv1d(x, w, b, 5, "right"), causal_trans_conv1d(x, w, b, 5, "both")
    assert yr.shape[-1] == 45 and yb.shape[-1] == 40 and torch.equal(yr[..., 5:], yb)


def test_dequantiser_restatement(pair):
    cfg, _, orc = pair
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, cfg.codebook_size, (7, cfg.num_quantizers), generator=g)
    got = orc.dequant(codes)[0].t()
    w = orc.w
    ref = torch.zeros(7, cfg.latent_dim)
    for t in range(7):
        first = sum(w[f"quantizer.codebook.{q}"][codes[t, q]] for q in range(cfg.num_semantic_quantizers))
        rest = sum(w[f"quantizer.codebook.{q}"][codes[t, q]] for q in range(cfg.num_semantic_quantizers, cfg.num_quantizers))
        ref[t] = w["quantizer.rvq_first.output_proj.weight"] @ first + w["quantizer.rvq_rest.output_proj.weight"] @ rest
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


def test_full_decode_is_causal_and_bounded(pair):
    """Streaming relies on causality: the first frames' samples do not change when later frames are appended."""
    cfg, _, orc = pair
    g = torch.Generator().manual_seed(4)
    codes = torch.randint(0, cfg.codebook_size, (6, cfg.num_quantizers), generator=g)
    a = orc.decode(codes[:4])
    b = orc.decode(codes)
    assert float(b.abs().max()) <= 1.0
    assert torch.allclose(a, b[: a.numel()], atol=1e-5)


def test_fuse_row_norms_rewrites_only_the_safe_pattern(monkeypatch):
    """codec.fuse_row_norms (host logic): an RMSNORM that reads exactly the rows the GEMM before it wrote becomes that GEMM's
    norm_out; anything else stays a separate op (different buffer / shape, SwiGLU or fp32 output, a tail-only GEMM, a GEMM that
    already carries a norm), and FQ3C_FUSE_NORM=0 turns the rewrite off."""
    from qwen3_tts_cuda_graphs_b200.codec import F_OUT_F32, F_RESID, F_SWIGLU, K_ATTN, K_GEMM, K_RMSNORM, Op, fuse_row_norms

    def gemm(C, M=8, N=64, flags=0, ldc=64, m_begin=0):
        o = Op()
        o.kind, o.M, o.N, o.K, o.flags, o.C, o.ldc, o.m_begin = K_GEMM, M, N, 64, flags, C, ldc, m_begin
        return o

    def norm(A, out, M=8, N=64, lda=64, rows=None):
        o = Op()
        o.kind, o.M, o.N, o.A, o.lda, o.a_rows, o.C, o.ldc, o.scale, o.f0 = K_RMSNORM, M, N, A, lda, M if rows is None else rows, out, 64, 0x9000, 1e-6
        return o

    ops = fuse_row_norms([gemm(0x1000, flags=F_RESID), norm(0x1000, 0x2000)])
    assert len(ops) == 1 and ops[0].norm_out == 0x2000 and ops[0].norm_w == 0x9000 and ops[0].norm_ld == 64 and abs(ops[0].norm_eps - 1e-6) < 1e-12
    keep = [
        [gemm(0x1000), norm(0x1100, 0x2000)],                    # reads another buffer
        [gemm(0x1000), norm(0x1000, 0x2000, N=32)],              # another width
        [gemm(0x1000), norm(0x1000, 0x2000, M=4)],               # another row count
        [gemm(0x1000, flags=F_SWIGLU), norm(0x1000, 0x2000)],    # the GEMM's output is not [M, N] bf16
        [gemm(0x1000, flags=F_OUT_F32), norm(0x1000, 0x2000)],
        [gemm(0x1000, m_begin=3), norm(0x1000, 0x2000)],         # tail-only GEMM: rows below m_begin are not written
        [norm(0x1000, 0x2000)],                                  # nothing in front
    ]
    for lst in keep:
        out = fuse_row_norms(list(lst))
        assert len(out) == len(lst) and not any(o.norm_out for o in out if o.kind == K_GEMM)
    a = Op()
    a.kind = K_ATTN
    ops = fuse_row_norms([gemm(0x1000), a, norm(0x1000, 0x2000)])  # not adjacent
    assert len(ops) == 3
    ops = fuse_row_norms([gemm(0x1000), norm(0x1000, 0x2000), norm(0x1000, 0x3000)])  # one norm per GEMM
    assert len(ops) == 2 and ops[0].norm_out == 0x2000 and ops[1].kind == K_RMSNORM
    monkeypatch.setenv("FQ3C_FUSE_NORM", "0")
    assert len(fuse_row_norms([gemm(0x1000), norm(0x1000, 0x2000)])) == 2
