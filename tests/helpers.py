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


# ---- a checkpoint directory of the tiny preset in the layout qwen3_tts_cuda_graphs_b200/checkpoint.py reads ----
TINY_WORDS = ["<unk>", "assistant", "user", "hello", "world", "short", "parity", "test", "the", "quick", "brown", "fox", "jumps",
              "over", "lazy", "dog", "speak", "slowly", "calm", "voice", "reference", "clip", "says", "this", ".", ",", "!", "?"]

TINY_SPEAKER_CFG = dict(mel_dim=128, enc_dim=256, enc_channels=(32, 32, 32, 32, 96), enc_kernel_sizes=(5, 3, 3, 3, 1),
                        enc_dilations=(1, 2, 3, 4, 1), enc_attention_channels=16, enc_res2net_scale=4, enc_se_channels=16,
                        sample_rate=24000)


def tiny_mimi_config():
    from transformers import MimiConfig

    return MimiConfig(hidden_size=32, num_filters=8, num_hidden_layers=2, intermediate_size=64, num_attention_heads=4,
                      num_key_value_heads=4, codebook_size=2048, codebook_dim=16, vector_quantization_hidden_dimension=16,
                      num_quantizers=16, upsample_groups=32, sliding_window=16)


def write_tiny_tokenizer(path, cfg):
    """tokenizer.json whose chat-template pieces map to the tiny preset's special ids (config.preset("tiny"))."""
    import os

    from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers

    vocab = {w: 16 + i for i, w in enumerate(TINY_WORDS)}
    vocab["<unk>"] = 15
    vocab.update({"<|im_start|>": cfg.im_start_token_id, "<|im_end|>": cfg.im_end_token_id, "assistant": cfg.assistant_token_id,
                  "user": cfg.user_token_id, "\n": cfg.newline_token_id})
    used = set(vocab.values())
    vocab.update({f"<reserved_{i}>": i for i in range(max(used) + 1) if i not in used})  # a dense id range, like a real vocab.json
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.normalizer = normalizers.Lowercase()
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(r"<\|im_start\|>|<\|im_end\|>|\n|[.,!?]"), behavior="isolated"),
        pre_tokenizers.Split(Regex(r" +"), behavior="removed"),
    ])
    os.makedirs(path, exist_ok=True)
    tok.save(os.path.join(path, "tokenizer.json"))


def write_tiny_checkpoint(path, cfg=None, seed=0, shards=1, mimi_codebooks=False, with_frontend=True):
    """Returns (cfg, lm weights, codec weights, speaker-encoder weights | None, MimiModel | None)."""
    from qwen3_tts_cuda_graphs_b200 import checkpoint as ck
    from qwen3_tts_cuda_graphs_b200 import frontend as fe
    from qwen3_tts_cuda_graphs_b200.codec import init_codec_synthetic

    cfg = cfg or preset("tiny-Base")
    lm = init_synthetic(cfg, seed=seed, norm_jitter=0.1)
    codec = init_codec_synthetic(cfg.codec, seed=seed + 1)
    extra = codec_extra = spk = mimi = None
    extra_cfg, codec_extra_cfg = {}, {}
    if with_frontend:
        spk = fe.init_speaker_encoder_synthetic(TINY_SPEAKER_CFG, seed=seed + 2)
        extra = {"speaker_encoder." + k: v for k, v in spk.items()}
        extra_cfg["speaker_encoder_config"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in TINY_SPEAKER_CFG.items()}
        from transformers import MimiModel

        torch.manual_seed(seed + 3)
        mc = tiny_mimi_config()
        mimi = MimiModel(mc).eval()
        keep = ("encoder.", "encoder_transformer.", "downsample.", "quantizer.")
        codec_extra = {"encoder." + k: v.detach().clone() for k, v in mimi.state_dict().items() if k.startswith(keep)}
        codec_extra_cfg["encoder_config"] = mc.to_dict()
    ck.export_checkpoint(path, cfg, lm, codec, extra=extra, codec_extra=codec_extra, shards=shards, mimi_codebooks=mimi_codebooks,
                         extra_config=extra_cfg, codec_extra_config=codec_extra_cfg)
    write_tiny_tokenizer(path, cfg)
    return cfg, lm, codec, spk, mimi
