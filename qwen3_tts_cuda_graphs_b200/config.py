"""Architecture configs for the Qwen3-TTS hot path (talker, code predictor, 12 Hz codec decoder).

Every dimension the kernels use is read from these dataclasses, never hard-coded: the numbers below
are the ones SURVEY.md §8(c) records for the public checkpoints (recalled, not verifiable offline),
so a real `config.json` can override any of them through `TTSConfig.from_dict`.

Reference provenance: talker/predictor dims are what `faster_qwen3_tts/talker_graph.py:27-59` and
`faster_qwen3_tts/predictor_graph.py:34-76` read from the HF configs of `qwen_tts`; special-token ids are
the attributes used by `faster_qwen3_tts/model.py:367-424` and `faster_qwen3_tts/generate.py:41-50`.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field, replace
from typing import Dict, Optional, Tuple


@dataclass(frozen=True)
class StackConfig:
    """One dense Qwen3 decoder stack (per-head q/k RMSNorm before RoPE, GQA, SwiGLU)."""

    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    vocab_size: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    sliding_window: Optional[int] = None  # the reference builds causal masks only when this is None

    @property
    def q_dim(self) -> int:
        return self.num_attention_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_key_value_heads * self.head_dim

    @property
    def qkv_dim(self) -> int:
        return self.q_dim + 2 * self.kv_dim

    def layer_params(self) -> int:
        h, i = self.hidden_size, self.intermediate_size
        return self.qkv_dim * h + h * self.q_dim + 3 * h * i + 2 * h + 2 * self.head_dim


@dataclass(frozen=True)
class TalkerConfig(StackConfig):
    """`talker_config` of the reference (`model.py:115`)."""

    num_code_groups: int = 16
    text_vocab_size: int = 151936
    text_hidden_size: int = 2048
    codec_eos_token_id: int = 2150
    codec_pad_id: int = 2148
    codec_bos_id: int = 2149
    codec_think_id: int = 2154
    codec_nothink_id: int = 2155
    codec_think_bos_id: int = 2156
    codec_think_eos_id: int = 2157
    codec_language_id: Dict[str, int] = field(
        default_factory=lambda: {
            "chinese": 2055, "english": 2050, "german": 2053, "italian": 2070, "portuguese": 2071,
            "spanish": 2054, "japanese": 2058, "korean": 2064, "french": 2061, "russian": 2069,
            "beijing_dialect": 2074, "sichuan_dialect": 2062,
        }
    )
    spk_id: Dict[str, int] = field(default_factory=dict)
    spk_is_dialect: Dict[str, object] = field(default_factory=dict)


@dataclass(frozen=True)
class PredictorConfig(StackConfig):
    """`talker.code_predictor.model.config` of the reference (`model.py:118-119`)."""

    num_code_groups: int = 16

    @property
    def num_codebooks(self) -> int:  # predictor_graph.py:44-46
        return self.num_code_groups - 1


@dataclass(frozen=True)
class CodecDecoderConfig:
    """12 Hz speech-tokenizer decoder (sibling: transformers Qwen3OmniMoeCode2Wav)."""

    codebook_size: int = 2048
    num_quantizers: int = 16
    num_semantic_quantizers: int = 1
    codebook_dim: int = 256
    latent_dim: int = 512
    hidden_size: int = 1024
    intermediate_size: int = 3072
    num_hidden_layers: int = 8
    num_attention_heads: int = 16
    num_key_value_heads: int = 16
    head_dim: int = 64
    sliding_window: int = 72
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    layer_scale_initial_scale: float = 0.01
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    decoder_dim: int = 1536
    sample_rate: int = 24000
    # How the vocoder's transposed convolutions (kernel 2r, stride r) are trimmed.
    #   "right": causal — drop the last r samples only: T frames give exactly total_upsample * T samples (every sample WAV the
    #            reference ships is k x 1920 samples long) and sample n depends on frames <= n // 1920 alone;
    #   "both":  r on each side, as the in-container sibling (transformers Qwen3OmniMoeCausalTransConvNet) does:
    #            1920 * T - 555 samples for the full-size rates, and a block looks one input row ahead.
    trans_conv_trim: str = "right"

    def n_samples(self, frames: int) -> int:
        rows = frames
        for f in self.upsampling_ratios:
            rows *= f
        for r in self.upsample_rates:
            rows = rows * r if self.trans_conv_trim == "right" else (rows - 1) * r
        return max(rows, 0)

    @property
    def total_upsample(self) -> int:
        n = 1
        for r in self.upsampling_ratios + self.upsample_rates:
            n *= r
        return n


@dataclass(frozen=True)
class TTSConfig:
    """Top-level config (`base.model.config` in the reference)."""

    talker: TalkerConfig
    predictor: PredictorConfig
    codec: CodecDecoderConfig
    tts_model_type: str = "base"  # base | custom_voice | voice_design (model.py:843, :1017)
    tts_model_size: str = "0b6"  # model.py:849
    tts_bos_token_id: int = 151672
    tts_eos_token_id: int = 151673
    tts_pad_token_id: int = 151671
    im_start_token_id: int = 151644
    im_end_token_id: int = 151645
    assistant_token_id: int = 77091
    user_token_id: int = 872
    newline_token_id: int = 198
    speaker_embed_dim: int = 0  # 0 => talker hidden size

    def to_dict(self) -> dict:
        return asdict(self)

    @staticmethod
    def from_dict(d: dict) -> "TTSConfig":
        d = dict(d)
        talker = TalkerConfig(**d.pop("talker"))
        predictor = PredictorConfig(**d.pop("predictor"))
        codec_d = d.pop("codec")
        for k in ("upsampling_ratios", "upsample_rates"):
            codec_d[k] = tuple(codec_d[k])
        return TTSConfig(talker=talker, predictor=predictor, codec=CodecDecoderConfig(**codec_d), **d)


_CUSTOM_SPEAKERS = {
    "aiden": 3000, "vivian": 3001, "serena": 3002, "uncle_fu": 3003, "dylan": 3004,
    "eric": 3005, "ryan": 3006, "ono_anna": 3007, "sohee": 3008,
}
_CUSTOM_DIALECTS = {k: False for k in _CUSTOM_SPEAKERS}
_CUSTOM_DIALECTS.update({"dylan": "beijing_dialect", "eric": "sichuan_dialect"})


def _talker(hidden: int, inter: int, **kw) -> TalkerConfig:
    return TalkerConfig(
        hidden_size=hidden, intermediate_size=inter, num_hidden_layers=28, num_attention_heads=16,
        num_key_value_heads=8, head_dim=128, vocab_size=3072, **kw,
    )


def _predictor() -> PredictorConfig:
    return PredictorConfig(
        hidden_size=1024, intermediate_size=3072, num_hidden_layers=5, num_attention_heads=16,
        num_key_value_heads=8, head_dim=128, vocab_size=2048,
    )


def preset(name: str) -> TTSConfig:
    """Named architecture presets.

    "0.6B-Base", "1.7B-Base", "0.6B-CustomVoice", "1.7B-CustomVoice", "1.7B-VoiceDesign" are the
    checkpoints BASELINE.json's configs name; "tiny" is a 2-layer stand-in (same head_dim/GQA ratio)
    used by CPU tests and the committed golden fixtures.
    """
    key = name.lower().replace("qwen/", "").replace("qwen3-tts-12hz-", "").replace("synthetic://", "")
    size, _, kind = key.partition("-")
    kind = {"": "base", "base": "base", "customvoice": "custom_voice", "voicedesign": "voice_design"}.get(kind, kind)
    if size == "tiny":
        talker = TalkerConfig(
            hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, head_dim=128, vocab_size=3072, text_vocab_size=512, text_hidden_size=128,
            spk_id=dict(_CUSTOM_SPEAKERS), spk_is_dialect=dict(_CUSTOM_DIALECTS),
        )
        pred = PredictorConfig(
            hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, head_dim=128, vocab_size=2048,
        )
        codec = CodecDecoderConfig(
            codebook_dim=32, latent_dim=64, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
            num_attention_heads=4, num_key_value_heads=4, head_dim=16, sliding_window=8, decoder_dim=128,
        )
        return TTSConfig(
            talker=talker, predictor=pred, codec=codec, tts_model_type=kind if kind != "" else "base",
            tts_model_size="tiny", tts_bos_token_id=501, tts_eos_token_id=502, tts_pad_token_id=500,
            im_start_token_id=503, im_end_token_id=504, assistant_token_id=505, user_token_id=506,
            newline_token_id=507,
        )
    if size in ("0.6b", "0b6"):
        hidden, inter, sz = 1024, 3072, "0b6"
    elif size in ("1.7b", "1b7"):
        hidden, inter, sz = 2048, 6144, "1b7"
    else:
        raise ValueError(f"unknown Qwen3-TTS preset {name!r}")
    kw = {}
    if kind == "custom_voice":
        kw = dict(spk_id=dict(_CUSTOM_SPEAKERS), spk_is_dialect=dict(_CUSTOM_DIALECTS))
    return TTSConfig(
        talker=_talker(hidden, inter, **kw), predictor=_predictor(), codec=CodecDecoderConfig(),
        tts_model_type=kind, tts_model_size=sz,
    )


def with_layers(cfg: TTSConfig, talker_layers: int, predictor_layers: Optional[int] = None) -> TTSConfig:
    """Same architecture with fewer layers (parity tests that must finish in seconds on CPU)."""
    t = replace(cfg.talker, num_hidden_layers=talker_layers)
    p = cfg.predictor if predictor_layers is None else replace(cfg.predictor, num_hidden_layers=predictor_layers)
    return replace(cfg, talker=t, predictor=p)
