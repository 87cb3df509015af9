"""Reference-audio front end: wav -> x-vector (speaker encoder) and wav -> reference codes (codec encoder).

In the reference this is `self.model.create_voice_clone_prompt(ref_audio=..., ref_text=..., x_vector_only_mode=...)`
(`faster_qwen3_tts/model.py:234-254`), a call into the un-vendored `qwen_tts` package; its result is cached per voice
(`_voice_prompt_cache`, `model.py:230-233`), so this code runs once per voice and is OFF the steady-state path — it is host-side
plumbing over torch ops on the model's device, not one of the hand-written kernels.

* `load_audio` / `resample`: what `sf.read(..., dtype="float32")` + mono mix-down do at `model.py:194-200` (soundfile is not in
  this image: PCM / float WAV through the standard library and scipy).
* `mel_spectrogram`: the 24 kHz log-mel the speaker encoder consumes (n_fft 1024, hop 256, 128 slaney mels, 0–12 kHz,
  reflect-padded, log(clamp(·, 1e-5))) — parameters from `speaker_encoder_config`, these are the recalled defaults.
* `SpeakerEncoder`: ECAPA-TDNN (TDNN → 3 × SE-Res2Net → multi-layer aggregation → attentive statistics pooling → 1×1 conv) as
  pure functions over a state dict.  Pinned against the in-container sibling `transformers...qwen2_5_omni.ECAPA_TimeDelayNet`
  run on the same weights (`tests/test_frontend_cpu.py`); upstream's `speaker_encoder.*` key names are recalled, not verified.
* `CodecEncoder`: the 12.5 Hz Mimi-style encoder (SEANet convs → 8-layer transformer → stride-2 downsample → split RVQ
  nearest-neighbour search).  This one wraps the library implementation `transformers.MimiModel` (its encoder half) instead of
  restating it: once per voice, ~1 s of audio per ms, nothing to win by hand.
"""
from __future__ import annotations

import math
import wave
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# audio in
# ------------------------------------------------------------------------------------------------
def load_audio(ref_audio) -> Tuple[np.ndarray, int]:
    """(float32 mono samples in [-1, 1], sample rate) from a path or an (array, sr) pair (model.py:194-197)."""
    if isinstance(ref_audio, (tuple, list)) and len(ref_audio) == 2 and not isinstance(ref_audio[0], str):
        a = np.asarray(ref_audio[0], dtype=np.float32)
        if a.ndim > 1:
            a = a.mean(axis=1 if a.shape[1] < a.shape[0] else 0)
        return np.ascontiguousarray(a), int(ref_audio[1])
    path = str(ref_audio)
    try:
        with wave.open(path, "rb") as wf:
            sr, nch, width, n = wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()
            raw = wf.readframes(n)
        if width == 2:
            a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 1:
            a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            a = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
        else:
            raise ValueError(f"{path}: unsupported PCM width {width}")
        a = a.reshape(-1, nch)
    except wave.Error:  # IEEE-float / extensible WAV
        from scipy.io import wavfile

        sr, a = wavfile.read(path)
        if a.dtype.kind == "i":
            a = a.astype(np.float32) / float(2 ** (8 * a.dtype.itemsize - 1))
        elif a.dtype.kind == "u":
            a = (a.astype(np.float32) - 128.0) / 128.0
        a = a.astype(np.float32).reshape(len(a), -1)
    return np.ascontiguousarray(a.mean(axis=1), dtype=np.float32), int(sr)


def resample(x: np.ndarray, sr: int, target: int) -> np.ndarray:
    if sr == target or len(x) == 0:
        return x.astype(np.float32, copy=False)
    from scipy.signal import resample_poly

    g = math.gcd(int(sr), int(target))
    return resample_poly(x, target // g, sr // g).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# log-mel
# ------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log_start = 1000.0 / (200.0 / 3)
    return np.where(f >= 1000.0, log_start + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    log_start = 1000.0 / (200.0 / 3)
    return np.where(m >= log_start, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - log_start)), m * (200.0 / 3))


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> torch.Tensor:
    """Slaney-scale, slaney-normalised triangular filters [n_mels, n_fft // 2 + 1] (librosa.filters.mel defaults)."""
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(pts)
    ramps = pts[:, None] - freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return torch.from_numpy(w.astype(np.float32))


def mel_spectrogram(wav: torch.Tensor, sr: int = 24000, n_fft: int = 1024, hop: int = 256, win: int = 1024, n_mels: int = 128,
                    fmin: float = 0.0, fmax: Optional[float] = 12000.0) -> torch.Tensor:
    """wav [N] float32 -> log-mel [frames, n_mels] (frames = N // hop for N a multiple of hop)."""
    fmax = sr / 2.0 if fmax is None else fmax
    pad = (n_fft - hop) // 2
    x = F.pad(wav.reshape(1, 1, -1).float(), (pad, pad), mode="reflect").reshape(1, -1)
    window = torch.hann_window(win, device=wav.device, dtype=torch.float32)
    spec = torch.stft(x, n_fft, hop_length=hop, win_length=win, window=window, center=False, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)[0]  # [freq, frames]
    mel = mel_filterbank(sr, n_fft, n_mels, fmin, fmax).to(wav.device) @ mag
    return torch.log(torch.clamp(mel, min=1e-5)).transpose(0, 1).contiguous()


# ------------------------------------------------------------------------------------------------
# speaker encoder
# ------------------------------------------------------------------------------------------------
SPEAKER_ENCODER_DEFAULTS = dict(
    mel_dim=128, enc_dim=1024, enc_channels=(512, 512, 512, 512, 1536), enc_kernel_sizes=(5, 3, 3, 3, 1),
    enc_dilations=(1, 2, 3, 4, 1), enc_attention_channels=128, enc_res2net_scale=8, enc_se_channels=128, sample_rate=24000,
)


def speaker_encoder_specs(c: dict) -> list:
    """(name, shape) of every tensor of the ECAPA stack described by config dict `c`."""
    ch, ks = list(c["enc_channels"]), list(c["enc_kernel_sizes"])
    sc, se, att = c["enc_res2net_scale"], c["enc_se_channels"], c["enc_attention_channels"]

    def conv(p, cout, cin, k):
        return [(f"{p}.weight", (cout, cin, k)), (f"{p}.bias", (cout,))]

    s = conv("blocks.0.conv", ch[0], c["mel_dim"], ks[0])
    for i in range(1, len(ch) - 1):
        p = f"blocks.{i}"
        s += conv(f"{p}.tdnn1.conv", ch[i], ch[i - 1], 1)
        for j in range(sc - 1):
            s += conv(f"{p}.res2net_block.blocks.{j}.conv", ch[i] // sc, ch[i] // sc, ks[i])
        s += conv(f"{p}.tdnn2.conv", ch[i], ch[i], 1)
        s += conv(f"{p}.se_block.conv1", se, ch[i], 1) + conv(f"{p}.se_block.conv2", ch[i], se, 1)
    s += conv("mfa.conv", ch[-1], ch[-1], ks[-1])
    s += conv("asp.tdnn.conv", att, ch[-1] * 3, 1) + conv("asp.conv", ch[-1], att, 1)
    s += conv("fc", c["enc_dim"], ch[-1] * 2, 1)
    return s


def init_speaker_encoder_synthetic(c: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in speaker_encoder_specs(c):
        if name.endswith(".bias"):
            out[name] = 0.02 * torch.randn(shape, generator=g)
        else:
            out[name] = torch.randn(shape, generator=g) / math.sqrt(shape[1] * shape[2])
    return out


class SpeakerEncoder:
    """ECAPA-TDNN forward over a state dict; fp32 on `device`."""

    def __init__(self, cfg: Optional[dict], weights: Dict[str, torch.Tensor], device="cpu"):
        self.cfg = dict(SPEAKER_ENCODER_DEFAULTS)
        self.cfg.update({k: v for k, v in (cfg or {}).items() if k in self.cfg and v is not None})
        self.device = torch.device(device)
        self.w = {}
        for name, shape in speaker_encoder_specs(self.cfg):
            if name not in weights:
                raise KeyError(f"speaker encoder tensor {name!r} is missing")
            t = weights[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"speaker encoder tensor {name!r} has shape {tuple(t.shape)}, the config implies {shape}")
            self.w[name] = t.to(self.device, torch.float32).contiguous()

    def _conv(self, x, p, dilation=1, relu=True):
        w = self.w[f"{p}.weight"]
        total = dilation * (w.shape[-1] - 1)
        if total:  # padding="same", padding_mode="reflect"
            x = F.pad(x, (total // 2, total - total // 2), mode="reflect")
        y = F.conv1d(x, w, self.w[f"{p}.bias"], dilation=dilation)
        return F.relu(y) if relu else y

    def _se_res2net(self, x, p, k, dilation):
        sc = self.cfg["enc_res2net_scale"]
        h = self._conv(x, f"{p}.tdnn1.conv")
        parts, prev = [], None
        for i, part in enumerate(torch.chunk(h, sc, dim=1)):
            if i == 0:
                prev = part
            elif i == 1:
                prev = self._conv(part, f"{p}.res2net_block.blocks.0.conv", dilation)
            else:
                prev = self._conv(part + prev, f"{p}.res2net_block.blocks.{i - 1}.conv", dilation)
            parts.append(prev)
        h = self._conv(torch.cat(parts, dim=1), f"{p}.tdnn2.conv")
        s = h.mean(dim=2, keepdim=True)
        s = torch.sigmoid(self._conv(self._conv(s, f"{p}.se_block.conv1"), f"{p}.se_block.conv2", relu=False))
        return h * s + x

    @staticmethod
    def _stats(x, wgt, eps=1e-12):
        mean = (wgt * x).sum(2)
        std = torch.sqrt((wgt * (x - mean.unsqueeze(2)).pow(2)).sum(2).clamp(eps))
        return mean, std

    @torch.inference_mode()
    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, frames, mel_dim] -> x-vector [B, enc_dim]."""
        c = self.cfg
        x = mel.to(self.device, torch.float32).transpose(1, 2)
        ks, ds = list(c["enc_kernel_sizes"]), list(c["enc_dilations"])
        x = self._conv(x, "blocks.0.conv", ds[0])
        feats = []
        for i in range(1, len(ks) - 1):
            x = self._se_res2net(x, f"blocks.{i}", ks[i], ds[i])
            feats.append(x)
        h = self._conv(torch.cat(feats, dim=1), "mfa.conv", ds[-1])
        L = h.shape[-1]
        mean, std = self._stats(h, torch.full((1, 1, L), 1.0 / L, device=h.device))
        att = torch.cat([h, mean.unsqueeze(2).expand(-1, -1, L), std.unsqueeze(2).expand(-1, -1, L)], dim=1)
        att = self._conv(torch.tanh(self._conv(att, "asp.tdnn.conv")), "asp.conv", relu=False)
        mean, std = self._stats(h, F.softmax(att, dim=2))
        pooled = torch.cat([mean, std], dim=1).unsqueeze(2)
        return self._conv(pooled, "fc", relu=False).squeeze(-1)

    def embed_wav(self, wav: np.ndarray, sr: int) -> torch.Tensor:
        """float32 mono samples -> x-vector [enc_dim] (resampled to the encoder's rate first)."""
        c = self.cfg
        x = torch.from_numpy(resample(wav, sr, c["sample_rate"])).to(self.device)
        mel = mel_spectrogram(x, sr=c["sample_rate"], n_mels=c["mel_dim"])
        return self(mel.unsqueeze(0))[0]


# ------------------------------------------------------------------------------------------------
# codec encoder
# ------------------------------------------------------------------------------------------------
_MIMI_ENCODER_SIDE = ("encoder.", "encoder_transformer.", "downsample.", "quantizer.")


class CodecEncoder:
    """wav -> codes [T, n_quantizers] through `transformers.MimiModel`'s encoder half (library code, once per voice)."""

    def __init__(self, encoder_config: dict, weights: Dict[str, torch.Tensor], n_quantizers: int, device="cpu"):
        from transformers import MimiConfig, MimiModel

        known = set(MimiConfig().to_dict())
        self.config = MimiConfig(**{k: v for k, v in (encoder_config or {}).items() if k in known})
        self.n_quantizers = int(n_quantizers)
        self.sample_rate = int(self.config.sampling_rate)
        self.device = torch.device(device)
        model = MimiModel(self.config)
        res = model.load_state_dict(weights, strict=False)
        missing = [k for k in res.missing_keys if k.startswith(_MIMI_ENCODER_SIDE) and not k.endswith(("initialized",))]
        if missing:
            raise KeyError(f"codec encoder: {len(missing)} tensors missing, e.g. {missing[:4]}")
        self.model = model.to(self.device).eval()

    @torch.inference_mode()
    def encode(self, wav: np.ndarray, sr: int) -> torch.Tensor:
        x = torch.from_numpy(resample(wav, sr, self.sample_rate)).to(self.device).reshape(1, 1, -1)
        codes = self.model.encode(x, num_quantizers=self.n_quantizers, return_dict=True).audio_codes  # [1, Q, T]
        return codes[0].transpose(0, 1).contiguous().to(torch.long)


# ------------------------------------------------------------------------------------------------
# the pair, as the base model holds it
# ------------------------------------------------------------------------------------------------
class VoiceFrontEnd:
    def __init__(self, speaker: SpeakerEncoder, codec: Optional[CodecEncoder]):
        self.speaker, self.codec = speaker, codec

    def x_vector(self, ref_audio) -> torch.Tensor:
        wav, sr = load_audio(ref_audio)
        return self.speaker.embed_wav(wav, sr)

    def ref_codes(self, ref_audio) -> torch.Tensor:
        if self.codec is None:
            raise RuntimeError("this checkpoint has no codec encoder weights: only x_vector_only_mode voice cloning is available")
        wav, sr = load_audio(ref_audio)
        return self.codec.encode(wav, sr)
