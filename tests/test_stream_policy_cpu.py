"""SURVEY.md §8 row a11, the POLICY half: `model.WindowedDecode` (used by the public streaming generators and by the serving
scheduler) against the oracle restatement of the reference's streaming decode policy (`oracle/stream_policy_oracle.py`,
`faster_qwen3_tts/model.py:737-826`), sample for sample, around a toy causal decoder with unbounded memory —
so a wrong window, a wrong cut or a wrong calibration changes samples.  The GPU tests hold the real decoder's arithmetic; this one
holds the bookkeeping: accumulate -> calibrate at max(25, chunk_size) frames -> 25-frame left context, ICL reference codes in
front while accumulating, both length laws (exact 1920 T, and 1920 T - 555 where samples-per-frame is fractional)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.stream_policy_oracle import stream_decode_policy  # noqa: E402

from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer  # noqa: E402
from qwen3_tts_cuda_graphs_b200.model import WindowedDecode  # noqa: E402

SPF = 1920


def toy_wave(codes: np.ndarray, trim: int) -> np.ndarray:
    """Causal toy decoder with unbounded memory: frame t -> SPF samples of a running state s_t = 0.9 s_{t-1} + key_t (zero at the
    start of whatever was passed in) plus a ramp over the sample offset.  A decode of a window therefore differs from a decode of
    the whole history, so the comparison is sensitive to WHICH frames a window holds, not only to where it is cut; `trim` samples
    are cut from the end (the "both" trim mode's length law 1920 T - 555)."""
    key = codes[:, 0].astype(np.float64) * 3.0 + codes[:, 1].astype(np.float64)
    state, s = np.empty(len(key)), 0.0
    for t, k in enumerate(key):
        s = 0.9 * s + k
        state[t] = s
    wav = (state[:, None] * 1e-4 + np.arange(SPF)[None, :] * 1e-6).astype(np.float32).reshape(-1)
    return wav[: len(wav) - trim] if trim else wav


class ToyDecoder:
    """codec.CodecDecoder's surface as WindowedDecode uses it: cfg.trans_conv_trim, n_samples(T), decode(codes, skip_samples) — the
    first `skip_samples` samples are not computed (NaN here: using one of them fails the comparison)."""

    def __init__(self, trim_mode: str):
        self.cfg = types.SimpleNamespace(trans_conv_trim=trim_mode, sample_rate=24000)
        self.trim = 0 if trim_mode == "right" else 555
        self.calls = []

    def n_samples(self, T: int) -> int:
        return T * SPF - self.trim

    def decode(self, codes: torch.Tensor, skip_samples: int = 0) -> torch.Tensor:
        self.calls.append((int(codes.shape[0]), int(skip_samples)))
        wav = toy_wave(codes.numpy(), self.trim).copy()
        wav[:skip_samples] = np.nan
        return torch.from_numpy(wav)


def _chunks(n_frames: int, chunk: int, seed: int):
    g = np.random.default_rng(seed)
    codes = g.integers(0, 2048, size=(n_frames, 16))
    return [codes[i:i + chunk] for i in range(0, n_frames, chunk)]


@pytest.mark.parametrize("trim_mode", ["right", "both"])
@pytest.mark.parametrize("n_frames,chunk,ref_frames", [
    (61, 8, 0),      # the benchmark's chunk size, partial last chunk
    (50, 12, 0),     # the API's default chunk size
    (9, 4, 0),       # ends before calibration: accumulated decodes only
    (70, 30, 0),     # chunk_size > 25: calibration waits for a whole chunk
    (26, 1, 0),      # frame-by-frame streaming
    (45, 8, 37),     # ICL: reference codes in front while accumulating, proportional cut
    (20, 8, 11),     # ICL, never calibrated
])
def test_windowed_decode_equals_the_reference_policy(trim_mode, n_frames, chunk, ref_frames):
    chunks = _chunks(n_frames, chunk, seed=n_frames * 100 + chunk)
    ref = np.random.default_rng(7).integers(0, 2048, size=(ref_frames, 16)) if ref_frames else None
    dec = ToyDecoder(trim_mode)
    want = list(stream_decode_policy(chunks, lambda c: toy_wave(np.asarray(c), dec.trim), ref, chunk))
    wd = WindowedDecode(SpeechTokenizer(dec), None if ref is None else torch.from_numpy(ref), chunk)
    got = []
    for c in chunks:
        audio, sr = wd.push(torch.from_numpy(c))
        assert sr == 24000
        got.append(audio.numpy())
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape, (i, a.shape, b.shape)
        assert not np.isnan(a).any(), f"chunk {i}: a skipped (never computed) sample was handed out"
        assert np.array_equal(a, b), f"chunk {i} differs from the reference policy"
    total = sum(len(a) for a in got)
    if trim_mode == "right":
        assert total == n_frames * SPF  # exact length law: every frame's samples exactly once
    # the window never grows beyond 25 context frames + the new chunk once calibrated
    calibrated_at = next((sum(len(c) for c in chunks[:k + 1]) for k in range(len(chunks))
                          if sum(len(c) for c in chunks[:k + 1]) >= max(25, chunk)), None)
    if calibrated_at is not None and ref is None:
        later = dec.calls[[sum(len(c) for c in chunks[:k + 1]) for k in range(len(chunks))].index(calibrated_at) + 1:]
        assert all(T <= 25 + chunk for T, _ in later)


def test_the_comparison_notices_a_different_window_or_cut():
    """Guard of the test itself: 24 instead of 25 context frames, or a cut that is one sample off, must not pass."""
    chunks = _chunks(61, 8, seed=1)
    dec = ToyDecoder("right")
    want = list(stream_decode_policy(chunks, lambda c: toy_wave(np.asarray(c), 0), None, 8))
    wd = WindowedDecode(SpeechTokenizer(dec), None, 8, context_frames=24)
    got = [wd.push(torch.from_numpy(c))[0].numpy() for c in chunks]
    assert any(a.shape != b.shape or not np.array_equal(a, b) for a, b in zip(got, want))
    wd = WindowedDecode(SpeechTokenizer(dec), None, 8)
    got = []
    for c in chunks:
        got.append(wd.push(torch.from_numpy(c))[0].numpy())
        if wd.spf is not None:
            wd.spf = SPF + 1.0 / 25  # one sample too many per 25 context frames
    assert any(a.shape != b.shape for a, b in zip(got, want))


@pytest.mark.parametrize("trim_mode", ["right", "both"])
@pytest.mark.parametrize("ref_frames", [0, 37])
def test_non_streaming_decode_cuts_the_reference_part_like_the_reference(trim_mode, ref_frames):
    """`FasterQwen3TTS._decode_full` against the oracle's restatement of model.py:634-656 (tail-only decode of the ICL reference part
    included: no skipped sample may come out)."""
    from oracle.stream_policy_oracle import full_decode_policy

    from qwen3_tts_cuda_graphs_b200.model import FasterQwen3TTS

    codes = np.random.default_rng(3).integers(0, 2048, size=(41, 16))
    ref = np.random.default_rng(4).integers(0, 2048, size=(ref_frames, 16)) if ref_frames else None
    dec = ToyDecoder(trim_mode)
    want = full_decode_policy(codes, lambda c: toy_wave(np.asarray(c), dec.trim), ref)
    self_ = types.SimpleNamespace(_to_numpy=FasterQwen3TTS._to_numpy)
    inner = types.SimpleNamespace(speech_tokenizer=SpeechTokenizer(dec))
    got, sr = FasterQwen3TTS._decode_full(self_, inner, torch.from_numpy(codes), None if ref is None else torch.from_numpy(ref))
    assert sr == 24000 and len(got) == 1 and got[0].dtype == np.float32
    assert not np.isnan(got[0]).any() and np.array_equal(got[0], want)
    if trim_mode == "right":
        assert len(got[0]) == 41 * SPF
