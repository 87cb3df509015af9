"""Reference-audio front end (reference call site faster_qwen3_tts/model.py:234-254): the ECAPA speaker
encoder against the in-container sibling on the same weights, the log-mel against torchaudio's filterbank, the codec-encoder
wrapper against transformers.MimiModel, audio loading."""
import types
import wave

import numpy as np
import pytest
import torch

from helpers import TINY_SPEAKER_CFG, tiny_mimi_config
from qwen3_tts_cuda_graphs_b200 import frontend as fe


def test_speaker_encoder_equals_the_sibling_ecapa():
    from transformers.models.qwen2_5_omni.modeling_qwen2_5_omni import ECAPA_TimeDelayNet

    cfg = dict(TINY_SPEAKER_CFG, mel_dim=20, enc_dim=48)
    torch.manual_seed(0)
    sib = ECAPA_TimeDelayNet(types.SimpleNamespace(**cfg)).eval()
    sd = {k: v.detach().clone() for k, v in sib.state_dict().items()}
    assert set(sd) == {n for n, _ in fe.speaker_encoder_specs(cfg)}
    enc = fe.SpeakerEncoder(cfg, sd)
    mel = torch.randn(2, 57, 20)
    with torch.no_grad():
        want = sib(mel)
    got = enc(mel)
    assert got.shape == (2, 48)
    assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max()) + 1e-7


def test_speaker_encoder_names_missing_and_misshaped_tensors():
    w = fe.init_speaker_encoder_synthetic(TINY_SPEAKER_CFG)
    bad = dict(w)
    bad.pop("asp.conv.weight")
    with pytest.raises(KeyError, match="asp.conv.weight"):
        fe.SpeakerEncoder(TINY_SPEAKER_CFG, bad)
    bad = dict(w, fc=w["fc.weight"])
    bad["fc.weight"] = w["fc.weight"][:, :-1]
    with pytest.raises(ValueError, match="fc.weight"):
        fe.SpeakerEncoder(TINY_SPEAKER_CFG, bad)


def test_mel_filterbank_and_frame_count():
    import torchaudio

    want = torchaudio.functional.melscale_fbanks(513, 0.0, 12000.0, 128, 24000, norm="slaney", mel_scale="slaney").T
    assert float((fe.mel_filterbank(24000, 1024, 128, 0.0, 12000.0) - want).abs().max()) < 1e-6
    m = fe.mel_spectrogram(torch.randn(256 * 40))
    assert m.shape == (40, 128) and torch.isfinite(m).all()
    # a pure tone lights up the band that holds it
    t = torch.arange(24000) / 24000.0
    m = fe.mel_spectrogram(torch.sin(2 * np.pi * 1000.0 * t))
    centres = fe._mel_to_hz(np.linspace(fe._hz_to_mel(0.0), fe._hz_to_mel(12000.0), 130))[1:-1]
    assert abs(centres[int(m.mean(0).argmax())] - 1000.0) < 60.0


def test_codec_encoder_wrapper_equals_mimi_encode():
    from transformers import MimiModel

    mc = tiny_mimi_config()
    torch.manual_seed(1)
    mm = MimiModel(mc).eval()
    wav = (0.1 * np.random.RandomState(0).randn(24000 * 2)).astype(np.float32)
    with torch.no_grad():
        want = mm.encode(torch.from_numpy(wav).reshape(1, 1, -1), num_quantizers=16).audio_codes[0]
    enc_side = {k: v for k, v in mm.state_dict().items() if k.startswith(fe._MIMI_ENCODER_SIDE)}
    got = fe.CodecEncoder(mc.to_dict(), enc_side, 16).encode(wav, 24000)
    assert got.shape == (25, 16) and got.dtype == torch.long  # 12.5 frames per second
    assert torch.equal(got.T, want)
    with pytest.raises(KeyError, match="missing"):
        fe.CodecEncoder(mc.to_dict(), {k: v for k, v in enc_side.items() if "downsample" not in k}, 16)


def test_load_audio_pcm_stereo_pair_and_resample(tmp_path):
    sr = 16000
    t = np.arange(sr) / sr
    left, right = 0.5 * np.sin(2 * np.pi * 200 * t), 0.25 * np.sin(2 * np.pi * 300 * t)
    pcm = (np.stack([left, right], 1) * 32767).astype(np.int16)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as wf:
        wf.setnchannels(2)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    a, got_sr = fe.load_audio(p)
    assert got_sr == sr and a.dtype == np.float32 and a.shape == (sr,)
    assert np.abs(a - (left + right) / 2).max() < 1e-3  # mono mix-down (model.py:196-197)
    b, sr2 = fe.load_audio((np.stack([left, right], 1), sr))
    assert sr2 == sr and np.abs(b - (left + right) / 2).max() < 1e-6
    r = fe.resample(a, sr, 24000)
    assert r.dtype == np.float32 and len(r) == 24000
    assert fe.resample(a, sr, sr) is a or np.array_equal(fe.resample(a, sr, sr), a)
