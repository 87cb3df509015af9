"""`FasterQwen3TTS.from_pretrained(<checkpoint directory>)` (reference: faster_qwen3_tts/model.py:71-163) on a directory of the
tiny architecture written by tests/helpers.write_tiny_checkpoint: safetensors -> arena / codec decoder, the directory's
tokenizer, the speaker encoder and the codec encoder for the reference clip."""
import dataclasses
import wave

import numpy as np
import pytest
import torch

from helpers import TINY_SPEAKER_CFG, write_tiny_checkpoint
from qwen3_tts_cuda_graphs_b200.config import preset

pytestmark = pytest.mark.gpu

TEXT = "Hello world, the quick brown fox!"


@pytest.fixture(scope="module")
def ref_wav(tmp_path_factory):
    p = tmp_path_factory.mktemp("audio") / "ref.wav"
    sr = 16000  # not the encoders' rate: the front end resamples
    t = np.arange(int(1.6 * sr)) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) * 32767 / 1.5).astype(np.int16)
    with wave.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    return str(p)


@pytest.fixture(scope="module")
def loaded(tmp_path_factory):
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS

    d = str(tmp_path_factory.mktemp("ckpt") / "Qwen3-TTS-12Hz-tiny-Base")
    cfg0 = dataclasses.replace(preset("tiny-Base"), tts_model_size="0b6")
    cfg, lm, codec, spk, mimi = write_tiny_checkpoint(d, cfg=cfg0, shards=2, mimi_codebooks=True)
    tts = FasterQwen3TTS.from_pretrained(d, device="cuda", dtype=torch.bfloat16, max_seq_len=256)
    yield tts, cfg, lm, codec, spk, mimi
    tts.model.engine.close()


def test_directory_load_equals_in_memory_weights(loaded):
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
    from qwen3_tts_cuda_graphs_b200.base_model import HFTokenizer

    tts, cfg, lm, codec, _, _ = loaded
    assert tts.model.cfg == cfg and isinstance(tts.model.tokenizer, HFTokenizer) and tts.model.frontend is not None
    ref = FasterQwen3TTS.from_pretrained("tiny-Base", device="cuda", dtype=torch.bfloat16, max_seq_len=256, weights=lm, cfg=cfg)
    try:
        assert tts.model.arena.offsets == ref.model.arena.offsets
        assert torch.equal(tts.model.arena.buf, ref.model.arena.buf)
        codes = torch.randint(0, cfg.codec.codebook_size, (1, 9, 16), device="cuda")
        a, sr = tts.model.model.speech_tokenizer.decode({"audio_codes": codes})
        b, _ = ref.model.model.speech_tokenizer.decode({"audio_codes": codes})  # synthetic(seed + 1) == the exported decoder
        assert sr == 24000 and torch.equal(a[0], b[0])
    finally:
        ref.model.engine.close()


def test_xvector_voice_clone_uses_the_speaker_encoder(loaded, ref_wav):
    from qwen3_tts_cuda_graphs_b200 import frontend as fe

    tts, cfg, _, _, spk, _ = loaded
    audio, sr = tts.generate_voice_clone(TEXT, "English", ref_wav, "", max_new_tokens=10, do_sample=False)
    assert sr == 24000 and audio[0].dtype == np.float32 and audio[0].size == cfg.codec.n_samples(10) and np.isfinite(audio[0]).all()
    (vcp, _), = [v for k, v in tts._voice_prompt_cache.items() if k[0] == ref_wav and k[2]]
    got = vcp["ref_spk_embedding"][0].float().cpu()
    wav, wsr = fe.load_audio(ref_wav)
    want = fe.SpeakerEncoder(TINY_SPEAKER_CFG, spk, "cpu").embed_wav(wav, wsr)
    assert got.shape == (cfg.talker.hidden_size,)
    assert float((got - want).abs().max()) <= 2e-2 * float(want.abs().max())  # bf16 storage of the x-vector + GPU fp32 convs


def test_icl_voice_clone_uses_the_codec_encoder(loaded, ref_wav):
    from qwen3_tts_cuda_graphs_b200 import frontend as fe

    tts, cfg, _, _, _, mimi = loaded
    chunks = list(tts.generate_voice_clone_streaming(TEXT, "English", ref_wav, "reference clip says this.", max_new_tokens=16,
                                                     do_sample=False, xvec_only=False, chunk_size=8))
    assert len(chunks) == 2 and all(np.isfinite(c[0]).all() for c in chunks)
    (vcp, ref_ids), = [v for k, v in tts._voice_prompt_cache.items() if k[0] == ref_wav and not k[2]]
    wav, wsr = fe.load_audio(ref_wav)
    wav = np.concatenate([wav, np.zeros(int(0.5 * wsr), np.float32)])  # model.py:198-200
    with torch.no_grad():
        want = mimi.encode(torch.from_numpy(fe.resample(wav, wsr, 24000)).reshape(1, 1, -1), num_quantizers=16).audio_codes[0].T
    got = vcp["ref_code"][0].cpu()
    assert got.shape == want.shape == (27, 16)  # 2.1 s at 12.5 frames per second, rounded up
    assert float((got == want).float().mean()) >= 0.9  # nearest-neighbour search: GPU/CPU fp32 summation order can flip near-ties
    assert ref_ids[0] is not None and ref_ids[0].shape[1] == 3 + 5 + 2


def test_hub_id_without_a_local_copy_raises():
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS

    with pytest.raises(FileNotFoundError, match="never downloads"):
        FasterQwen3TTS.from_pretrained("Qwen/Qwen3-TTS-12Hz-0.6B-Base", device="cuda")
