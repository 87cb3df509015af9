"""Continuous-batching scheduler on the tiny preset: requests that join and leave a
running lock-step group produce exactly the codes and the audio of their single-stream streaming run (greedy), a length cap /
a cancellation / a bad request retire one slot without disturbing the others."""
import time
import wave

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]  # a dead scheduler thread must fail the test, not hang the suite

TEXTS = [
    "Short parity test.",
    "The quick brown fox jumps over the lazy dog and keeps running through the field.",
    "Hello world!",
    "Speak slowly, in a calm voice, and do not stop before the sentence has ended.",
    "One more request that arrives late.",
    "And the last one of the batch.",
]
LENGTHS = [40, 21, 64, 33, 16, 50]


@pytest.fixture(scope="module")
def ref_wav(tmp_path_factory):
    p = tmp_path_factory.mktemp("audio") / "ref.wav"
    sr = 24000
    t = np.arange(int(1.2 * sr)) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 220 * t) * 32767).astype(np.int16)
    with wave.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    return str(p)


@pytest.fixture(scope="module")
def tts():
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS

    m = FasterQwen3TTS.from_pretrained("tiny-Base", device="cuda", dtype=torch.bfloat16, max_seq_len=256, max_streams=6)
    m.predictor_graph.do_sample = False  # greedy predictor too (the reference's tests set it the same way, test_e2e_parity.py:208-215)
    yield m
    m.model.engine.close()


def _single_stream(tts, ref_wav, text, n, chunk=8):
    codes, audio = [], []
    for a, sr, info in tts.generate_voice_clone_streaming(text, "English", ref_wav, "", max_new_tokens=n, do_sample=False, chunk_size=chunk):
        audio.append(a)
    st_codes = tts.model.engine.read_codes(0, 0, n)
    return st_codes, np.concatenate(audio)


def test_requests_joining_and_leaving_match_their_single_stream_runs(tts, ref_wav):
    from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler, TTSRequest

    want = [_single_stream(tts, ref_wav, t, n) for t, n in zip(TEXTS, LENGTHS)]
    sched = BatchScheduler(tts, chunk_frames=8, codec_mode="windowed").start()
    try:
        def req(i):
            return TTSRequest(TEXTS[i], ref_audio=ref_wav, language="English", max_new_tokens=LENGTHS[i], do_sample=False)

        handles = {0: sched.submit(req(0)), 1: sched.submit(req(1))}
        first = next(iter(handles[0]))            # request 0 is mid-utterance ...
        handles[2] = sched.submit(req(2))         # ... when the next ones arrive
        handles[3] = sched.submit(req(3))
        time.sleep(0.05)
        handles[4] = sched.submit(req(4))
        handles[5] = sched.submit(req(5))
        audio = {0: [first[0]] + [a for a, _, _ in handles[0]]}
        for i in range(1, 6):
            audio[i] = [a for a, _, _ in handles[i]]
    finally:
        sched.stop()
    for i in range(6):
        h = handles[i]
        got = torch.cat(h.codes)
        assert got.shape == (LENGTHS[i], 16) and h.finish_reason == "length"
        assert torch.equal(got, want[i][0]), f"request {i}: codes differ from the single-stream run"
        a = np.concatenate(audio[i])
        assert a.shape == want[i][1].shape and np.array_equal(a, want[i][1]), f"request {i}: audio differs"
        assert h.ttfa_s is not None and h.ttfa_s > 0
    assert sched.stats["max_batch"] >= 3 and sched.stats["frames"] == sum(LENGTHS)
    # the scheduler shared launches: far fewer than one chunk launch per request and chunk
    assert sched.stats["launches"] < sum((n + 7) // 8 for n in LENGTHS)


def test_cancel_bad_request_and_policy_cohorts(tts, ref_wav):
    from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler, TTSRequest

    with BatchScheduler(tts, chunk_frames=8, codec_mode="windowed") as sched:
        long_one = sched.submit(TTSRequest(TEXTS[1], ref_audio=ref_wav, language="English", max_new_tokens=120, do_sample=False))
        bad = sched.submit(TTSRequest("x", ref_audio=ref_wav, language="Klingon", max_new_tokens=8, do_sample=False))
        victim = sched.submit(TTSRequest(TEXTS[2], ref_audio=ref_wav, language="English", max_new_tokens=200, do_sample=False))
        sampled = sched.submit(TTSRequest(TEXTS[0], ref_audio=ref_wav, language="English", max_new_tokens=16))  # another policy: next cohort
        with pytest.raises(NotImplementedError, match="Language"):
            bad.result()
        it = iter(victim)
        next(it)
        victim.cancel()
        rest = list(it)
        assert victim.finish_reason == "cancelled" and len(rest) <= 2
        a, sr = long_one.result()
        assert long_one.finish_reason == "length" and a.size == 120 * 1920 and sr == 24000
        b, _ = sampled.result()
        assert sampled.finish_reason in ("length", "stop") and b.size >= 1920 and np.isfinite(b).all()
    want, _ = _single_stream(tts, ref_wav, TEXTS[1], 120)
    assert torch.equal(torch.cat(long_one.codes), want)


def test_stateful_codec_mode_gives_the_full_decode_of_every_request(tts, ref_wav, monkeypatch):
    """codec_mode="stateful": one CodecStream per slot.  Without split-K on either side the audio of every request is the full
    non-streaming decode of its codes bit for bit — also for a slot that is reused by a later request and for an ICL request
    whose reference codes are decoded for their state first."""
    from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler, TTSRequest

    monkeypatch.setenv("FQ3C_SPLITK", "0")
    m = tts.model.model
    n_req = 9  # more than the six slots: slots are reused
    with BatchScheduler(tts, chunk_frames=8, codec_mode="stateful", codec_split_k=False) as sched:
        hs = [sched.submit(TTSRequest(TEXTS[i % 6], ref_audio=ref_wav, ref_text="reference clip says this." if i == 4 else "",
                                      xvec_only=i != 4, language="English", max_new_tokens=LENGTHS[i % 6] + i, do_sample=False))
              for i in range(n_req)]
        audio = [h.result()[0] for h in hs]
    for i, h in enumerate(hs):
        codes = torch.cat(h.codes).cuda()
        assert codes.shape[0] == LENGTHS[i % 6] + i
        if i == 4:
            (vcp, _), = [v for k, v in tts._voice_prompt_cache.items() if k[0] == ref_wav and not k[2]]
            ref = vcp["ref_code"][0].cuda()
            full = m.speech_tokenizer.decoder.decode(torch.cat([ref, codes]))[ref.shape[0] * 1920:]
        else:
            full = m.speech_tokenizer.decoder.decode(codes)
        assert audio[i].shape == (codes.shape[0] * 1920,)
        assert np.array_equal(audio[i], full.cpu().numpy()), f"request {i}: max diff {np.abs(audio[i] - full.cpu().numpy()).max()}"
