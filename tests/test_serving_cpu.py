"""Host logic of the continuous-batching scheduler (serving.BatchScheduler) against doubles of the engine
and the model — no GPU: admission at chunk boundaries, one launch for all running utterances, EOS / length / cancel / bad-request
retirement, policy cohorts, slot reuse only after the last chunk went through the codec, audio order.  The GPU tests
(tests/test_serving_gpu.py) hold the same loop against the real engine bit for bit."""
import threading
import time
import types

import numpy as np
import pytest
import torch

from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler, TTSRequest

pytestmark = pytest.mark.timeout(120)
SPF = 1920


class FakeEngine:
    """Streams keyed by slot; a frame's codes row is [request key, frame index, 0, ...]; EOS after `eos_at` frames."""

    def __init__(self, max_streams=4, frame_sleep=0.0):
        self.device = torch.device("cpu")
        self.max_streams, self.lockstep_group, self.max_seq_len, self.max_frames = max_streams, 16, 64, 4096
        self.st = {s: dict(n=0, done=3, key=-1, eos_at=10 ** 9, codes=[]) for s in range(max_streams)}
        self.launches = []   # (policy tuple, [live slots], hi)
        self.events = []     # ("prefill" | "retire", slot)
        self.frame_sleep = frame_sleep

    def retire_stream(self, s):
        self.st[s]["done"] = self.st[s]["done"] or 3
        self.events.append(("retire", s))

    def set_text_conditioning(self, s, tth, tpe):
        pass

    def prefill(self, s, embeds, n_pad, policy):
        key, eos_at = int(embeds[0, 0]), int(embeds[0, 1])
        self.st[s] = dict(n=0, done=0, key=key, eos_at=eos_at, codes=[])
        self.events.append(("prefill", s))

    def decode_frames(self, hi, n, policy, sub):
        live = [s for s in range(hi) if self.st[s]["done"] == 0]
        self.launches.append(((policy.do_sample, policy.top_k, policy.temperature), live, hi))
        for s in live:
            st = self.st[s]
            for _ in range(n):
                if st["n"] >= st["eos_at"]:
                    st["done"] = 1  # EOS is tested before the frame is appended (generate.py:150)
                    break
                st["codes"].append([st["key"], st["n"]] + [0] * 14)
                st["n"] += 1
        if self.frame_sleep:
            time.sleep(self.frame_sleep * n)

    def status(self, s):
        st = self.st[s]
        return types.SimpleNamespace(n_frames=st["n"], done=st["done"] if st["done"] in (1, 2) else 0, error=0)

    def read_codes(self, s, first, n):
        return torch.tensor(self.st[s]["codes"][first:first + n], dtype=torch.int64)


class FakeDecoder:
    """decode(codes) -> every frame becomes 1920 samples holding its frame index; counts the frames it was asked for."""

    def __init__(self):
        self.cfg = types.SimpleNamespace(trans_conv_trim="right", sample_rate=24000)
        self.device = torch.device("cpu")
        self.frames_decoded = 0
        self._plans = {}
        self.stream_log = []

    def decode(self, codes, skip_samples=0):
        self.frames_decoded += int(codes.shape[0])
        return codes[:, 1].float().repeat_interleave(SPF)

    def open_stream(self, max_chunk_frames, split_k=None):
        outer = self

        class S:
            def reset(self):
                outer.stream_log.append(("reset", id(self)))

            def decode(self, codes):
                outer.stream_log.append(("decode", id(self), int(codes[0, 0]), int(codes[0, 1])))
                return codes[:, 1].float().repeat_interleave(SPF)

            def close(self):
                pass
        return S()


class FakeTTS:
    def __init__(self, engine):
        from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer

        self.decoder = FakeDecoder()
        tok = SpeechTokenizer(self.decoder)
        inner = types.SimpleNamespace(speech_tokenizer=tok, tts_model_type="base", tts_model_size="0b6")
        self.model = types.SimpleNamespace(engine=engine, model=inner)
        self.sample_rate = 24000
        self.predictor_graph = types.SimpleNamespace(policy=lambda: None)

    def _prepare_generation(self, text, ref_audio, ref_text, language="Auto", **kw):
        if language == "Klingon":
            raise NotImplementedError("Language Klingon not implemented")
        key, eos_at = (int(x) for x in text.split(","))
        tie = torch.zeros(1, 5, 8)
        tie[0, 0, 0], tie[0, 0, 1] = key, eos_at
        return self.model.model, None, None, tie, torch.ones(1, 5, dtype=torch.long), torch.zeros(1, 1, 8), torch.zeros(1, 1, 8), None

    @staticmethod
    def _to_numpy(a):
        return a.flatten().float().cpu().numpy()


def _req(key, frames, eos_at=10 ** 9, **kw):
    return TTSRequest(f"{key},{eos_at}", ref_audio="v.wav", language="English", max_new_tokens=frames, do_sample=False, **kw)


def _audio_ok(h, n):
    a, sr = h.result()
    assert sr == 24000 and a.shape == (n * SPF,)
    assert np.array_equal(a, np.repeat(np.arange(n, dtype=np.float32), SPF))  # every frame once, in order
    codes = torch.cat(h.codes)
    assert codes.shape == (n, 16) and torch.equal(codes[:, 1], torch.arange(n)) and int(codes[0, 0]) == h.id_key


@pytest.mark.parametrize("mode,overlap", [("windowed", "auto"), ("windowed", "host"), ("windowed", "gpu"), ("windowed", "none"), ("stateful", "host")])
def test_more_requests_than_slots_all_complete_in_shared_launches(mode, overlap):
    eng = FakeEngine(max_streams=4)
    tts = FakeTTS(eng)
    lengths = [20, 9, 33, 16, 8, 41, 5, 24, 12]
    with BatchScheduler(tts, chunk_frames=8, codec_mode=mode, overlap_codec=overlap) as sched:
        hs = []
        for i, n in enumerate(lengths):
            h = sched.submit(_req(100 + i, n))
            h.id_key = 100 + i
            hs.append(h)
        for h, n in zip(hs, lengths):
            _audio_ok(h, n)
            assert h.finish_reason == "length" and h.ttfa_s is not None
    assert max(len(live) for _, live, _ in eng.launches) == 4                       # the slots were used ...
    assert sum(len(live) for _, live, _ in eng.launches) < 2 * sum((n + 7) // 8 for n in lengths)
    assert len(eng.launches) < sum((n + 7) // 8 for n in lengths)                  # ... by shared launches
    assert all(hi <= 4 for _, _, hi in eng.launches)
    if mode == "stateful":
        # a slot's codec stream is reset only between utterances: never between two chunks of one request
        last = {}
        for ev in tts.decoder.stream_log:
            if ev[0] == "decode":
                _, sid, key, first = ev
                if sid in last and last[sid][0] == key:
                    assert first == last[sid][1] + 8, "a chunk of another decode order slipped in"
                elif first != 0:
                    raise AssertionError(f"stream {sid} started request {key} at frame {first} without a reset")
                last[sid] = (key, first)
            else:
                last.pop(ev[1], None)


def test_eos_cancel_error_and_late_arrivals():
    eng = FakeEngine(max_streams=3, frame_sleep=0.002)
    tts = FakeTTS(eng)
    with BatchScheduler(tts, chunk_frames=8) as sched:
        long_one = sched.submit(_req(1, 200))
        eos = sched.submit(_req(2, 200, eos_at=13))
        bad = sched.submit(TTSRequest("3,99", ref_audio="v.wav", language="Klingon", max_new_tokens=8))
        victim = sched.submit(_req(4, 10 ** 4))
        it = iter(victim)
        next(it)
        victim.cancel()
        assert len(list(it)) <= 3 and victim.finish_reason == "cancelled"  # stops at the next chunk boundary
        with pytest.raises(NotImplementedError):
            bad.result()
        eos.id_key, long_one.id_key = 2, 1
        _audio_ok(eos, 13)
        assert eos.finish_reason == "stop"
        late = sched.submit(_req(5, 17))  # joins while request 1 is mid-utterance
        late.id_key = 5
        _audio_ok(late, 17)
        _audio_ok(long_one, 200)
    joined = [live for _, live, _ in eng.launches if len(live) >= 2]
    assert joined, "the late request never shared a launch with the long one"
    assert ("retire", 0) in eng.events or ("retire", 1) in eng.events


def test_policy_cohorts_never_share_a_launch_and_empty_generation():
    eng = FakeEngine(max_streams=4)
    tts = FakeTTS(eng)
    with BatchScheduler(tts, chunk_frames=8) as sched:
        greedy = [sched.submit(_req(10 + i, 24)) for i in range(3)]
        sampled = [sched.submit(TTSRequest(f"{20 + i},{10 ** 9}", ref_audio="v.wav", language="English", max_new_tokens=16, temperature=0.7))
                   for i in range(2)]
        empty = sched.submit(_req(30, 8, eos_at=0))
        for i, h in enumerate(greedy):
            h.id_key = 10 + i
            _audio_ok(h, 24)
        for i, h in enumerate(sampled):
            h.id_key = 20 + i
            _audio_ok(h, 16)
        a, _ = empty.result()
        assert a.shape == (1,) and a[0] == 0.0 and empty.finish_reason == "stop"  # model.py:630-632
    pols = {}
    for pol, live, _ in eng.launches:
        for s in live:
            pols.setdefault(len(pols), None)
        assert pol in ((False, 50, 0.9), (True, 50, 0.7))
    kinds = [pol for pol, live, _ in eng.launches if live]
    switches = sum(1 for a, b in zip(kinds, kinds[1:]) if a != b)
    assert switches <= 2  # cohort after cohort, not interleaved per launch


def test_emit_every_hands_out_the_first_chunk_at_once_then_pairs():
    eng = FakeEngine(max_streams=2)
    tts = FakeTTS(eng)
    with BatchScheduler(tts, chunk_frames=8, emit_every=2) as sched:
        h = sched.submit(_req(7, 45))
        h.id_key = 7
        sizes = [len(a) // SPF for a, _, _ in h]
    assert sizes[0] == 8 and sum(sizes) == 45 and all(s in (16, 13, 5) or s == 8 for s in sizes[1:]) and len(sizes) <= 4


def test_a_waiting_policy_is_not_starved_and_stop_releases_everyone():
    eng = FakeEngine(max_streams=2, frame_sleep=0.001)
    tts = FakeTTS(eng)
    sched = BatchScheduler(tts, chunk_frames=8).start()
    first = sched.submit(_req(1, 64))
    other = sched.submit(TTSRequest(f"2,{10 ** 9}", ref_audio="v.wav", language="English", max_new_tokens=16, temperature=0.5))
    behind = [sched.submit(_req(10 + i, 16)) for i in range(3)]  # same policy as `first`, but they arrived after `other`
    other.id_key = 2
    _audio_ok(other, 16)
    first_other = [pol for pol, _, _ in eng.launches].index((True, 50, 0.5))
    # the cohort of `first` drained before `other` ran, and none of the later same-policy requests was admitted before it
    assert all(len(live) <= 1 for pol, live, _ in eng.launches[:first_other])
    endless = sched.submit(_req(99, 10 ** 6))
    it = iter(endless)
    next(it)
    sched.stop()
    list(it)  # returns: the scheduler released the request on shutdown
    assert endless.finish_reason in ("shutdown", "length")
    for i, h in enumerate(behind):
        assert h.finish_reason in ("length", "shutdown")


@pytest.mark.parametrize("where", ["status", "launch"])
def test_a_device_fault_reaches_every_waiter_and_later_submits_are_refused(where):
    """The loop dies of a device fault (engine.status().error != 0): running requests, the request whose last chunk was still
    waiting for the codec, requests waiting for a slot or for another policy's cohort — all get the exception, none hangs;
    submit() refuses afterwards and `healthy` turns False (server.Dispatcher routes around the replica)."""

    class FaultyEngine(FakeEngine):
        fault_after = 3

        def status(self, s):
            st = super().status(s)
            if where == "status" and len(self.launches) >= self.fault_after:
                st.error = 7
            return st

        def decode_frames(self, hi, n, policy, sub):
            if where == "launch" and len(self.launches) + 1 >= self.fault_after:  # fq3_decode_frames -> FQ3_E_DEVICE_FAULT
                raise RuntimeError("device fault 7 at launch")
            super().decode_frames(hi, n, policy, sub)

    eng = FaultyEngine(max_streams=2, frame_sleep=0.001)
    tts = FakeTTS(eng)
    sched = BatchScheduler(tts, chunk_frames=8, overlap_codec="host").start()
    running = sched.submit(_req(1, 10 ** 4))
    ending = sched.submit(_req(2, 16))            # retired after launch 2; its final chunk waits for the codec beside launch 3
    queued = sched.submit(_req(3, 16))            # no free slot yet
    other = sched.submit(TTSRequest(f"4,{10 ** 9}", ref_audio="v.wav", language="English", max_new_tokens=16, temperature=0.5))
    for h in (running, ending, queued, other):
        if where == "status" and h is ending:  # its last chunk went out beside launch 3, before the fault was read
            assert len(h.result()[0]) == 16 * SPF and h.finish_reason == "length"
            continue
        with pytest.raises(RuntimeError, match="device fault 7"):
            h.result()
        assert h.finish_reason == "error"
    assert sched.healthy is False
    with pytest.raises(RuntimeError, match="scheduler stopped"):
        sched.submit(_req(5, 8))
    sched.stop()


def test_stop_then_start_serves_again():
    eng = FakeEngine(max_streams=2)
    tts = FakeTTS(eng)
    sched = BatchScheduler(tts, chunk_frames=8).start()
    a = sched.submit(_req(1, 12))
    a.id_key = 1
    _audio_ok(a, 12)
    sched.stop()
    sched.start()
    b = sched.submit(_req(2, 9))
    b.id_key = 2
    _audio_ok(b, 9)
    sched.stop()
    assert sched.healthy


def test_result_timeout_raises_and_cancels():
    eng = FakeEngine(max_streams=1, frame_sleep=0.002)
    tts = FakeTTS(eng)
    with BatchScheduler(tts, chunk_frames=8) as sched:
        slow = sched.submit(_req(1, 10 ** 5))
        with pytest.raises(TimeoutError):
            slow.result(timeout=0.2)
        assert slow.cancelled
        nxt = sched.submit(_req(2, 8))  # the only slot is free again at the next chunk boundary
        nxt.id_key = 2
        _audio_ok(nxt, 8)
