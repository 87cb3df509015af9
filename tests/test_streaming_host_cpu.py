"""Host logic of the fused streaming generator (reference `faster_qwen3_tts/streaming.py:19-188`) against an
engine double — no GPU: one launch per chunk, chunk shapes and the reference's timing keys (`streaming.py:162-169`), only a trailing
PARTIAL chunk flagged final (`:175-188`), EOS and length cap, the reference's over-long-prompt error, and the overlapped mode's
bookkeeping (next chunk launched before the current one is yielded, full grid for the first chunk and restored at the end or when
the consumer closes the generator).  The GPU tests hold the same generator against the real engine."""
import types

import pytest
import torch

from qwen3_tts_cuda_graphs_b200 import streaming
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.predictor_graph import PredictorGraph
from qwen3_tts_cuda_graphs_b200.talker_graph import TalkerGraph

KEYS = {"chunk_index", "chunk_steps", "prefill_ms", "decode_ms", "total_steps_so_far", "is_final"}


class EngineDouble:
    def __init__(self, eos_at=10 ** 9, reduced=0, max_seq_len=64):
        self.cfg = preset("tiny-Base")
        self.device = torch.device("cpu")
        self.max_seq_len, self.max_frames = max_seq_len, 4096
        self.eos_at, self.reduced = eos_at, reduced
        self.n, self.done = 0, 0
        self.log = []       # ("launch", n_frames, grid) | ("status",) | ("grid", g) | ("prefill", T, pads)
        self.grid = 0

    def reduced_grid(self):
        return self.reduced

    def set_decode_grid(self, g):
        self.grid = g
        self.log.append(("grid", g))

    def set_text_conditioning(self, idx, tth, tpe):
        pass

    def prefill(self, idx, embeds, n_pad, policy):
        self.n, self.done = 0, 0
        self.log.append(("prefill", embeds.shape[0], n_pad))

    def decode_frames(self, n_streams, n_frames, policy, sub):
        self.log.append(("launch", n_frames, self.grid))
        for _ in range(n_frames):
            if self.n >= self.eos_at:
                self.done = 1
                break
            self.n += 1

    def status(self, idx=0):
        self.log.append(("status",))
        return types.SimpleNamespace(n_frames=self.n, done=self.done, error=0)

    def read_codes(self, idx, first, n):
        return torch.arange(first, first + n).reshape(n, 1).repeat(1, 16)


class _Event:
    def record(self, *a):
        pass


@pytest.fixture()
def no_cuda(monkeypatch):
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: _Event())


def _gen(eng, T=14, **kw):
    tg, pg = TalkerGraph(eng), PredictorGraph(eng)
    talker = types.SimpleNamespace(engine=eng)
    H = eng.cfg.talker.hidden_size
    return streaming.fast_generate_streaming(
        talker=talker, talker_input_embeds=torch.zeros(1, T, H), attention_mask=torch.ones(1, T, dtype=torch.long),
        trailing_text_hiddens=torch.zeros(1, 1, H), tts_pad_embed=torch.zeros(1, 1, H), config=eng.cfg.talker,
        predictor_graph=pg, talker_graph=tg, seed=1, **kw)


def test_chunks_keys_and_final_flag(no_cuda):
    eng = EngineDouble()
    out = list(_gen(eng, max_new_tokens=21, chunk_size=8))
    assert [c.shape for c, _ in out] == [(8, 16), (8, 16), (5, 16)] and all(c.dtype == torch.int64 for c, _ in out)
    assert torch.equal(torch.cat([c for c, _ in out])[:, 0], torch.arange(21))               # every frame once, in order
    assert all(set(info) >= KEYS for _, info in out)                                           # streaming.py:162-169
    assert [i["is_final"] for _, i in out] == [False, False, True]                             # only the trailing partial chunk
    assert [i["chunk_index"] for _, i in out] == [0, 1, 2] and [i["total_steps_so_far"] for _, i in out] == [8, 16, 21]
    assert out[0][1]["prefill_ms"] > 0 and out[1][1]["prefill_ms"] == 0 == out[2][1]["prefill_ms"]
    assert [e for e in eng.log if e[0] == "launch"] == [("launch", 8, 0), ("launch", 8, 0), ("launch", 5, 0)]  # one launch per chunk
    assert eng.log[0] == ("prefill", 14, 0)


def test_eos_inside_a_chunk_and_at_a_chunk_boundary(no_cuda):
    out = list(_gen(EngineDouble(eos_at=13), max_new_tokens=100, chunk_size=8))
    assert [c.shape[0] for c, _ in out] == [8, 5] and [i["is_final"] for _, i in out] == [False, True]
    eng = EngineDouble(eos_at=16)
    out = list(_gen(eng, max_new_tokens=100, chunk_size=8))
    # EOS right after a full chunk: the reference yields no empty trailing chunk and never sets is_final then (streaming.py:175)
    assert [c.shape[0] for c, _ in out] == [8, 8] and [i["is_final"] for _, i in out] == [False, False]
    assert len([e for e in eng.log if e[0] == "launch"]) == 3                                   # the third launch found EOS at once
    assert list(_gen(EngineDouble(eos_at=0), max_new_tokens=100, chunk_size=8)) == []           # empty generation: nothing is yielded
    out = list(_gen(EngineDouble(), max_new_tokens=16, chunk_size=8))                           # the length cap on a boundary
    assert [c.shape[0] for c, _ in out] == [8, 8]


def test_prompt_longer_than_the_cache_raises_the_reference_error(no_cuda):
    with pytest.raises(RuntimeError, match=r"Input is too long: prefill has 65 tokens but max_seq_len=64\."):
        next(_gen(EngineDouble(max_seq_len=64), T=65, max_new_tokens=8, chunk_size=8))


def test_overlapped_mode_launches_ahead_and_restores_the_grid(no_cuda, monkeypatch):
    monkeypatch.setenv("FQ3_OVERLAP_CODEC", "1")
    eng = EngineDouble(reduced=128)
    seen = []
    for chunk, info in _gen(eng, max_new_tokens=40, chunk_size=8):
        launches = [e for e in eng.log if e[0] == "launch"]
        seen.append((chunk.shape[0], len(launches), "codes_ready" in info))
    # chunk 0: its own launch only (time to first audio does not wait behind a speculative launch); from chunk 1 on the next
    # chunk is already running when the current one is handed out; the last one has nothing to launch ahead
    assert seen == [(8, 1, False), (8, 3, True), (8, 4, True), (8, 5, True), (8, 5, False)]
    launches = [e for e in eng.log if e[0] == "launch"]
    assert launches[0] == ("launch", 8, 0) and launches[1] == ("launch", 8, 0) and all(g == 128 for _, _, g in launches[2:])
    assert eng.grid == 0                                                                        # other callers get the full grid back
    assert torch.equal(torch.cat([c for c, _ in _gen(EngineDouble(reduced=128), max_new_tokens=21, chunk_size=8)])[:, 0], torch.arange(21))

    eng = EngineDouble(reduced=128)
    g = _gen(eng, max_new_tokens=400, chunk_size=8)
    next(g), next(g)                                                                            # chunk 2 is in flight now
    n_status = len([e for e in eng.log if e[0] == "status"])
    g.close()                                                                                   # benchmarks/throughput.py:63
    assert len([e for e in eng.log if e[0] == "status"]) == n_status + 1                        # the speculative chunk was waited for
    assert eng.grid == 0

    monkeypatch.setenv("FQ3_OVERLAP_CODEC", "0")
    eng = EngineDouble(reduced=128)
    out = list(_gen(eng, max_new_tokens=24, chunk_size=8))
    assert all("codes_ready" not in i for _, i in out) and all(g == 0 for _, _, g in [e for e in eng.log if e[0] == "launch"])


def test_non_streaming_generate_returns_codes_and_the_reference_timing_keys(no_cuda):
    """generate.py:205-215: (int64 [T, 16] | None, {prefill_ms, decode_s, steps, ms_per_step, steps_per_s}); one launch for the
    whole utterance, one status read."""
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate

    def run(eng, **kw):
        tg, pg = TalkerGraph(eng), PredictorGraph(eng)
        H = eng.cfg.talker.hidden_size
        return fast_generate(types.SimpleNamespace(engine=eng), torch.zeros(1, 14, H), torch.ones(1, 14, dtype=torch.long),
                             torch.zeros(1, 1, H), torch.zeros(1, 1, H), eng.cfg.talker, pg, tg, seed=3, **kw)

    eng = EngineDouble(eos_at=37)
    codes, timing = run(eng, max_new_tokens=2048)
    assert codes.shape == (37, 16) and codes.dtype == torch.int64 and torch.equal(codes[:, 0], torch.arange(37))
    assert set(timing) == {"prefill_ms", "decode_s", "steps", "ms_per_step", "steps_per_s"} and timing["steps"] == 37
    assert [e for e in eng.log if e[0] == "launch"] == [("launch", 2048, 0)] and len([e for e in eng.log if e[0] == "status"]) == 1
    codes, timing = run(EngineDouble(eos_at=0), max_new_tokens=64)
    assert codes is None and timing["steps"] == 0 and timing["ms_per_step"] == 0                # generate.py:213-215
    eng = EngineDouble()
    eng.max_frames = 100
    codes, _ = run(eng, max_new_tokens=2048)
    assert codes.shape[0] == 100                                                                 # the engine's frame store bounds the launch
    with pytest.raises(NotImplementedError):
        run(EngineDouble(), parity_mode=True)


def test_batch_generate_groups_requests_by_what_one_launch_takes(no_cuda):
    """fast_generate_batch: one launch per engine-full of requests with the wide program (lock-step groups of 16 formed on the
    device), groups of four otherwise (1.7B dims); every request gets its own codes back, a request that produced nothing None."""
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate_batch

    class Batched(EngineDouble):
        def __init__(self, max_streams, lockstep_group):
            super().__init__()
            self.max_streams, self.lockstep_group = max_streams, lockstep_group
            self.frames = {}

        def prefill(self, idx, embeds, n_pad, policy):
            self.frames[idx] = -int(embeds[0, 0])       # the request's length rides in its first embedding value
            self.log.append(("prefill", idx, n_pad))

        def decode_frames(self, n_streams, n_frames, policy, sub):
            self.log.append(("launch", n_streams, n_frames))
            for s in range(n_streams):
                self.frames[s] = min(-self.frames[s], n_frames)

        def status(self, idx=0):
            return types.SimpleNamespace(n_frames=self.frames[idx], done=0, error=0)

        def read_codes(self, idx, first, n):
            return torch.full((n, 16), idx, dtype=torch.int64)

    def reqs(lengths, pads=0):
        H = 8
        out = []
        for n in lengths:
            tie = torch.zeros(1, 6, H)
            tie[0, 0, 0] = n
            tam = torch.ones(1, 6, dtype=torch.long)
            tam[0, :pads] = 0
            out.append((tie, tam, torch.zeros(1, 1, H), torch.zeros(1, 1, H)))
        return out

    lengths = [5, 0, 9, 3, 7, 2]
    for max_streams, group, want_launches in ((16, 16, [6]), (4, 16, [4, 2]), (16, 4, [4, 2])):
        eng = Batched(max_streams, group)
        tg, pg = TalkerGraph(eng), PredictorGraph(eng)
        codes, timing = fast_generate_batch(tg, pg, reqs(lengths, pads=2), max_new_tokens=50, seed=5)
        assert [e[1] for e in eng.log if e[0] == "launch"] == want_launches
        assert [None if c is None else c.shape[0] for c in codes] == [5, None, 9, 3, 7, 2]
        assert timing["frames"] == sum(lengths) and {"total_s", "frames", "audio_s_per_s"} <= set(timing)
        assert all(e[2] == 2 for e in eng.log if e[0] == "prefill")                             # left pads counted from the mask
    with pytest.raises(RuntimeError, match="Input is too long"):
        eng = Batched(4, 16)
        eng.max_seq_len = 5
        fast_generate_batch(TalkerGraph(eng), PredictorGraph(eng), reqs([3]))
