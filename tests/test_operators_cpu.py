"""Host logic of the operator seam (reference `talker_graph.py:149-214`, `predictor_graph.py:34-214`) against an
engine double — no GPU: what `TalkerGraph` / `PredictorGraph` hand to the C ABI for the arguments the reference's callers pass
(`generate.py:137,140,156,179`), the reference's error text for an over-long prefill, the static output buffer, the mutable
sampler attributes, the single-stream guard."""
import types

import pytest
import torch

from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.predictor_graph import PredictorGraph
from qwen3_tts_cuda_graphs_b200.talker_graph import TalkerGraph


class EngineDouble:
    def __init__(self, max_seq_len=32):
        self.cfg = preset("tiny-Base")
        self.device = torch.device("cpu")
        self.max_seq_len = max_seq_len
        self.calls = []

    def reset_stream(self, idx):
        self.calls.append(("reset", idx))

    def import_kv(self, idx, layer, k, v):
        self.calls.append(("kv", idx, layer, tuple(k.shape), tuple(v.shape)))

    def set_generation_state(self, idx, n_pad, delta):
        self.calls.append(("state", idx, n_pad, delta))

    def talker_step(self, idx, embeds, position, want_logits=True):
        self.calls.append(("step", idx, tuple(embeds.shape), position))
        H = self.cfg.talker.hidden_size
        return torch.full((H,), float(position), dtype=torch.bfloat16), torch.zeros(self.cfg.talker.vocab_size)

    def predictor_run(self, idx, pred_input, sub, seed=0, want_logits=False):
        self.calls.append(("pred", idx, tuple(pred_input.shape), (sub.do_sample, sub.top_k, sub.top_p, sub.temperature), seed))
        return torch.arange(self.cfg.predictor.num_codebooks), None


def test_talker_graph_prefill_kv_and_generation_state():
    eng = EngineDouble(max_seq_len=32)
    tg = TalkerGraph(eng)
    t = eng.cfg.talker
    assert tg.max_seq_len == 32 and tg.output_buf.shape == (1, 1, t.hidden_size)
    kv = [(torch.zeros(1, t.num_key_value_heads, 9, t.head_dim), torch.zeros(1, t.num_key_value_heads, 9, t.head_dim))
          for _ in range(t.num_hidden_layers)]
    assert tg.prefill_kv(kv) == 9                                              # talker_graph.py:153-170
    assert eng.calls[0] == ("reset", 0) and [c[2] for c in eng.calls[1:]] == list(range(t.num_hidden_layers))
    assert eng.calls[1][3] == (t.num_key_value_heads, 9, t.head_dim)
    too_long = [(torch.zeros(1, t.num_key_value_heads, 33, t.head_dim),) * 2 for _ in range(t.num_hidden_layers)]
    with pytest.raises(RuntimeError, match=r"Input is too long: prefill has 33 tokens but max_seq_len=32\. Use shorter text or shorter reference audio\."):
        tg.prefill_kv(too_long)                                                # talker_graph.py:163-167, same text
    eng.calls.clear()
    mask = torch.tensor([[0, 0, 0, 1, 1, 1, 1]])
    tg.set_generation_state(mask, torch.tensor([[-3.0]]))                      # talker_graph.py:172-196: left-padded batch row
    tg.set_generation_state(None, None)                                        # bs = 1, no pads: delta 0
    tg.set_generation_state(torch.ones(1, 5, dtype=torch.long), torch.tensor([2.0]))   # 1-D rope_deltas (talker_graph.py:193-194)
    assert eng.calls == [("state", 0, 3, -3), ("state", 0, 0, 0), ("state", 0, 0, 2)]
    prefilled = types.SimpleNamespace(engine=eng, length=40)                   # KV already written in place by the engine's prefill
    with pytest.raises(RuntimeError, match="prefill has 40 tokens but max_seq_len=32"):
        tg.prefill_kv(prefilled)
    prefilled.length = 17
    assert tg.prefill_kv(prefilled) == 17


def test_talker_graph_run_returns_the_static_buffer():
    eng = EngineDouble()
    tg = TalkerGraph(eng)
    H = eng.cfg.talker.hidden_size
    a = tg.run(torch.zeros(1, 1, H, dtype=torch.bfloat16), position=11)
    assert a is tg.output_buf and a.shape == (1, 1, H) and float(a[0, 0, 0]) == 11.0
    kept = a.clone()
    b = tg.run(torch.zeros(1, 1, H, dtype=torch.bfloat16), position=12)
    assert b is a and float(a[0, 0, 0]) == 12.0 and float(kept[0, 0, 0]) == 11.0   # "use immediately or clone" (talker_graph.py:214)
    assert tg.last_logits.shape == (eng.cfg.talker.vocab_size,)
    tg.capture(prefill_len=100, num_warmup=3)
    tg.reset(prefill_len=0)
    assert tg.captured and eng.calls[-1] == ("reset", 0)


def test_predictor_graph_reads_its_sampler_attributes_at_every_run():
    eng = EngineDouble()
    pg = PredictorGraph(eng, do_sample=True, top_k=50, temperature=0.9)        # model.py:124-133
    H = eng.cfg.talker.hidden_size
    x = torch.zeros(1, 2, H, dtype=torch.bfloat16)
    codes = pg.run(x)
    assert codes.shape == (eng.cfg.predictor.num_codebooks,) and pg.max_seq == 2 + eng.cfg.predictor.num_codebooks
    pg.do_sample, pg.top_k, pg.top_p, pg.temperature = False, 20, 0.8, 0.5     # tests/test_e2e_parity.py:211-214 mutates them
    pg.run(x)
    (_, _, shape, pol1, seed1), (_, _, _, pol2, seed2) = eng.calls
    assert shape == (1, 2, H) and pol1 == (True, 50, 1.0, 0.9) and pol2 == (False, 20, 0.8, 0.5)
    assert seed1 != seed2                                                      # a fresh draw per call, like torch.multinomial
    pg.capture(num_warmup=3)
    assert pg.captured


def test_the_operator_seam_is_single_stream_like_the_reference():
    eng = EngineDouble()
    with pytest.raises(ValueError):
        TalkerGraph(eng, stream_idx=1)                                         # talker_graph.py:46-47: bs = 1 buffers
    with pytest.raises(ValueError):
        PredictorGraph(eng, stream_idx=2)                                      # predictor_graph.py:70-71


def test_sampling_wrappers_have_no_cpu_fallback_and_pass_the_reference_arguments_on():
    """`sampling.py:10-66` signatures: without an engine both entry points refuse (no CPU arithmetic behind them); with one,
    suppression is applied before the engine call (mask and id list, sampling.py:41-47), the policy carries temperature / top-k /
    top-p / do_sample, the penalty is a no-op for 1.0 or an empty history (sampling.py:17-18)."""
    from qwen3_tts_cuda_graphs_b200 import sampling

    sampling.set_default_engine(None)
    logits = torch.tensor([[0.5, 2.0, -1.0, 3.0]])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sampling.sample_logits(logits, temperature=0.9, top_k=50, top_p=1.0, do_sample=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sampling.apply_repetition_penalty(logits.clone().unsqueeze(0), torch.tensor([1]), 1.1)
    same = logits.clone()
    assert sampling.apply_repetition_penalty(same, torch.tensor([1]), 1.0) is same          # penalty 1.0: untouched, no engine needed
    assert sampling.apply_repetition_penalty(same, torch.tensor([], dtype=torch.long), 1.3) is same

    seen = {}

    class Eng:
        def sample(self, x, history, pol, eos_id, suppress_eos, draw_index):
            seen.update(x=x.clone(), history=history, pol=pol, eos_id=eos_id, suppress_eos=suppress_eos, draw_index=draw_index)
            return torch.tensor([int(torch.argmax(x))])

        def apply_repetition_penalty(self, lg, hist, p):
            seen.update(rep=(tuple(lg.shape), hist.tolist(), p))
            return lg

    mask = torch.tensor([False, False, False, True])
    tok = sampling.sample_logits(logits, temperature=0.7, top_k=5, top_p=0.9, do_sample=False, suppress_mask=mask, suppress_tokens=[1],
                                 engine=Eng(), seed=11, draw_index=4)
    assert tok.tolist() == [0]                                                               # ids 3 (mask) and 1 (list) were suppressed
    assert seen["x"].tolist() == [0.5, float("-inf"), -1.0, float("-inf")] and logits[0, 1] == 2.0   # on a clone (sampling.py:40)
    p = seen["pol"]
    assert (p.do_sample, p.top_k, p.top_p, p.temperature, p.seed, p.repetition_penalty) == (False, 5, 0.9, 0.7, 11, 1.0)
    assert seen["draw_index"] == 4 and seen["history"] is None and seen["suppress_eos"] is False
    sampling.set_default_engine(Eng())
    try:
        sampling.apply_repetition_penalty(logits.clone().unsqueeze(0), torch.tensor([1, 1, 3]), 1.1)
        assert seen["rep"] == ((1, 1, 4), [1, 1, 3], 1.1)
    finally:
        sampling.set_default_engine(None)
