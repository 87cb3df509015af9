"""CPU suite: pins the oracle (test infrastructure) to the reference's own outputs and to transformers."""
import os

import pytest
import torch

from helpers import make_cfg, make_oracle, make_weights, synth_prompt

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sgold():
    return torch.load(os.path.join(GOLD, "sampling_golden.pt"), weights_only=False)


def test_repetition_penalty_known_answer(sgold):
    """tests/test_sampling.py:10-21 of the reference, and the reference's own output for it."""
    from oracle.qwen3_tts_oracle import apply_repetition_penalty
    out = apply_repetition_penalty(sgold["kat_logits"].clone(), sgold["kat_history"], 1.1)
    assert pytest.approx(out[0, 0, 7].item(), rel=1e-6) == 1.0 / 1.1
    assert pytest.approx(out[0, 0, 8].item(), rel=1e-6) == -1.0 * 1.1
    assert torch.equal(out, sgold["kat_out"])


def test_sampling_matches_reference_vectors(sgold):
    from oracle.qwen3_tts_oracle import apply_repetition_penalty, candidate_set, sample_logits
    for c in sgold["cases"]:
        logits = c["logits"]
        pen = apply_repetition_penalty(logits.clone().unsqueeze(0), c["history"], 1.05)[0]
        assert torch.equal(pen, c["penalised"])
        greedy = sample_logits(logits, temperature=0.9, top_k=50, top_p=1.0, do_sample=False,
                               suppress_mask=c["smask"], suppress_tokens=[c["eos"]])
        assert torch.equal(greedy, c["greedy"])
        for (k, p), drawn in c["drawn"].items():
            cand = candidate_set(logits, temperature=0.9, top_k=k, top_p=p, suppress_mask=c["smask"])[0]
            # everything the reference drew must be a candidate; with 400 draws the head of the set is covered
            assert not (drawn & ~cand).any(), (c["V"], k, p)
            assert drawn.sum() >= 1


def test_oracle_stack_equals_transformers_golden():
    from oracle.qwen3_tts_oracle import OracleStack
    from qwen3_tts_cuda_graphs_b200.weights import init_synthetic
    gold = torch.load(os.path.join(GOLD, "qwen3_stack_golden.pt"), weights_only=False)
    cfg = make_cfg("tiny")
    w32 = init_synthetic(cfg, seed=3, norm_jitter=0.1, dtype=torch.float32, skip_text_embedding=True)
    for dt in (torch.float32, torch.bfloat16):
        g = gold[str(dt)]
        st = OracleStack(cfg.talker, {k: v.to(dt) for k, v in w32.items()}, "talker.model", attn="eager")
        o = st.forward(g["x"], torch.arange(7))
        o2 = st.forward(g["x2"], torch.tensor([7]))
        assert torch.equal(o, g["prefill"]), dt
        assert torch.equal(o2, g["step"]), dt


def test_oracle_stack_equals_transformers_live():
    """Same check run live against the installed transformers (the image ships it on the GPU box too)."""
    transformers = pytest.importorskip("transformers")
    from oracle.qwen3_tts_oracle import OracleStack
    from qwen3_tts_cuda_graphs_b200.weights import init_synthetic
    cfg = make_cfg("tiny")
    t = cfg.predictor
    w = init_synthetic(cfg, seed=5, norm_jitter=0.1, dtype=torch.float32, skip_text_embedding=True)
    hc = transformers.Qwen3Config(
        vocab_size=32, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
        num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
        num_key_value_heads=t.num_key_value_heads, head_dim=t.head_dim, rms_norm_eps=t.rms_norm_eps,
        rope_theta=t.rope_theta, attention_bias=False, max_position_embeddings=4096)
    hc._attn_implementation = "eager"
    m = transformers.Qwen3Model(hc).eval()
    pre = "talker.code_predictor.model."
    m.load_state_dict({k[len(pre):]: v for k, v in w.items() if k.startswith(pre + "layers") or k == pre + "norm.weight"},
                      strict=False)
    x = torch.randn(1, 2, t.hidden_size, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(inputs_embeds=x).last_hidden_state
    st = OracleStack(t, w, "talker.code_predictor.model")
    assert torch.equal(st.forward(x, torch.arange(2)), ref)


def test_loop_semantics_eos_and_min_new_tokens():
    """tests/test_sampling.py:24-118 restated: EOS is greedy-best from the start, min_new_tokens delays it."""
    from dataclasses import replace
    cfg = make_cfg("tiny")
    cfg = replace(cfg, talker=replace(cfg.talker, codec_eos_token_id=0))
    w = make_weights(cfg, seed=2)
    w["talker.codec_head.weight"].zero_()
    orc = make_oracle(cfg, w)
    orc.sub.do_sample = False
    tie, tam, tth, tpe = synth_prompt(cfg, T=3)
    for mn, expect in [(2, 2), (5, 5)]:
        codes, timing = orc.fast_generate(tie, tam, tth, tpe, max_new_tokens=16, min_new_tokens=mn, do_sample=False,
                                          repetition_penalty=1.0)
        assert codes.shape == (expect, 16) and (codes[:, 0] != 0).all() and timing["steps"] == expect
    codes, _ = orc.fast_generate(tie, tam, tth, tpe, max_new_tokens=16, min_new_tokens=0, do_sample=False, repetition_penalty=1.0)
    assert codes is None


def test_oracle_streaming_equals_nonstreaming():
    """tests/test_e2e_parity.py:726-780 on the oracle itself."""
    cfg = make_cfg("tiny")
    w = make_weights(cfg, seed=0)
    orc = make_oracle(cfg, w)
    orc.sub.do_sample = False
    tie, tam, tth, tpe = synth_prompt(cfg, T=14, R=2)
    kw = dict(max_new_tokens=20, do_sample=False, repetition_penalty=1.05)
    full, _ = orc.fast_generate(tie, tam, tth, tpe, **kw)
    chunks = list(orc.fast_generate_streaming(tie, tam, tth, tpe, chunk_size=8, **kw))
    assert [c.shape[0] for c, _ in chunks] == [8, 8, 4]
    assert chunks[-1][1]["is_final"] and not chunks[0][1]["is_final"]
    assert torch.equal(torch.cat([c for c, _ in chunks]), full)
    assert full.shape == (20, 16) and (full[:, 0] < cfg.talker.vocab_size - 1024).all()


def test_teacher_forcing_reproduces_free_run():
    cfg = make_cfg("tiny")
    w = make_weights(cfg, seed=0)
    orc = make_oracle(cfg, w)
    orc.sub.do_sample = False
    tie, tam, tth, tpe = synth_prompt(cfg, T=9)
    kw = dict(max_new_tokens=6, do_sample=False, repetition_penalty=1.05)
    free = torch.stack(list(orc.generate_frames(tie, tam, tth, tpe, **kw)))
    tr = {}
    forced = torch.stack(list(orc.generate_frames(tie, tam, tth, tpe, trace=tr, forced=free, **kw)))
    assert torch.equal(free, forced)
    for i in range(5):
        assert int(tr["talker_final"][i].argmax()) == int(free[i + 1, 0])
