"""The same-box GPU anchor (oracle/graph_anchor.py: the reference's StaticCache + CUDA-graph style on plain torch ops) must
compute what the oracle computes, so that its timing in bench.py is the timing of the real work."""
import pytest
import torch

from helpers import make_cfg, make_oracle, make_weights, rel_err, synth_prompt

pytestmark = pytest.mark.gpu


def test_graph_anchor_matches_oracle():
    from oracle.graph_anchor import GraphAnchor
    cfg = make_cfg("0.6B-Base", 2, 2)
    w = make_weights(cfg, seed=3)
    orc = make_oracle(cfg, w)
    orc.sub.do_sample = False
    anc = GraphAnchor(cfg, w, max_seq_len=64, device="cuda")
    anc.capture()
    tie, tam, tth, tpe = synth_prompt(cfg, T=14)
    tok, ph, T = anc.prefill(tie)
    ref_logits, ref_ph, _ = orc.talker_prefill(tie, tam)
    assert T == 14
    assert rel_err(ph, ref_ph) <= 3e-2
    pad = tpe.cuda().view(1, 1, -1)
    frames = []
    for i in range(3):
        codes, tok, ph = anc.frame(tok, ph, T + i, pad)
        frames.append(codes.cpu())
    frames = torch.stack(frames)
    # teacher-force the oracle with the anchor's ids: every id within tolerance of the oracle's maximum
    trace = {}
    list(orc.generate_frames(tie, tam, tpe, tpe, max_new_tokens=3, min_new_tokens=0, do_sample=False, repetition_penalty=1.0,
                             max_seq_len=64, trace=trace, forced=frames))
    for i in range(3):
        pl = trace["pred_logits"][i]
        for c in range(orc.ncb):
            assert float(pl[c].max() - pl[c][int(frames[i, c + 1])]) <= 3e-2 * float(pl.abs().max()), (i, c)
    # a replayed frame equals the eager run of the same body
    tok2, ph2, _ = anc.prefill(tie)
    c_graph, n_graph, h_graph = anc.frame(tok2, ph2, T, pad, use_graphs=True)
    tok3, ph3, _ = anc.prefill(tie)
    c_eager, n_eager, h_eager = anc.frame(tok3, ph3, T, pad, use_graphs=False)
    assert torch.equal(c_graph, c_eager) and torch.equal(h_graph, h_eager)
    assert anc.time_frames(tie, tpe, 4) > 0
