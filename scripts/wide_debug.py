"""Debug: where does a lock-step batch first differ from the single-stream runs?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("FQ3_VARIANT"): sys.path.insert(0, os.path.join(ROOT, "variants", os.environ["FQ3_VARIANT"]))  # A/B builds (scripts/build_variant.sh)
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
name, tl, pl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lens = [int(x) for x in sys.argv[5].split(",")]
cfg = make_cfg(name, tl, pl)
w = make_weights(cfg, seed=0)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.05)
sub = SubPolicy(do_sample=False)
nf = 8
for ns in [int(x) for x in sys.argv[4].split(",")]:
    eng = make_engine(cfg, w, max_streams=ns, max_seq_len=256)
    prompts = [synth_prompt(cfg, T=lens[i % len(lens)], seed=10 + i) for i in range(ns)]
    singles = []
    for tie, tam, tth, tpe in prompts:
        eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
        eng.prefill(0, tie[0].cuda(), 0, pol)
        eng.decode_frames(1, nf, pol, sub)
        singles.append(eng.read_codes(0, 0, eng.status(0).n_frames))
    for i, (tie, tam, tth, tpe) in enumerate(prompts):
        eng.set_text_conditioning(i, tth[0].cuda(), tpe.cuda())
        eng.prefill(i, tie[0].cuda(), 0, pol)
    eng.decode_frames(ns, nf, pol, sub)
    out = []
    for i in range(ns):
        got = eng.read_codes(i, 0, eng.status(i).n_frames)
        n = min(got.shape[0], singles[i].shape[0])
        d = (got[:n] != singles[i][:n]).nonzero()
        out.append("ok" if (d.numel() == 0 and got.shape == singles[i].shape) else f"f{int(d[0,0])}c{int(d[0,1])}" if d.numel() else "len")
    print(ns, "streams:", " ".join(out), flush=True)
    eng.close()
