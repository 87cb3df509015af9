"""Prefill time: chunked decode-kernel path vs dense tensor-core path, for a few prompt lengths."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy
name = sys.argv[1] if len(sys.argv) > 1 else "0.6B-Base"
cfg = make_cfg(name)
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=64)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.0)
ev = lambda: torch.cuda.Event(enable_timing=True)
for T in [int(t) for t in os.environ.get("PF_T", "39,120,240,480").split(",")]:
    tie, tam, tth, tpe = synth_prompt(cfg, T=T)
    x = tie[0].cuda()
    res = {}
    for dense in (False, True):
        for _ in range(3): eng.prefill(0, x, 0, pol, dense=dense)
        ts = []
        for _ in range(5):
            a, b = ev(), ev(); torch.cuda.synchronize(); a.record(); eng.prefill(0, x, 0, pol, dense=dense); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[dense] = sorted(ts)[2]
    print(f"{name} T={T:4d}: chunked {res[False]:7.3f} ms   dense {res[True]:7.3f} ms   ({res[False]/res[True]:.1f}x)")
