"""Minimal driver for ncu: 16 lock-step streams through the wide frame program (0.6B dims), a few 4-frame launches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=512, max_streams=ns, max_frames=64)
pol = SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
for s in range(ns):
    tie, tam, tth, tpe = synth_prompt(cfg, T=39, seed=1 + s)
    eng.set_text_conditioning(s, tth[0].cuda(), tpe.cuda())
    eng.prefill(s, tie[0].cuda(), 0, pol)
for _ in range(3):
    eng.decode_frames(ns, 4, pol, sub)
torch.cuda.synchronize()
print("ok", [eng.status(s).n_frames for s in range(ns)][:4])
