"""Minimal device-loop driver for ncu: prefill + a few 8-frame launches of the persistent decode kernel (0.6B dims)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=256)
tie, tam, tth, tpe = synth_prompt(cfg, T=39)
pol = SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
eng.prefill(0, tie[0].cuda(), 0, pol)
for _ in range(4):
    eng.decode_frames(1, 8, pol, sub)
torch.cuda.synchronize()
st = eng.status(0)
print("ok", st.error, st.n_frames)
