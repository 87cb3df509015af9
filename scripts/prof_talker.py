"""Minimal talker-step driver for ncu: prefill (2 launches) + N talker steps."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=64)
tie, tam, tth, tpe = synth_prompt(cfg, T=14)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.0)
eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
eng.prefill(0, tie[0].cuda(), 0, pol)
x = torch.randn(cfg.talker.hidden_size).to(torch.bfloat16).cuda()
for _ in range(5): eng.talker_step(0, x, 14, want_logits=False)
torch.cuda.synchronize()
print("ok", eng.status(0).error)
