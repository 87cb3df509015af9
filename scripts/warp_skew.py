"""Per-warp arrival times inside CTA 0 for one talker step (FQ3_PROF=-3)."""
import os, sys, ctypes as C
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
for _ in range(3): eng.talker_step(0, x, 14, want_logits=False)
torch.cuda.synchronize()
N = 512 * 160 * 2
buf = (C.c_longlong * N)()
eng.lib.fq3_debug_read_prof(eng.h, buf, N)
eng.talker_step(0, x, 14, want_logits=False)
eng.lib.fq3_debug_read_prof(eng.h, buf, N)
names = {0: "qkv", 1: "attn", 2: "o", 3: "gu", 4: "down"}
prev_end = None
for ph in list(range(0, 15)) + list(range(135, 141)):
    kind = names[ph % 5] if ph < 140 else "head"
    if kind == "attn": continue
    def col(k): return [buf[(ph * 160 + 16 * (k >> 1) + wp) * 2 + (k & 1)] for wp in range(12)]
    a = col(0)
    if not any(a): continue
    t0 = min(v for v in a if v)
    print(f"ph {ph:3d} {kind:5s} (cycles from the first warp entering the phase; warps 0-7 poll, leaders sit on the high warps)")
    for k, nm in ((0, "enter"), (5, "poll done"), (6, "at stage bar"), (1, "past stage bar"), (2, "mma start"), (3, "mma end"), (7, "published"), (4, "leave")):
        c = col(k)
        if any(c): print(f"    {nm:>14s}: " + " ".join(f"{(v - t0) if v else -1:5d}" for v in c))
