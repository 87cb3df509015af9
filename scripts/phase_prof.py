"""Per-phase timeline of one CTA for a talker step (FQ3_PROF=<cta>)."""
import os, sys, ctypes as C, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
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
buf = (C.c_longlong * (1024 * 8))()
eng.lib.fq3_debug_read_prof(eng.h, buf, 1024 * 8)   # clear
eng.talker_step(0, x, 14, want_logits=False)
eng.lib.fq3_debug_read_prof(eng.h, buf, 1024 * 8)
n = 141
names = {0: "qkv", 1: "attn", 2: "o", 3: "gu", 4: "down"}
t0 = min(buf[i * 8] for i in range(n) if buf[i * 8])
print("phase kind     start  d(load/step1) d(tiles/sweep) d(tail)   [cycles]")
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
prev_end = None
for i in range(n):
    m = [buf[i * 8 + k] for k in range(4)]
    if m[0] == 0: continue
    kind = names.get(i % 5, "?") if i < 140 else "head"
    d = [m[1] - m[0], m[2] - m[1], m[3] - m[2]]
    pr = [buf[i * 8 + 4], buf[i * 8 + 5]]
    if i < 12: print("       gemv_tile cycles", buf[i * 8 + 6], "calls", buf[i * 8 + 7])
    if i < 12 or i >= 136: print(f"{i:4d} {kind:5s} {m[0]-t0:9d} {d[0]:9d} {d[1]:9d} {d[2]:9d}   producer first/last issue at {pr[0]-t0 if pr[0] else 0:9d} {pr[1]-t0 if pr[1] else 0:9d}")
    a = agg[kind]; a[0] += 1; a[1] += d[0]; a[2] += d[1]; a[3] += d[2]
    if prev_end is not None: a[4] += m[0] - prev_end
    prev_end = m[3]
print("kind    n  avg load/step1  avg tiles/sweep  avg tail  avg gap-before")
for k, a in agg.items(): print(f"{k:5s} {a[0]:3d} {a[1]/a[0]:12.0f} {a[2]/a[0]:14.0f} {a[3]/a[0]:10.0f} {a[4]/a[0]:12.0f}")
last = max(buf[i * 8 + 3] for i in range(n)); print("total cycles", last - t0)
