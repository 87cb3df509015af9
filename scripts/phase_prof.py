"""Per-phase timeline of one CTA for a talker step (FQ3_PROF=<cta>)."""
import os, sys, ctypes as C, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
cfg = make_cfg(sys.argv[1] if len(sys.argv) > 1 else "0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=64)
tie, tam, tth, tpe = synth_prompt(cfg, T=14)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.0)
eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
eng.prefill(0, tie[0].cuda(), 0, pol)
x = torch.randn(cfg.talker.hidden_size).to(torch.bfloat16).cuda()
for _ in range(3): eng.talker_step(0, x, 14, want_logits=False)
torch.cuda.synchronize()
buf = (C.c_longlong * (512 * 16))()
eng.lib.fq3_debug_read_prof(eng.h, buf, 512 * 16)   # clear
eng.talker_step(0, x, 14, want_logits=False)
eng.lib.fq3_debug_read_prof(eng.h, buf, 512 * 16)
n = 141
names = {0: "qkv", 1: "attn", 2: "o", 3: "gu", 4: "down"}
S = 16
t0 = min(buf[i * S] for i in range(n) if buf[i * S])
cols = ["poll", "n0", "n1", "n2", "ring", "mma", "rest", "bar", "f0", "sum", "epi", "st+tl"]
print("GEMV: poll = own input words visible; n0 = sum of squares + norm weights; n1 = first barrier (all warps' words visible); n2 = rescale + barrier;")
print("      ring = loop set-up + wait full; mma = one unit; rest = partial store / other rounds; bar = barrier; f0/sum/epi/st/tail = finish")
print(f"{'ph':>4s} {'kind':5s} {'start':>9s} " + " ".join(f"{x:>6s}" for x in cols))
agg = collections.defaultdict(lambda: [0] + [0] * 13)
dist = collections.defaultdict(list)
prev_end = None
for i in range(n):
    m = [buf[i * S + k] for k in range(S)]
    if m[0] == 0: continue
    kind = names.get(i % 5, "?") if i < 140 else "head"
    if kind == "attn":
        d = [m[1] - m[0], m[2] - m[1], m[3] - m[2]] + [0] * 9
    else:
        n1 = (m[14] - m[7]) if m[14] else (m[1] - m[7])
        n2 = (m[1] - m[14]) if m[14] else 0
        n0 = (m[15] - m[7]) if m[15] else 0
        if m[15]: n1 = m[14] - m[15]
        d = [m[7] - m[0], n0, n1, n2, m[8] - m[1], m[9] - m[8], m[6] - m[9], m[2] - m[6],
             m[10] - m[2], m[11] - m[10], m[12] - m[11], m[3] - m[12]]
    if i < 12 or i >= 136:
        print(f"{i:4d} {kind:5s} {m[0]-t0:9d} " + " ".join(f"{x:6d}" for x in d) + f"   prod {m[4]-t0 if m[4] else 0} {m[5]-t0 if m[5] else 0}")
    dist[kind].append((d[0] + d[1] + d[2] + d[3], d[4], d[5] + d[6], sum(d[8:12])))
    a = agg[kind]; a[0] += 1
    for k in range(12): a[1 + k] += d[k]
    if prev_end is not None: a[13] += m[0] - prev_end
    prev_end = m[3]
print("averages (last column: gap before the phase)")
for k, a in agg.items(): print(f"{k:5s} {a[0]:3d}           " + " ".join(f"{a[1+j]/a[0]:6.0f}" for j in range(12)) + f" {a[13]/a[0]:8.0f}")
last = max(buf[i * S + 3] for i in range(n)); print("total cycles", last - t0)
print("distribution per kind: min / median / max of  wait(poll+n1+n2) | ring | mma+rest | finish(all)")
for k, v in dist.items():
    cs = list(zip(*v))
    def q(c): c = sorted(c); return f"{c[0]:6d}/{c[len(c)//2]:6d}/{c[-1]:6d}"
    print(f"{k:5s} " + " | ".join(q(c) for c in cs))
