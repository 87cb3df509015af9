"""Per-phase publish-time skew across CTAs for one talker step (FQ3_PROF=-2)."""
import os, sys, ctypes as C, collections
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
G = eng.num_sms
late = collections.Counter(); agg = collections.defaultdict(list)
prev_last = None
for ph in range(141):
    kind = names[ph % 5] if ph < 140 else "head"
    if kind == "attn": continue
    st = [buf[(ph * 160 + c) * 2] for c in range(G)]
    pb = [(buf[(ph * 160 + c) * 2 + 1], c) for c in range(G) if buf[(ph * 160 + c) * 2 + 1]]
    if not pb: continue
    ts = sorted(t for t, _ in pb)
    last_t, last_c = max(pb)
    med = ts[len(ts) // 2]
    late[last_c] += 1
    s0 = sorted(s for s in st if s)
    agg[kind].append((last_t - ts[0], last_t - med, s0[-1] - s0[0], med - s0[len(s0)//2], (ts[0] - prev_last) if prev_last else 0))
    if ph < 10 or ph > 135:
        print(f"ph {ph:3d} {kind:5s} publish spread {last_t - ts[0]:6d} ns, last-median {last_t - med:6d} ns (last cta {last_c}), start spread {s0[-1]-s0[0]:6d} ns, median start->publish {med - s0[len(s0)//2]:6d} ns")
    prev_last = last_t
print("kind: avg publish spread | avg last-median | avg start spread | median start->median publish   (ns)")
for k, v in agg.items():
    n = len(v)
    print(f"{k:5s} {sum(a[0] for a in v)/n:8.0f} {sum(a[1] for a in v)/n:8.0f} {sum(a[2] for a in v)/n:8.0f} {sum(a[3] for a in v)/n:8.0f}")
print("most often last:", late.most_common(12))

import statistics
lag = collections.defaultdict(lambda: [0.0] * G)
cnt = collections.Counter()
for ph in range(141):
    kind = names[ph % 5] if ph < 140 else "head"
    if kind == "attn": continue
    pb = [buf[(ph * 160 + c) * 2 + 1] for c in range(G)]
    if not all(pb): continue
    t0 = min(pb)
    for c in range(G): lag[kind][c] += pb[c] - t0
    cnt[kind] += 1
for kind in lag:
    print(f"average publish lag per CTA [{kind}] (ns), rows of 37 CTAs:")
    v = [x / cnt[kind] for x in lag[kind]]
    for r in range(0, G, 37): print("   " + " ".join(f"{int(x):4d}" for x in v[r:r+37]))
