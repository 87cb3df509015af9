"""Per-phase durations of one predictor run on CTA 0 (FQ3_PROF=0): GEMV kinds, attention, sampling."""
import os, sys, ctypes as C, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine
from qwen3_tts_cuda_graphs_b200.engine import SubPolicy
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=64)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
pin = torch.randn(2, cfg.talker.hidden_size).to(torch.bfloat16).cuda()
for _ in range(3): eng.predictor_run(0, pin, sub)
torch.cuda.synchronize()
N = 512 * 160 * 2
buf = (C.c_longlong * N)()
eng.lib.fq3_debug_read_prof(eng.h, buf, N)
eng.predictor_run(0, pin, sub)
eng.lib.fq3_debug_read_prof(eng.h, buf, N)
S = 16
per_pass = 27  # 5 layers x 5 phases + head + sample
names = ["qkv", "attn", "o", "gu", "down"]
starts = {}
ends = {}
for i in range(15 * per_pass):
    if buf[i * S]: starts[i] = buf[i * S]
    if buf[i * S + 3]: ends[i] = buf[i * S + 3]
agg = collections.defaultdict(list)
for i in range(15 * per_pass - 1):
    j = i % per_pass
    kind = names[j % 5] if j < 25 else ("head" if j == 25 else "sample")
    if kind == "sample":
        if (i - 1) in ends and (i + 1) in starts: agg[kind].append(starts[i + 1] - ends[i - 1])
    elif i in starts and (i + 1) in starts and ((i + 1) % per_pass) != 26:
        agg[kind].append(starts[i + 1] - starts[i])
    elif i in starts and i in ends:
        agg[kind].append(ends[i] - starts[i])
tot = 0
for k, v in agg.items():
    if not v: continue
    v2 = sorted(v)
    print(f"{k:7s} n={len(v):3d} median {v2[len(v2)//2]:6d} cycles  min {v2[0]:6d} max {v2[-1]:6d}  sum {sum(v):8d}")
    tot += sum(v)
print("sum of phases", tot, "cycles =", tot / 1.965e3, "us")
pp = [starts[(k + 1) * per_pass] - starts[k * per_pass] for k in range(14) if (k + 1) * per_pass in starts and k * per_pass in starts]
print("cycles per pass (start of qkv to start of the next pass's qkv):", pp)
for k in (0, 1):
    print(f"pass {k} phases:", [(starts[k * per_pass + j + 1] - starts[k * per_pass + j]) for j in range(per_pass) if k * per_pass + j + 1 in starts and k * per_pass + j in starts])

print("sample phases: start->logits polled | prep+max | top-k | softmax+draw | publish | end barrier   (cycles)")
rows = []
for i in range(26, 15 * per_pass, per_pass):
    m = [buf[i * S + k] for k in range(8)]
    if not m[0] or not m[3]: continue
    rows.append((m[1] - m[0], m[2] - m[1], m[4] - m[2], m[5] - m[4], m[6] - m[5], m[3] - m[6], m[3] - m[0]))
for r in rows[:4]: print("   ", r)
n = max(len(rows), 1)
print("avg ", tuple(int(sum(r[k] for r in rows) / n) for k in range(7)))

if os.environ.get("FQ3_PROF") == "-3":
    for i in (26, 53, 80):
        a = [buf[(i * 160 + wp) * 2] for wp in range(12)]
        b = [buf[(i * 160 + wp) * 2 + 1] for wp in range(12)]
        if not any(a): continue
        t0 = min(v for v in a + b if v)
        print(f"sample phase {i}: poll-done per warp " + " ".join(f"{v - t0:6d}" for v in a))
        print(f"                 at barrier 1      " + " ".join(f"{v - t0:6d}" for v in b))
