"""Per-phase timeline of one CTA for a talker step (FQ3_PROF=<cta>): clock64 marks of thread 0, the MMA thread and the producer."""
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
L = cfg.talker.num_hidden_layers
n = 5 * L + 1
names = {0: "qkv", 1: "attn", 2: "o", 3: "gu", 4: "down"}
S = 16
t0 = min(buf[i * S] for i in range(n) if buf[i * S])
cols = ["poll", "n0", "n1", "n2", "pre", "mma", "red", "epi", "back"]
print("GEMV (thread 0 = a polling warp; red / epi / back = lane 0 of warp 11, the leader of group 0): poll = own input words visible; n0 = sum of squares + norm weights; n1 = barrier; n2 = rescale + stage x + barrier; pre = to the first MMA;")
print("      mma = k loop of the warp's unit; red = k-parts through shared memory; epi = epilogue + publish; back = hand the stages back")
print(f"{'ph':>4s} {'kind':5s} {'start':>9s} " + " ".join(f"{x:>6s}" for x in cols))
agg = collections.defaultdict(lambda: [0] * 12)
prev_end = None
for i in range(n):
    m = [buf[i * S + k] for k in range(S)]
    if m[0] == 0: continue
    kind = names.get(i % 5, "?") if i < 5 * L else "head"
    if kind == "attn":
        d = [m[1] - m[0], m[2] - m[1], m[3] - m[2]] + [0] * 6
    else:
        n0 = (m[15] - m[7]) if m[15] else 0
        n1 = (m[14] - m[15]) if m[15] else 0
        n2 = (m[1] - m[14]) if m[14] else (m[1] - m[7])
        d = [m[7] - m[0], n0, n1, n2, m[8] - m[1], m[9] - m[8], m[10] - m[9], m[11] - m[10], m[3] - m[11]]
    if i < 12 or i >= n - 5:
        print(f"{i:4d} {kind:5s} {m[0]-t0:9d} " + " ".join(f"{x:6d}" for x in d) + f"   prod {m[4]-t0 if m[4] else 0} {m[5]-t0 if m[5] else 0}")
    a = agg[kind]; a[0] += 1
    for k in range(9): a[1 + k] += d[k]
    if prev_end is not None: a[10] += m[0] - prev_end
    a[11] += m[3] - m[0]
    prev_end = m[3]
print("averages (last columns: gap before the phase, phase total)")
for k, a in agg.items(): print(f"{k:5s} {a[0]:3d}           " + " ".join(f"{a[1+j]/a[0]:6.0f}" for j in range(9)) + f" {a[10]/a[0]:8.0f} {a[11]/a[0]:8.0f}")
last = max(buf[i * S + 3] for i in range(n)); print("total cycles", last - t0)
