"""Where a frame of the wide program goes (build with scripts/build_variant.sh wprof -DFQ3_WIDE_PROF=1; FQ3_VARIANT=wprof FQ3_PROF=<cta>): cycles per category of one CTA's thread 0, per frame."""
import os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("FQ3_VARIANT"): sys.path.insert(0, os.path.join(ROOT, "variants", os.environ["FQ3_VARIANT"]))  # A/B builds (scripts/build_variant.sh)
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 14
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=1024, max_streams=ns, max_frames=512)
pol = SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
for s in range(ns):
    tie, tam, tth, tpe = synth_prompt(cfg, T=T, seed=1 + s)
    eng.set_text_conditioning(s, tth[0].cuda(), tpe.cuda())
    eng.prefill(s, tie[0].cuda(), 0, pol)
eng.decode_frames(ns, 4, pol, sub)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
eng.lib.fq3_debug_read_prof(eng.h, buf, 64)
frames = 8
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.decode_frames(ns, frames, pol, sub); b.record(); torch.cuda.synchronize()
eng.lib.fq3_debug_read_prof(eng.h, buf, 64)
names = ["entry barrier", "poll + stage", "norm", "mma + reduce + epilogue", "attention", "sample"]
tot = sum(buf[i] for i in range(6))
print(f"{ns} streams, context {T}: {a.elapsed_time(b) / frames:.3f} ms per frame-step; CTA {os.environ.get('FQ3_PROF')} thread 0, {buf[6] // frames} GEMV phases per frame")
if tot == 0:
    print("  (the per-category cycle marks of the wide kernel are compiled out by default: build with -DFQ3_WIDE_PROF=1 for the split)")
    sys.exit(0)
for i, n in enumerate(names):
    print(f"  {n:26s} {buf[i] / frames:10.0f} cycles/frame  {100 * buf[i] / tot:5.1f} %")
for i, n in ((12, "multi-attn: entry barrier"), (13, "multi-attn: q / fresh K,V"), (14, "multi-attn: scores + request"), (15, "multi-attn: barrier 1"),
             (18, "multi-attn: probabilities"), (19, "multi-attn: barrier 2"), (16, "multi-attn: output + merge"), (17, "multi-attn: barrier 3"), (8, "leader: wait for weights"), (9, "leader: k loop"), (10, "leader: k-parts"), (11, "leader: epilogue + publish")):
    print(f"  {n:26s} {buf[i] / frames:10.0f} cycles/frame")
print(f"  {'sum':26s} {tot / frames:10.0f} cycles/frame = {tot / frames / 1.965e3:.0f} us at 1.965 GHz")
