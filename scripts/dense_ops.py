"""Per-op device time of the dense prefill plan for R rows."""
import os, sys, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy
from qwen3_tts_cuda_graphs_b200.codec import Op
R = int(sys.argv[1]) if len(sys.argv) > 1 else 239
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=64)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.0)
tie, tam, tth, tpe = synth_prompt(cfg, T=R + 1)
eng.prefill(0, tie[0].cuda(), 0, pol); torch.cuda.synchronize()
d = eng._dense
st = torch.cuda.current_stream().cuda_stream
kinds = {0: "gemm", 2: "rmsnorm", 4: "attn", 8: "qknorm"}
agg = {}
for i in range(len(d.ops)):
    o = d.arr[i]
    arr = (Op * 1)(o)
    for _ in range(2): d.lib.fq3c_run(arr, 1, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): d.lib.fq3c_run(arr, 1, st)
    b.record(); torch.cuda.synchronize()
    key = (kinds.get(o.kind, str(o.kind)), o.M, o.N, o.K, o.flags)
    e = agg.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b) / 5 * 1000
tot = sum(v[1] for v in agg.values())
print(f"R={R}: {len(d.ops)} ops, sum {tot:.0f} us")
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    k, M, N, K, fl = key
    gf = 2.0 * M * N * K * n / (us * 1e-6) / 1e12 if k == "gemm" else 0
    print(f"  {k:8s} M={M:5d} N={N:5d} K={K:5d} flags={fl:3d} x{n:2d}: {us:8.1f} us ({us/n:6.1f} each)  {gf:6.1f} TFLOP/s")
