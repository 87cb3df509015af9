"""Per-op device time of the codec decode plan (T frames); run once per FQ3C_TCGEN05 setting."""
import os, sys, torch, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer, Op
T = int(sys.argv[1]) if len(sys.argv) > 1 else 33
cfg = preset("0.6B-Base")
tok = SpeechTokenizer.synthetic(cfg.codec, torch.device("cuda"), seed=1)
dec = tok.decoder
codes = torch.randint(0, cfg.codec.codebook_size, (T, cfg.codec.num_quantizers)).cuda()
dec.decode(codes); torch.cuda.synchronize()
plan = dec._plans[T]
kinds = {0: "gemm", 1: "rvq", 2: "rmsnorm", 3: "rope", 4: "attn", 5: "dwconv", 6: "layernorm", 7: "snake"}
rows = []
st = torch.cuda.current_stream().cuda_stream
for i, o in enumerate(plan.ops):
    arr = (Op * 1)(o)
    for _ in range(2): dec.lib.fq3c_run(arr, 1, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): dec.lib.fq3c_run(arr, 1, st)
    b.record(); torch.cuda.synchronize()
    rows.append((a.elapsed_time(b) / 5 * 1000, i, kinds[o.kind], o.M, o.N, o.K, o.flags))
tot = sum(r[0] for r in rows)
print(f"T={T} FQ3C_TCGEN05={os.environ.get('FQ3C_TCGEN05','1')}: {len(rows)} ops, sum {tot:.0f} us")
agg = {}
for us, i, k, M, N, K, fl in rows:
    key = (k, M, N, K, fl)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    k, M, N, K, fl = key
    gf = 2.0 * M * N * K * n / (us * 1e-6) / 1e12 if k == "gemm" else 0
    print(f"  {k:8s} M={M:6d} N={N:5d} K={K:5d} flags={fl:4d} x{n:2d}: {us:8.1f} us  {gf:6.1f} TFLOP/s")
