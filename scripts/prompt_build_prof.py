"""Host profile of the prompt assembly (model._prepare_generation) — the part of TTFA in front of the prefill."""
import os, sys, time, cProfile, pstats, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
m = FasterQwen3TTS.from_pretrained("synthetic://0.6B-Base", device="cuda:0", dtype=torch.bfloat16, attn_implementation="eager", max_seq_len=2048, seed=0)
ref = bench.make_ref_wav()
f = lambda: m._prepare_generation(bench.TEXT, ref, bench.REF_TEXT, language="English", non_streaming_mode=True)
for _ in range(5): f()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): f()
torch.cuda.synchronize(); print("prepare_generation: %.3f ms per call" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): f()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
