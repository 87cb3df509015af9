"""Windowed streaming decode cost vs window size (new frames + 25 context frames, tail-only): device time per call, graph replay."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import CodecDecoder, init_codec_synthetic
cfg = preset("0.6B-Base").codec
dec = CodecDecoder(cfg, init_codec_synthetic(cfg, seed=1), "cuda")
codes = torch.randint(0, cfg.codebook_size, (96, cfg.num_quantizers), generator=torch.Generator().manual_seed(0)).cuda()
import time
for new in (8, 12, 16, 24):
    T, skip = new + 25, 25 * cfg.total_upsample
    fn = lambda: dec.decode(codes[:T], skip_samples=skip)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(20): fn()
    b.record(); host = (time.perf_counter() - t0) / 20; torch.cuda.synchronize()
    plan = dec._plans[(T, skip)]
    print(f"{new} new + 25 context = {T} frames: {a.elapsed_time(b) / 20:.3f} ms device, {host * 1000:.3f} ms host enqueue, graph={'yes' if plan.graph is not None else 'NO'} ops={len(plan.ops)}")
