"""Codec cost per streaming chunk: windowed policy (33-frame window, tail-only) vs stateful stream (8 new frames)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import CodecDecoder, init_codec_synthetic
cfg = preset("0.6B-Base").codec
dec = CodecDecoder(cfg, init_codec_synthetic(cfg, seed=1), "cuda")
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, cfg.codebook_size, (64, cfg.num_quantizers), generator=g).cuda()
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
skip = 25 * cfg.total_upsample
print(f"windowed: 33-frame window, tail-only (skip 25 frames): {timed(lambda: dec.decode(codes[:33], skip_samples=skip)):.3f} ms")
print(f"full 8-frame decode (first chunk): {timed(lambda: dec.decode(codes[:8])):.3f} ms")
st = dec.open_stream(8)
print(f"stateful stream, 8 new frames: {timed(lambda: st.decode(codes[:8])):.3f} ms; ops per chunk {len(st._plans[8].ops)}")
