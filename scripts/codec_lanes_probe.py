"""Codec cost of ONE serving cycle (16 utterances x one chunk) on an otherwise idle GPU: one CUDA stream vs several lanes,
windowed (33-frame window, tail-only) vs stateful (8 new frames per utterance).  Wall clock around enqueue + synchronize."""
import copy, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import CodecDecoder, init_codec_synthetic
cfg = preset("0.6B-Base").codec
dec = CodecDecoder(cfg, init_codec_synthetic(cfg, seed=1), "cuda")
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, cfg.codebook_size, (64, cfg.num_quantizers), generator=g).cuda()
N = 16
skip = 25 * cfg.total_upsample
for lanes in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    decs = []
    for _ in range(lanes):
        d = copy.copy(dec); d._plans = {}; decs.append(d)
    cs = [decs[i % lanes].open_stream(8) for i in range(N)]
    def cycle(kind):
        outs = []
        for i in range(N):
            with torch.cuda.stream(streams[i % lanes]):
                outs.append(cs[i].decode(codes[:8]) if kind == "stateful" else decs[i % lanes].decode(codes[:33], skip_samples=skip))
        host_enqueue = time.perf_counter()
        for s in streams:
            s.synchronize()
        return host_enqueue
    for kind in ("windowed", "stateful"):
        for _ in range(3):
            cycle(kind)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); enq = 0.0
        for _ in range(5):
            ta = time.perf_counter(); enq += cycle(kind) - ta
        dt = (time.perf_counter() - t0) / 5
        print(f"lanes {lanes} {kind}: {dt * 1000:.2f} ms per cycle of {N} utterances ({dt * 1000 / N:.2f} ms each), host enqueue {enq / 5 * 1000:.2f} ms")
