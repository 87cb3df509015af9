#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_codec_gpu.py -q -m gpu --tb=short -x -k "prefill or codec or oracle or waveform or transformer" 2>&1 | tail -8 | tee gpurun_out/tests_dense2.log
timeout 300 python scripts/prefill_perf.py 0.6B-Base 2>&1 | tail -4 | tee gpurun_out/prefill_perf2.log
timeout 300 python /dev/stdin <<'PY' 2>&1 | tail -3 | tee gpurun_out/codec_time3.log
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer
cfg = preset("0.6B-Base")
dec = SpeechTokenizer.synthetic(cfg.codec, torch.device("cuda"), seed=1).decoder
g = torch.Generator().manual_seed(0)
for T in (8, 33):
    codes = torch.randint(0, cfg.codec.codebook_size, (T, cfg.codec.num_quantizers), generator=g).cuda()
    for _ in range(3): w = dec.decode(codes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): w = dec.decode(codes)
    b.record(); torch.cuda.synchronize()
    print(f"T={T}: {a.elapsed_time(b)/10:.3f} ms per decode, sum {float(w.double().sum()):.6f}")
PY
