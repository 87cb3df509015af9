#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -8 | tee gpurun_out/tests_tail.log
timeout 300 python - <<'PY' 2>&1 | tail -4 | tee gpurun_out/tail_perf.log
import os, sys, torch
sys.path.insert(0, os.getcwd())
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer
cfg = preset("0.6B-Base")
dec = SpeechTokenizer.synthetic(cfg.codec, torch.device("cuda"), seed=1).decoder
codes = torch.randint(0, cfg.codec.codebook_size, (33, cfg.codec.num_quantizers), generator=torch.Generator().manual_seed(0)).cuda()
n = dec.n_samples(33)
for skip in (0, int(round(25 * n / 33))):
    for _ in range(3): w = dec.decode(codes, skip)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): w = dec.decode(codes, skip)
    b.record(); torch.cuda.synchronize()
    print(f"T=33 skip={skip}: {a.elapsed_time(b)/10:.3f} ms per decode")
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batched-streams 0 2>gpurun_out/bench_dbg.err | tail -1 > gpurun_out/bench_tail.json; python -c "
import json; d=json.load(open('gpurun_out/bench_tail.json')); print(d['value'], d['e2e']['value'], d['ttfa_ms']['mean'])"
