#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codec_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -25 | tee gpurun_out/tests_codec_tc5.log
cat > /tmp/codec_time.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer
cfg = preset("0.6B-Base")
tok = SpeechTokenizer.synthetic(cfg.codec, torch.device("cuda"), seed=1)
dec = tok.decoder
g = torch.Generator().manual_seed(0)
for T in (8, 33):
    codes = torch.randint(0, cfg.codec.codebook_size, (T, cfg.codec.num_quantizers), generator=g).cuda()
    for _ in range(3): w = dec.decode(codes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): w = dec.decode(codes)
    b.record(); torch.cuda.synchronize()
    print(f"T={T}: {a.elapsed_time(b)/10:.3f} ms per decode, wav {tuple(w.shape)} absmax {float(w.abs().max()):.4f} sum {float(w.double().sum()):.6f}")
PY
for v in 1 0; do echo "FQ3C_TCGEN05=$v"; FQ3C_TCGEN05=$v timeout 300 python /tmp/codec_time.py 2>&1 | tail -3; done | tee gpurun_out/codec_time.log
