#!/bin/bash
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=3000
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -4 | tee gpurun_out/tests_tp.log
timeout 300 python scripts/prompt_build_prof.py 2>&1 | grep "prepare_generation:" | tee gpurun_out/tp_perf.log
FQ3_TEXTPROJ_GEMM=0 timeout 300 python scripts/prompt_build_prof.py 2>&1 | grep "prepare_generation:" | tee -a gpurun_out/tp_perf.log
timeout 300 python - <<'PY' 2>&1 | tail -3 | tee -a gpurun_out/tp_perf.log
import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ["FQ3_TEXTPROJ_GEMM"] = "1"
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
m = FasterQwen3TTS.from_pretrained("Qwen/Qwen3-TTS-12Hz-0.6B-Base", device="cuda:0", dtype=torch.bfloat16, attn_implementation="eager", max_seq_len=256, seed=0)
tp = m.model.model.talker.text_projection
x = torch.randn(37, tp.w1.shape[1], generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).cuda()
a = tp(x)
g, tp._gemm = tp._gemm, None
b = tp(x)
tp._gemm = g
ref = torch.nn.functional.linear(torch.nn.functional.silu(torch.nn.functional.linear(x.float(), tp.w1.float(), tp.b1.float()).to(torch.bfloat16).float()).to(torch.bfloat16).float(), tp.w2.float(), tp.b2.float())
print("gemm vs streaming-kernel linear: max abs diff", float((a.float() - b.float()).abs().max()), "of", float(b.float().abs().max()),
      "| vs fp32 reference:", float((a.float() - ref).abs().max()), float((b.float() - ref).abs().max()))
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batched-streams 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['ttfa_ms']['mean'])" | tee -a gpurun_out/tp_perf.log
