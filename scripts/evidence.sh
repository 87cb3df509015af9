#!/bin/bash
# Round evidence run: all GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture.
# usage: scripts/evidence.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
export FQ3_WATCHDOG_MS=${FQ3_WATCHDOG_MS:-3000}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/tests_gpu_${tag}.log
echo "tests rc=$?" >> gpurun_out/tests_gpu_${tag}.log
tail -3 gpurun_out/tests_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; tail -1 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${tag}.log 2>gpurun_out/bench_${tag}.err; tail -1 gpurun_out/bench_${tag}.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.log 2>&1; tail -1 gpurun_out/bench_ref_${tag}.log | cut -c1-400
if [ "$2" != "nonu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq3_stream -s 3 -c 1 -o gpurun_out/prof_frames_${tag} python scripts/prof_frames.py > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
# tensor-pipe utilisation of the tcgen05 GEMM (codec decode of a 33-frame window): a dozen launches of one decode; the report is
# summarised here and deleted (gpurun_out/ must stay under 64 MiB)
timeout 600 ncu --set full --clock-control none -k regex:fq3c_gemm_tc5 -s 150 -c 14 -o /tmp/prof_gemm_${tag} python scripts/codec_time.py > gpurun_out/ncu_gemm_${tag}.log 2>&1
tail -2 gpurun_out/ncu_gemm_${tag}.log
python scripts/summarize_ncu.py gemm /tmp/prof_gemm_${tag}.ncu-rep gpurun_out/ncu_gemm_summary_${tag}.txt > /dev/null 2>&1
fi
FQ3_PROF=0 timeout 300 python scripts/wide_prof.py 16 14 > gpurun_out/wide_profile_${tag}.log 2>&1
FQ3_PROF=77 timeout 300 python scripts/wide_prof.py 16 300 >> gpurun_out/wide_profile_${tag}.log 2>&1
timeout 600 python scripts/batch_perf.py 0.6B-Base 16 1,2,4,8,16,32,64 > gpurun_out/batch_perf_${tag}.log 2>&1
timeout 300 python scripts/codec_stream_time.py > gpurun_out/codec_stream_${tag}.log 2>&1
timeout 300 python scripts/ttfa_breakdown.py > gpurun_out/ttfa_${tag}.log 2>&1
timeout 300 python scripts/quick_perf.py 0.6B-Base 64 14 > gpurun_out/quick_perf_${tag}.log 2>&1
timeout 300 python scripts/serving_bench.py --concurrent 16 --requests 48 --frames 250 2>&1 | grep "^{" > gpurun_out/serving_${tag}.log
timeout 100 python scripts/serving_bench.py --concurrent 16 --requests 1 --frames 250 2>&1 | grep "^{" >> gpurun_out/serving_${tag}.log
