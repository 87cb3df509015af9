#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_talker.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fq3_stream -s 4 -c 1 -o gpurun_out/prof_talker_r1c python scripts/prof_talker.py > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/ncu.log
