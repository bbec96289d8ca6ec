#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_head -s 4 -c 1 -f -o gpurun_out/k_head_v4 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-wav2lip --no-musetalk --no-asr --no-mixed > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
