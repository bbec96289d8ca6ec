#!/bin/bash
# ncu launch list of the bench command (ErNeRF legs only: single frame, 2048 rays, 4 batched sessions).  usage: launchlist.sh <tag>
TAG=${1:-v5}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ernerf_$TAG.csv \
  python bench.py --steps 20 --warmup 3 --no-wav2lip --no-musetalk --no-asr --no-cpu-baseline --no-mixed --no-reference-gpu > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
wc -l gpurun_out/launches_ernerf_$TAG.csv
