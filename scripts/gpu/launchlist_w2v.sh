#!/bin/bash
# ncu launch list of one wav2vec2 window (fused-stack engine).  usage: launchlist_w2v.sh <tag>
TAG=${1:-v1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w2v_$TAG.csv \
  python scripts/time_w2v.py > gpurun_out/ncu_w2v.log 2>&1
tail -3 gpurun_out/ncu_w2v.log
wc -l gpurun_out/launches_w2v_$TAG.csv
