#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python scripts/time_wav2lip.py 16 50 96 2>&1 | tail -1 | tee gpurun_out/time_w2l96.log
timeout 200 python scripts/time_wav2lip.py 16 50 256 2>&1 | tail -1 | tee gpurun_out/time_w2l256.log
timeout 300 python scripts/bench_conv.py 0,1,2,4 w2l256 2>&1 | tail -4 | tee gpurun_out/bench_conv_w2l256.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w2l256.csv python scripts/time_wav2lip.py 16 2 256 > gpurun_out/ncu_w2l256.log 2>&1
