#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_convnet_gpu.py tests/test_wav2lip_gpu.py tests/test_scheduler_gpu.py -m gpu -q 2>&1 | tail -3
for i in 1 2; do
timeout 200 python scripts/time_wav2lip.py 16 50 96 2>&1 | tail -1 | tee -a gpurun_out/bench_conv_exp.log
timeout 200 python scripts/time_wav2lip.py 16 50 256 2>&1 | tail -1 | tee -a gpurun_out/bench_conv_exp.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w2l256.csv python scripts/time_wav2lip.py 16 2 256 > gpurun_out/ncu_w2l256.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w2l96.csv python scripts/time_wav2lip.py 16 2 96 > gpurun_out/ncu_w2l96.log 2>&1
