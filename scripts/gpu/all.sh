#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_musetalk_gpu.py -m gpu -q -s -k small_config 2>&1 | grep -E "PSNR|passed|failed" | tee gpurun_out/pytest_muse.log
timeout -s KILL 900 python scripts/time_musetalk.py 16 2>&1 | tail -4 | tee gpurun_out/time_muse.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/launches_muse.csv python scripts/time_musetalk.py 16 > gpurun_out/ncu_muse.log 2>&1
tail -2 gpurun_out/ncu_muse.log
