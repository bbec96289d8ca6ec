#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_convnet_gpu.py tests/test_musetalk_gpu.py tests/test_wav2lip_gpu.py tests/test_whisper_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_conv.log
tail -8 gpurun_out/pytest_conv.log
timeout -s KILL 300 python scripts/time_musetalk.py 16 2>&1 | tail -3 | tee gpurun_out/time_muse.log
timeout -s KILL 200 python scripts/time_wav2lip.py 16 50 2>&1 | tail -3 | tee gpurun_out/time_w2l.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 800 --csv --log-file gpurun_out/launches_muse.csv python scripts/time_musetalk.py 16 > gpurun_out/ncu_muse.log 2>&1
tail -2 gpurun_out/ncu_muse.log
