#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_convnet_gpu.py tests/test_wav2lip_gpu.py tests/test_plugin_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for d in 0 7; do echo "dbg=$d (no graph)"; MF_NO_GRAPH=1 MF_CONV_DBG=$d timeout 120 python scripts/time_wav2lip.py 1 30 2>&1 | tail -1; done | tee gpurun_out/conv_dbg.log
timeout 300 python scripts/time_wav2lip.py 1 50 2>&1 | tail -1 | tee gpurun_out/time_w2l.log
timeout 300 python scripts/time_wav2lip.py 16 50 2>&1 | tail -1 | tee -a gpurun_out/time_w2l.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 70 --csv --log-file gpurun_out/launches_w2l.csv python scripts/time_wav2lip.py 16 2 > gpurun_out/ncu_w2l.log 2>&1
