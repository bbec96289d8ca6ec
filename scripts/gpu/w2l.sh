#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/time_wav2lip.py 16 50 2>&1 | tail -3 | tee gpurun_out/time_w2l.log
timeout 300 python scripts/time_wav2lip.py 1 50 2>&1 | tail -1 | tee -a gpurun_out/time_w2l.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 70 --csv --log-file gpurun_out/launches_w2l.csv python scripts/time_wav2lip.py 16 2 > gpurun_out/ncu_w2l.log 2>&1
tail -2 gpurun_out/ncu_w2l.log
