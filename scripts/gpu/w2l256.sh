#!/bin/bash
# launch list + timing of the 256x256 extended Wav2Lip generator (B = 16)
mkdir -p gpurun_out
timeout 300 python scripts/time_wav2lip.py 16 50 256 2>&1 | tail -2 | tee gpurun_out/time_w2l256.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w2l256.csv python scripts/time_wav2lip.py 16 2 256 > gpurun_out/ncu_w2l256.log 2>&1
tail -2 gpurun_out/ncu_w2l256.log
