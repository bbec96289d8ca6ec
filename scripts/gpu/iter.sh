#!/bin/bash
# one GPU iteration: all gpu tests, conv microbench, head timings (NCU=1 adds the musetalk launch list)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python scripts/bench_conv.py ${MODES:-0} 2>&1 | tee gpurun_out/bench_conv.log
timeout -s KILL 300 python scripts/time_musetalk.py 16 2>&1 | tail -2 | tee gpurun_out/time_muse.log
timeout -s KILL 200 python scripts/time_wav2lip.py 16 50 2>&1 | tail -1 | tee gpurun_out/time_w2l.log
if [ -n "$NCU" ]; then
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1100 --csv --log-file gpurun_out/launches_muse.csv python scripts/time_musetalk.py 16 > gpurun_out/ncu_muse.log 2>&1
fi
