#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_convnet_gpu.py tests/test_musetalk_gpu.py tests/test_wav2lip_gpu.py tests/test_whisper_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_halo.log
tail -12 gpurun_out/pytest_halo.log
MF_CONV_VERBOSE=1 timeout -s KILL 200 python scripts/bench_conv.py 0 "vae" 2>&1 | grep -v "op [123] " | tee gpurun_out/bench_conv_halo.log
MF_CONV_HALO=0 timeout -s KILL 200 python scripts/bench_conv.py 0 "vae" 2>&1 | tee gpurun_out/bench_conv_nohalo.log
timeout -s KILL 300 python scripts/time_musetalk.py 16 2>&1 | tail -2 | tee gpurun_out/time_muse.log
