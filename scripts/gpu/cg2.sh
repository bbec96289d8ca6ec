#!/bin/bash
mkdir -p gpurun_out
echo "== forced CG=2 correctness"
MF_CONV_CG=2 timeout -s KILL 240 python -m pytest tests/test_convnet_gpu.py tests/test_musetalk_gpu.py tests/test_wav2lip_gpu.py tests/test_whisper_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_cg2.log
tail -6 gpurun_out/pytest_cg2.log
echo "== default"
timeout -s KILL 240 python -m pytest tests/test_convnet_gpu.py tests/test_musetalk_gpu.py tests/test_wav2lip_gpu.py tests/test_whisper_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout -s KILL 200 python scripts/bench_conv.py 0 2>&1 | tee gpurun_out/bench_conv_cg2.log
MF_CONV_CG=1 timeout -s KILL 200 python scripts/bench_conv.py 0 "vae" 2>&1 | tee gpurun_out/bench_conv_cg1.log
timeout -s KILL 300 python scripts/time_musetalk.py 16 2>&1 | tail -2 | tee gpurun_out/time_muse.log
