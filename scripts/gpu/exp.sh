#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convnet_gpu.py tests/test_wav2lip_gpu.py tests/test_musetalk_gpu.py tests/test_whisper_gpu.py tests/test_wav2vec2_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python scripts/bench_conv.py 0 w2l256 2>&1 | tail -5 | tee gpurun_out/bench_conv_exp.log
timeout 300 python scripts/bench_conv.py 0 vae 2>&1 | tail -5 | tee -a gpurun_out/bench_conv_exp.log
timeout 200 python scripts/time_wav2lip.py 16 50 96 2>&1 | tail -1 | tee -a gpurun_out/bench_conv_exp.log
timeout 200 python scripts/time_wav2lip.py 16 50 256 2>&1 | tail -1 | tee -a gpurun_out/bench_conv_exp.log
timeout 300 python scripts/time_musetalk.py 16 2>&1 | tail -2 | tee -a gpurun_out/bench_conv_exp.log
