#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_ernerf_gpu.py tests/test_ernerf_ref_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_ernerf.log
cat gpurun_out/pytest_ernerf.log
timeout 300 python scripts/time_ernerf.py 2>&1 | tail -20 | tee gpurun_out/time_ernerf.log
