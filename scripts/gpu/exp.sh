#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ernerf_gpu.py tests/test_ernerf_ref_gpu.py tests/test_edge_cases_gpu.py tests/test_plugin_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_ernerf.log; tail -12 gpurun_out/pytest_ernerf.log
timeout 300 python scripts/time_ernerf.py 2>&1 | grep -v rounds | tail -6 | tee gpurun_out/time_ernerf.log
timeout 300 python scripts/time_ernerf.py 2>&1 | grep "H=" | cut -c1-60
