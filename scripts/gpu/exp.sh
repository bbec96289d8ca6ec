#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | tail -80 > gpurun_out/pytest_edge.log; tail -30 gpurun_out/pytest_edge.log
