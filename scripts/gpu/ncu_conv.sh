#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tma -s 3 -c 1 -f -o gpurun_out/conv64_256 python scripts/bench_conv.py 0 "64->64 @256 +res" > gpurun_out/ncu_conv64.log 2>&1
tail -3 gpurun_out/ncu_conv64.log
ls -la gpurun_out/*.ncu-rep
