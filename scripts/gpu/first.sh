#!/bin/bash
mkdir -p gpurun_out
python scripts/dump_level_scales.py > gpurun_out/scales.log 2>&1
cp gpurun_out/ernerf_level_scales.json tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_ernerf.json
# launch list (cold-cache, serialised) of a short bench run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ernerf.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_head -s 3 -c 2 -o gpurun_out/prof_k_head -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
