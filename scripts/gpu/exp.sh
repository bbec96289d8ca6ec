#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scheduler_gpu.py tests/test_ernerf_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_sched.log; tail -12 gpurun_out/pytest_sched.log
timeout 900 python bench.py --no-asr --no-cpu-baseline > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.log').read().strip().splitlines()[-1])
print('ernerf', d['value'], d['roofline']['frac'])
print(d['ernerf_batched_sessions'])
m=d['heads']['mixed_sessions']; print('mixed', m['value'], m['ms_per_step'], m['e2e']['value'])
PY
