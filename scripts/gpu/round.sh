#!/bin/bash
# one GPU call: the whole -m gpu suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout -s KILL 700 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1000 python bench.py > gpurun_out/bench_stdout.log 2> gpurun_out/bench_stderr.log
tail -1 gpurun_out/bench_stdout.log > gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_stderr.log
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n1.json'))
    print('ernerf', d['value'], 'e2e', d['e2e']['value'], 'p50', d['p50_chunk_to_frame_ms'], 'roof', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    for h,v in d['heads'].items(): print(h, round(v['value'],1), 'ms', round(v['ms_per_step'],3), 'e2e', round(v['e2e']['value'],1), 'roof', v.get('roofline',{}).get('frac'), {k:v[k] for k in ('wav2lip_engine_calls_per_step','realtime_sessions_capacity','algorithmic_tflops') if k in v})
except Exception as e:
    print('bench parse failed', e)
PY
