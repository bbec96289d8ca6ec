#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json'))
print('ernerf', d['value'], 'e2e', d['e2e']['value'], 'p50', d['p50_chunk_to_frame_ms'], 'cpu', d['cpu_baseline']['value'])
for h,v in d['heads'].items(): print(h, v['value'], 'e2e', v['e2e']['value'], 'p50', v.get('p50_chunk_to_frame_ms'), 'roof', v['roofline']['frac'], 'cpu', v.get('cpu_baseline',{}).get('value'))
"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
