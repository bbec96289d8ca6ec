#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json'))
print('ernerf', d['value'], 'e2e', d['e2e']['value'], 'p50', d['p50_chunk_to_frame_ms'])
for h,v in d['heads'].items(): print(h, v['value'], 'e2e', v['e2e']['value'], 'p50', v.get('p50_chunk_to_frame_ms'), 'roof', v['roofline']['frac'])
"
# ErNeRF: launch list + full capture of k_head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ernerf.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-wav2lip --no-musetalk > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_head -s 3 -c 2 -o gpurun_out/prof_k_head -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-wav2lip --no-musetalk > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
