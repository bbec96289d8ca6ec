mkdir -p gpurun_out; timeout 100 python bench.py --steps 50 --warmup 3 --no-musetalk --no-asr --no-mixed 2>gpurun_out/b_err.log | tail -1 > gpurun_out/bench_cfg1.json; tail -2 gpurun_out/b_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1.json')); print(d['value'], d['heads']['wav2lip']['cpu_plumbing_config1'], d['heads']['wav2lip']['cpu_baseline']['value'])"
