#!/bin/bash
# one `ncu --set full` capture of the dominant ErNeRF kernel (k_head) inside the bench command, and the per-launch DRAM traffic
# derived from it -> gpurun_out/k_head_traffic.json (copied to profiles/ by hand together with the summary).  usage: ncu_head.sh <tag>
TAG=${1:-v5}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_head -s 4 -c 1 -f -o gpurun_out/k_head_$TAG \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-wav2lip --no-musetalk --no-asr --no-mixed --no-reference-gpu > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log | cut -c1-300
ncu -i gpurun_out/k_head_$TAG.ncu-rep --page raw --csv > gpurun_out/k_head_${TAG}_raw.csv 2>/dev/null
python - "$TAG" <<'PY'
import csv, json, sys
tag = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/k_head_{tag}_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult
out = {"kernel": "k_head", "tag": tag, "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum")}
out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
for k in ("gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
          "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"):
    try:
        out[k] = get(k)
    except ValueError:
        pass
json.dump(out, open("gpurun_out/k_head_traffic.json", "w"), indent=1)
print(out)
PY
ls -la gpurun_out/*.ncu-rep
