#!/bin/bash
# usage: KREGEX=attn_bwd_tc SKIP=5 NAME=bwd_vit CMD="python scripts/attn_bench.py vit_self" bash scripts/gpu_ncu_cmd.sh
# one `ncu --set full` capture (source sampling on) of the SKIP+1-th launch matching KREGEX inside CMD
mkdir -p gpurun_out
OUT=gpurun_out/k_${NAME}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP:-0} -c 1 -f -o $OUT $CMD > gpurun_out/ncu_cmd.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/ncu_cmd.log
ncu -i $OUT.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
ncu -i $OUT.ncu-rep --page source --csv > ${OUT}_src.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("${OUT}_raw.csv")))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
r = rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for k in want:
    if k in idx: print("%-70s %s %s" % (k, r[idx[k]], rows[1][idx[k]]))
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
vals = sorted(((float(r[idx[k]]), k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for k in stall), reverse=True)[:8]
print("stalls per issue:", [(round(v,2),k) for v,k in vals])
PY
python scripts/ncu_hot.py ${OUT}_src.csv ${TOP:-30}
