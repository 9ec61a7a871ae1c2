#!/bin/bash
# usage: KREGEX=attn_bwd_tc SKIP=0 COUNT=2 [WL=vqa_step] bash scripts/gpu_ncu_kernel.sh  -> gpurun_out/k_<regex>_<wl>.ncu-rep + raw csv
mkdir -p gpurun_out
WL=${WL:-gd}
OUT=gpurun_out/k_${KREGEX}_${WL}
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${KREGEX} -s ${SKIP:-0} -c ${COUNT:-2} \
  -f -o ${OUT} python bench.py --workload ${WL} --profile-step --warmup 3 > gpurun_out/ncu_k.log 2>&1
echo "ncu exit=$?"; tail -n 3 gpurun_out/ncu_k.log
ncu -i ${OUT}.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
ncu -i ${OUT}.ncu-rep --page source --csv > ${OUT}_src.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("${OUT}_raw.csv")))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sectors_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum"]
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print({k: r[idx[k]] for k in want if k in idx})
    vals = sorted(((float(r[idx[k]]), k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for k in stall), reverse=True)[:7]
    print("   stalls:", [(round(v,2),k) for v,k in vals])
PY
