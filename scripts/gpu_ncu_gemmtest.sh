#!/bin/bash
# usage: CASE=act_fwd_fc1 VARIANT=v1 bash scripts/gpu_ncu_gemmtest.sh
# one `ncu --set full` capture (with SASS/source sampling) of a gemm_test case; prints key metrics, stall reasons, pipe use and
# the hottest source lines by warp-stall samples.
mkdir -p gpurun_out
BIN=efficientvlm_b200/csrc/test/gemm_test
V=${VARIANT:-}
[ -n "$V" ] && export LD_LIBRARY_PATH=$PWD/variants/$V
OUT=gpurun_out/gt_${CASE}_${V:-cur}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f -o $OUT $BIN $CASE > gpurun_out/ncu_gt.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/ncu_gt.log
ncu -i $OUT.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
ncu -i $OUT.ncu-rep --page source --csv > ${OUT}_src.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("${OUT}_raw.csv")))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
r = rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active"]
for k in want:
    if k in idx: print("%-70s %s %s" % (k, r[idx[k]], rows[1][idx[k]]))
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
vals = sorted(((float(r[idx[k]]), k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for k in stall), reverse=True)[:8]
print("stalls per issue:", [(round(v,2),k) for v,k in vals])
pipes = sorted(((float(r[idx[h]]), h) for h in hdr if h.startswith("sm__inst_executed_pipe_") and h.endswith("pct_of_peak_sustained_active") and r[idx[h]] not in ("", "n/a")), reverse=True)[:8]
print("pipes:", [(round(v,1), h.split("sm__inst_executed_pipe_")[1].split(".")[0]) for v, h in pipes])
# hottest SASS lines
src = list(csv.reader(open("${OUT}_src.csv")))
h = src[0]; ix = {n: i for i, n in enumerate(h)}
samp = next((n for n in h if n.startswith("# Samples") or n == "Warp Stall Sampling (All Samples)" or "Sampling (All" in n), None)
sass = next((n for n in h if n in ("Source", "SASS")), h[1])
print("columns:", [n for n in h][:14])
if samp:
    body = [x for x in src[1:] if len(x) == len(h)]
    def f(x):
        try: return float(x[ix[samp]])
        except ValueError: return 0.0
    tot = sum(f(x) for x in body) or 1
    for x in sorted(body, key=f, reverse=True)[:28]:
        print("%6.2f%%  %s" % (100 * f(x) / tot, x[ix[sass]][:130]))
PY
