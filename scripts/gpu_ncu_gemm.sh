#!/bin/bash
# one `ncu --set full` capture of the dominant kernel: the ViT fc1 (activation epilogue) and fc2 (residual epilogue) GEMMs of layer 0
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -s 3 -c 2 \
  -f -o gpurun_out/gemm_full python bench.py --profile-step --warmup 3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit=$?"; tail -n 5 gpurun_out/ncu_gemm.log; ls -la gpurun_out/gemm_full.ncu-rep
ncu -i gpurun_out/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/gemm_full_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "lts__t_bytes.sum"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print({k: r[idx[k]] for k in want if k in idx})
print("units:", {k: rows[1][idx[k]] for k in want if k in idx})
PY
