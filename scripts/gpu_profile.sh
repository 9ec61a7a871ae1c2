#!/bin/bash
# (1) GEMM per-shape breakdown, (2) ncu launch list (gpu__time_duration) of ONE steady-state GD step
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gemm-breakdown > gpurun_out/bench_bd.json 2> gpurun_out/gemm_breakdown.txt
tail -n 45 gpurun_out/gemm_breakdown.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python bench.py --profile-step --warmup 3 > gpurun_out/ncu_run.log 2>&1
echo "ncu exit=$?"; wc -l gpurun_out/launches.csv
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -60
