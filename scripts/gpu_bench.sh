#!/bin/bash
# first-light bench: ours (N=1) with a short run, then the reference arm; logs under gpurun_out/
mkdir -p gpurun_out
STEPS=${STEPS:-5}; WARM=${WARM:-3}; EXTRA=${EXTRA:-}
timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup $WARM $EXTRA > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench_ours.json; tail -n 15 gpurun_out/bench_ours.err
