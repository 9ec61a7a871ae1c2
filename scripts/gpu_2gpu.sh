#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_gd_2gpu.json 2> gpurun_out/bench_gd_2gpu.err
echo "== 2-GPU gd exit=$? =="; grep '^{' gpurun_out/bench_gd_2gpu.json | cut -c1-3000; tail -n 5 gpurun_out/bench_gd_2gpu.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gd_1gpu_same_box.json 2> gpurun_out/bench_gd_1gpu_same_box.err
echo "== 1-GPU gd (same box) exit=$? =="; cut -c1-2600 gpurun_out/bench_gd_1gpu_same_box.json; tail -n 3 gpurun_out/bench_gd_1gpu_same_box.err
