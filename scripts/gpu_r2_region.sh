#!/bin/bash
# region-batch GD step: fixture parity test + the bench line (graph replay), plus the ncu launch list of one eager step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "region" 2>&1 | tail -5 > gpurun_out/region_tests.log
python bench.py --workload gd_region --steps 8 --warmup 3 > gpurun_out/bench_gd_region.json 2> gpurun_out/bench_gd_region.err
tail -c 3000 gpurun_out/bench_gd_region.err
TAG=region BENCH_ARGS="--workload gd_region --no-cpu-baseline" scripts/gpu_r2_launches.sh > gpurun_out/region_launches.log 2>&1
head -30 gpurun_out/region_launches.log
cat gpurun_out/region_tests.log
head -c 2500 gpurun_out/bench_gd_region.json
