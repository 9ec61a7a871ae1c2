#!/bin/bash
# round 2, first GPU pass: every -m gpu test (incl. the two un-calibrated GD fixture tests) with every measured error logged,
# then the default bench line with the same-box eager-PyTorch comparator, then the vqa_infer line.
mkdir -p gpurun_out
rm -f gpurun_out/calib.jsonl
EVLM_UNCALIBRATED_GPU_TESTS=1 EVLM_CALIBRATE_LOG=gpurun_out/calib.jsonl EVLM_CALIBRATE_NOFAIL=1 \
  timeout 1500 python -m pytest tests -m gpu -q --no-header -rfs > gpurun_out/r2_tests.log 2>&1
echo "== tests exit=$? =="; tail -n 25 gpurun_out/r2_tests.log
timeout 900 python bench.py --torch-gpu-baseline > gpurun_out/r2_bench_gd_v0.json 2> gpurun_out/r2_bench_gd_v0.err
echo "== bench gd exit=$? =="; cat gpurun_out/r2_bench_gd_v0.json; tail -n 5 gpurun_out/r2_bench_gd_v0.err
timeout 600 python bench.py --workload vqa_infer > gpurun_out/r2_bench_vqa_infer_v0.json 2> gpurun_out/r2_bench_vqa_infer_v0.err
echo "== bench vqa_infer exit=$? =="; cat gpurun_out/r2_bench_vqa_infer_v0.json; tail -n 5 gpurun_out/r2_bench_vqa_infer_v0.err
