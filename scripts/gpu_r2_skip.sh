#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -rfs -x -k "compact or limits or zero_skip or gemm" > gpurun_out/r2_skip_tests.log 2>&1
echo "== skip tests exit=$? =="; tail -n 30 gpurun_out/r2_skip_tests.log
EVLM_CALIBRATE_LOG=gpurun_out/calib.jsonl timeout 2400 python -m pytest tests -m gpu -q --no-header -rfs > gpurun_out/r2_tests_strict.log 2>&1
echo "== all tests exit=$? =="; tail -n 30 gpurun_out/r2_tests_strict.log
