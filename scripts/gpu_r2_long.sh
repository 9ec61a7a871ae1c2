#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --no-header -rfs -x -k "attention or vqa or baseline_shape or itr" 2>&1 | tail -4
WORKLOADS="vqa_step itr_step" bash scripts/gpu_r2_workloads.sh 2>&1 | head -3
