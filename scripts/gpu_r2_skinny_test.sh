#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "small_m" 2>&1 | grep -E "^E  |passed|failed|Error" | head -30
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -x -k "caption" 2>&1 | grep -E "^E  |passed|failed|Error" | head -30
