#!/bin/bash
# attention kernel tests, then the GD bench (no side arms) and the attention launch times of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -rfs -x -k "attention or mse" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -rfs -x 2>&1 | tail -4
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/attn_bench.json 2>gpurun_out/attn_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/attn_bench.json") if l.startswith("{")][-1])
print("GD: %.2f ms/step value %.1f gemm %.2f ms" % (d["ms_per_step"], d["value"], d["roofline"]["gemm_ms_per_step"]))
PY
TAG=${TAG:-attn} bash scripts/gpu_r2_launches.sh 2>&1 | grep -E "attn|total kernel" 
