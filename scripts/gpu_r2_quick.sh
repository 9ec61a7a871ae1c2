#!/bin/bash
# kernel tests matching $K (pytest -k) + the GD bench line with the HBM-kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --no-header -rfs -x -k "${K:-mse or kd or gd}" 2>&1 | tail -3
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/quick.json 2>gpurun_out/quick.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/quick.json") if l.startswith("{")][-1])
print("GD: %.2f ms/step value %.1f gemm %.2f ms" % (d["ms_per_step"], d["value"], d["roofline"]["gemm_ms_per_step"]))
for k in d["hbm_kernels"]["kernels"]: print("   ", k)
PY
