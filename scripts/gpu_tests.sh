#!/bin/bash
# pytest -m gpu with per-file logs; each file in its own process so one CUDA fault does not poison the rest.
mkdir -p gpurun_out
for f in tests/test_gpu_kernels.py tests/test_gpu_models.py; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --no-header -rf ${PYTEST_EXTRA} > gpurun_out/$name.log 2>&1
  echo "== $f exit=$? ==" | tee -a gpurun_out/$name.log
  tail -n 40 gpurun_out/$name.log
done
