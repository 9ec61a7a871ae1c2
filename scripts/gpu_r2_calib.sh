#!/bin/bash
# every -m gpu test with the reference-noise bound, all measured errors logged (no failure on tolerance)
mkdir -p gpurun_out
rm -f gpurun_out/calib.jsonl
EVLM_CALIBRATE_LOG=gpurun_out/calib.jsonl EVLM_CALIBRATE_NOFAIL=1 \
  timeout 2400 python -m pytest tests -m gpu -q --no-header -rfs -x > gpurun_out/r2_tests.log 2>&1
echo "== tests exit=$? =="; tail -n 40 gpurun_out/r2_tests.log
