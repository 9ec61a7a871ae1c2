#!/bin/bash
# A/B: wide (128x256) vs narrow (128x128) tiles for the split-K weight-gradient GEMMs, same box, interleaved
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "gemm" 2>&1 | tail -2
for r in 1 2; do
for v in wide narrow; do
  if [ $v = narrow ]; then export EVLM_GEMM_NARROW_SPLITK=1; else unset EVLM_GEMM_NARROW_SPLITK; fi
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary --gemm-breakdown > gpurun_out/wg_$v.json 2> gpurun_out/wg_$v.txt
  python - <<PY
import json,re
d = json.loads([l for l in open("gpurun_out/wg_$v.json") if l.startswith("{")][-1])
tot=0
for l in open("gpurun_out/wg_$v.txt"):
    m=re.match(r"\s+\((\d+), (\d+), (\d+), 1, 1\)\s+(\d+)\s+([\d.]+) ms", l)
    if m: tot+=float(m.group(5))
print("%-7s %.2f ms/step  gemm %.2f ms  wgrad (a_mn=b_mn=1) total %.2f ms" % ("$v", d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], tot))
PY
done
done
grep "1, 1)" gpurun_out/wg_wide.txt | head -12; echo ---; grep "1, 1)" gpurun_out/wg_narrow.txt | head -12
