#!/bin/bash
# A/B of bench flags on one box: usage: VARIANTS="a:--flag1 b:--flag2" bash scripts/gpu_r2_ab_bench.sh
mkdir -p gpurun_out
for r in 1 2; do
for v in ${VARIANTS}; do
  name=${v%%:*}; flags=${v#*:}; flags=${flags//,/ }
  [ "$flags" = "$name" ] && flags=""
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary ${WL_ARGS} $flags > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/ab_$name.json") if l.startswith("{")][-1])
    print("%-12s %7.2f ms/step  value %8.1f  e2e %8.1f  gemm %6.2f ms  loss %s" % ("$name", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["config"].get("final_loss")))
except Exception as e:
    print("$name: parse error", e); print(open("gpurun_out/ab_$name.err").read()[-1500:])
PY
done
done
