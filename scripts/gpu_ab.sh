#!/bin/bash
# attention tests, then bench lines (no CPU baseline) for every workload; extra env knobs per line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --no-header -rf -k "attention" > gpurun_out/test_attn.log 2>&1
rc=$?; echo "== attention tests exit=$rc =="; tail -n 8 gpurun_out/test_attn.log
if [ $rc -ne 0 ]; then exit 1; fi
run() { # name, workload, steps, env...
  name=$1; wl=$2; steps=$3; shift 3
  env "$@" timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "== $name exit=$? =="; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("   %s: %.2f ms/step  value %.1f  e2e %.1f  gemm %.2f ms  clocks %s" % ("$name", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["clocks"]))
except Exception as e:
    print("   parse error", e); print(open("gpurun_out/ab_$name.err").read()[-1500:])
PY
}
for w in ${WORKLOADS:-gd vqa_step itr_step vqa_infer}; do
  case $w in gd) n=8;; itr_step) n=4;; *) n=5;; esac
  run $w $w $n A=1
done
