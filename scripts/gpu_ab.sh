#!/bin/bash
# attention tests, then A/B bench lines (no CPU baseline) for every workload; knobs come from the environment of each line
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
run gd gd 8 A=1
run vqa_step vqa_step 5 A=1
run itr_step itr_step 4 A=1
run itr_step_kps2 itr_step 4 EVLM_BWD_KT_PER_CTA=2
run vqa_step_kps1 vqa_step 5 EVLM_BWD_KT_PER_CTA=1
run vqa_infer vqa_infer 5 A=1
