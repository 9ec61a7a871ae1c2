#!/bin/bash
# zero-skip evidence: pruning steps with the kept-column path vs the dense gated path at three shares of exactly-zero gates
mkdir -p gpurun_out
for wl in ${WORKLOADS:-vqa_step itr_step}; do
  for la in ${LOGAS:-0.0 -1.0 -2.2}; do
    for mode in skip dense; do
      extra=""; [ $mode = dense ] && extra="--no-zero-skip"
      timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --gate-loga $la $extra > gpurun_out/skip_${wl}_${la}_${mode}.json 2> gpurun_out/skip_${wl}_${la}_${mode}.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/skip_${wl}_${la}_${mode}.json").read().strip().splitlines()[-1])
    z = d["config"]["zero_skip"]
    print("%-9s loga %5s %-5s: %7.2f ms/step  %8.1f units/s  gemm %6.2f ms  executed %8.0f / dense %8.0f GFLOP  calls %s" % ("$wl", "$la", "$mode", d["ms_per_step"], d["value"], d["roofline"]["gemm_ms_per_step"], z["gemm_gflop_executed"], z["gemm_gflop_dense_gated"], z["ffn_layer_calls"]))
except Exception as e:
    print("$wl $la $mode: parse error", e); print(open("gpurun_out/skip_${wl}_${la}_${mode}.err").read()[-1500:])
PY
    done
  done
done
