#!/bin/bash
# caption_infer with each decode fast path switched off in turn: device-timed value and end-to-end value
mkdir -p gpurun_out
for knob in NONE EVLM_GEMM_NO_SKINNY EVLM_ATTN_NO_DECODE EVLM_NO_KV_CACHE EVLM_NO_GREEDY_FUSED NONE; do
  unset EVLM_GEMM_NO_SKINNY EVLM_ATTN_NO_DECODE EVLM_NO_KV_CACHE EVLM_NO_GREEDY_FUSED
  [ $knob != NONE ] && export $knob=1
  python bench.py --workload caption_infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/cap_ab_$knob.json 2> gpurun_out/cap_ab_$knob.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/cap_ab_$knob.json") if l.startswith("{")][-1])
print("%-22s %.2f ms/step value %.1f e2e %.1f" % ("$knob", d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
done
