#!/bin/bash
# attention-map KD gradient formed inside the attention backward (ops.FUSED_ATTN_KD): tests, then GD / itr_step / vqa_step with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "kd_gradient or row_dots or mse" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
for knob in NONE EVLM_NO_FUSED_ATTN_KD; do
  unset EVLM_NO_FUSED_ATTN_KD
  [ $knob != NONE ] && export $knob=1
  for wl in gd itr_step vqa_step; do
    case $wl in itr_step) n=4;; *) n=8;; esac
    python bench.py --workload $wl --steps $n --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/kd_ab_${wl}_$knob.json 2> gpurun_out/kd_ab_${wl}_$knob.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/kd_ab_${wl}_$knob.json") if l.startswith("{")][-1])
    print("%-10s %-22s %.2f ms/step value %.1f e2e %.1f loss %s" % ("$wl", "$knob", d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("final_loss")))
    for k in d["hbm_kernels"]["kernels"][:6]: print("     ", k["kernel"], k["launches"], k["ms"], k["frac"])
except Exception as e:
    print("$wl $knob FAILED", e); print(open("gpurun_out/kd_ab_${wl}_$knob.err").read()[-2500:])
PY
  done
done
