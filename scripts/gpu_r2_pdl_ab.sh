#!/bin/bash
# programmatic dependent launch on the decode chain: tests, then caption_infer / vqa_infer / GD with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "small_m or decode or greedy or attention or layernorm or ln" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -x -k "caption or decode or vqa" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
for knob in NONE EVLM_NO_PDL NONE EVLM_NO_PDL; do
  unset EVLM_NO_PDL
  [ $knob != NONE ] && export $knob=1
  for wl in caption_infer vqa_infer; do
    python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/pdl_ab_${wl}_$knob.json 2> gpurun_out/pdl_ab_${wl}_$knob.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/pdl_ab_${wl}_$knob.json") if l.startswith("{")][-1])
    print("%-14s %-12s %.2f ms/step value %.1f e2e %.1f" % ("$wl", "$knob", d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$wl $knob FAILED", e); print(open("gpurun_out/pdl_ab_${wl}_$knob.err").read()[-1500:])
PY
  done
done
unset EVLM_NO_PDL
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/pdl_gd.json 2> gpurun_out/pdl_gd.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/pdl_gd.json") if l.startswith("{")][-1])
print("GD: %.2f ms/step value %.1f gemm %.2f ms" % (d["ms_per_step"], d["value"], d["roofline"]["gemm_ms_per_step"]))
PY
