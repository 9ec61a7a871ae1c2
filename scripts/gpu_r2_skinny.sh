#!/bin/bash
# decode fast paths (small-M GEMM, single-query attention, KV cache, fused token selection): tests, caption / VQA inference A/B, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -x -k "small_m or gemm or decode or greedy or attention" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -x -k "caption or decode or vqa or generat" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
for v in on off; do
  if [ $v = off ]; then export EVLM_GEMM_NO_SKINNY=1 EVLM_ATTN_NO_DECODE=1 EVLM_NO_KV_CACHE=1 EVLM_NO_GREEDY_FUSED=1; else unset EVLM_GEMM_NO_SKINNY EVLM_ATTN_NO_DECODE EVLM_NO_KV_CACHE EVLM_NO_GREEDY_FUSED; fi
  python bench.py --workload caption_infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/caption_fast_$v.json 2> gpurun_out/caption_fast_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/caption_fast_$v.json") if l.startswith("{")][-1])
print("caption_infer fast paths $v: %.2f ms/step %.1f captions/s e2e %.1f launches/step %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"] // d["steps"]))
PY
  python bench.py --workload vqa_infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/vqa_infer_fast_$v.json 2> gpurun_out/vqa_infer_fast_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/vqa_infer_fast_$v.json") if l.startswith("{")][-1])
print("vqa_infer fast paths $v: %.2f ms/step %.1f samples/s" % (d["ms_per_step"], d["value"]))
PY
done
unset EVLM_GEMM_NO_SKINNY EVLM_ATTN_NO_DECODE EVLM_NO_KV_CACHE EVLM_NO_GREEDY_FUSED
TAG=caption_fast BENCH_ARGS="--workload caption_infer --no-cpu-baseline" scripts/gpu_r2_launches.sh 2>/dev/null | head -16
