#!/bin/bash
# 8-GPU GD step: overlapped vs blocking gradient exchange (one box), then the 1-GPU line on the same box
mkdir -p gpurun_out
for mode in overlap overlap1 blocking; do
  extra=""; [ $mode = blocking ] && extra="--no-overlap"
  unset EVLM_OVERLAP_STAGES; [ $mode = overlap1 ] && export EVLM_OVERLAP_STAGES=1
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary $extra > gpurun_out/r02_bench_gd_8gpu_$mode.json 2> gpurun_out/r02_bench_gd_8gpu_$mode.err
  echo "== $mode exit=$? =="
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_gd_8gpu_$mode.json") if l.startswith("{")][-1])
    print("$mode: %.2f ms/step value %.1f e2e %.1f comm %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {k: round(d["comm"][k], 2) for k in ("grad_allreduce_ms", "algbw_gbps")}))
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_bench_gd_8gpu_$mode.err").read()[-2000:])
PY
done
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/r02_bench_gd_1gpu_samebox8.json 2>/dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_gd_1gpu_samebox8.json') if l.startswith('{')][-1]); print('1 GPU same box: %.2f ms/step value %.1f' % (d['ms_per_step'], d['value']))"
