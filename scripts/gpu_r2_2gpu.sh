#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_distributed.py -m gpu -q --no-header -rfs -x 2>&1 | tail -15 | tee gpurun_out/r02_ddp_parity_2gpu.log
for mode in overlap blocking; do
  extra=""; [ $mode = blocking ] && extra="--no-overlap"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 $extra > gpurun_out/r02_bench_gd_2gpu_$mode.json 2> gpurun_out/r02_bench_gd_2gpu_$mode.err
  echo "== $mode exit=$? =="
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_gd_2gpu_$mode.json") if l.startswith("{")][-1])
    print("$mode: %.2f ms/step value %.1f e2e %.1f comm %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {k: d["comm"][k] for k in ("grad_allreduce_ms", "algbw_gbps")}))
except Exception as e:
    print("parse error", e); print(open("gpurun_out/r02_bench_gd_2gpu_$mode.err").read()[-2000:])
PY
done
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > gpurun_out/r02_bench_gd_1gpu_samebox.json 2>/dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_gd_1gpu_samebox.json') if l.startswith('{')][-1]); print('1 GPU same box: %.2f ms/step value %.1f' % (d['ms_per_step'], d['value']))"
