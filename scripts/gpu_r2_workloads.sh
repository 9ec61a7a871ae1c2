#!/bin/bash
# every non-headline workload of BASELINE.json once (bench line without the CPU arm), saved under gpurun_out/r02_bench_<workload>.json
mkdir -p gpurun_out
for w in ${WORKLOADS:-vqa_step itr_step vqa_infer caption_infer}; do
  case $w in itr_step) n=4;; *) n=6;; esac
  timeout 900 python bench.py --workload $w --steps $n --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_$w.json") if l.startswith("{")][-1])
    print("%-14s %8.2f ms/step  value %8.1f %s  e2e %8.1f  gemm %6.2f ms (%.0f TFLOP/s, frac %.2f)  mfu %.2f" % ("$w", d["ms_per_step"], d["value"], d["unit"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["model_flops_utilization"]["frac_of_sustained_peak"]))
except Exception as e:
    print("$w: parse error", e); print(open("gpurun_out/r02_bench_$w.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload vqa_infer --materialize --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_vqa_infer_materialized.json 2>/dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_bench_vqa_infer_materialized.json') if l.startswith('{')][-1]); print('vqa_infer --materialize %.2f ms value %.1f' % (d['ms_per_step'], d['value']))"
