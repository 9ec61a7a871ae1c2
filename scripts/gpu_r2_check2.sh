#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/calib.jsonl
EVLM_CALIBRATE_LOG=gpurun_out/calib.jsonl timeout 2400 python -m pytest tests -m gpu -q --no-header -rfs > gpurun_out/r2_tests_strict.log 2>&1
echo "== tests exit=$? =="; grep -E "^E  |FAILED|passed|failed" gpurun_out/r2_tests_strict.log | tail -n 15
timeout 1200 python bench.py > gpurun_out/r2_bench_gd_v2.json 2> gpurun_out/r2_bench_gd_v2.err
echo "== bench gd exit=$? =="; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_gd_v2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "gemm frac", d["roofline"]["frac"], d["roofline"]["frac_of_burst_peak"], "gemm ms", d["roofline"]["gemm_ms_per_step"])
print("eager bf16", d.get("torch_eager_gpu",{}).get("bf16_autocast",{}).get("value"), "ratio", d.get("torch_eager_gpu",{}).get("ours_over_bf16_autocast"))
print({k:(v.get("value"),v.get("ms_per_step"),v.get("error")) for k,v in d.get("secondary",{}).items()})
for k in d["hbm_kernels"]["kernels"]: print("   ", k)
PY
tail -n 3 gpurun_out/r2_bench_gd_v2.err
timeout 600 python bench.py --gemm-breakdown --no-cpu-baseline --no-torch-gpu-baseline --no-secondary > /dev/null 2> gpurun_out/r2_gemm_breakdown_v2.txt; head -45 gpurun_out/r2_gemm_breakdown_v2.txt
