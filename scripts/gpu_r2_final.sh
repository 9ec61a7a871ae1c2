#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke(), the reference arm, the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_gpu_tests_final.log 2>&1; echo "== pytest -m gpu exit=$? =="; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "== reference arm exit=$? =="; head -c 400 gpurun_out/r02_bench_reference_arm.json; echo
timeout 1500 python bench.py > gpurun_out/r02_bench_gd_final.json 2> gpurun_out/r02_bench_gd_final.err; echo "== bench exit=$? =="
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_gd_final.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "gemm frac", round(d["roofline"]["frac"],3), round(d["roofline"]["frac_of_burst_peak"],3), "gemm ms", round(d["roofline"]["gemm_ms_per_step"],2), d["clocks"])
print("eager bf16", d.get("torch_eager_gpu",{}).get("bf16_autocast",{}).get("value"), "ratio", d.get("torch_eager_gpu",{}).get("ours_over_bf16_autocast"), "cpu", d["cpu_baseline"]["value"])
print({k:(v.get("value"),v.get("ms_per_step"),v.get("error")) for k,v in d.get("secondary",{}).items()})
PY
