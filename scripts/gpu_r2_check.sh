#!/bin/bash
# strict run of the -m gpu suite (what the driver runs) + the default bench line
mkdir -p gpurun_out
rm -f gpurun_out/calib.jsonl
EVLM_CALIBRATE_LOG=gpurun_out/calib.jsonl timeout 2400 python -m pytest tests -m gpu -q --no-header -rfs -x > gpurun_out/r2_tests_strict.log 2>&1
echo "== tests exit=$? =="; tail -n 15 gpurun_out/r2_tests_strict.log
timeout 1200 python bench.py > gpurun_out/r2_bench_gd_v1.json 2> gpurun_out/r2_bench_gd_v1.err
echo "== bench gd exit=$? =="; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_gd_v1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["frac_of_burst_peak"])
print(d.get("torch_eager_gpu"))
print({k:(v.get("value"),v.get("ms_per_step"),v.get("error")) for k,v in d.get("secondary",{}).items()})
PY
tail -n 5 gpurun_out/r2_bench_gd_v1.err
