#!/bin/bash
# What the driver runs at round end, in one call: pytest -m gpu, smoke(), bench (ours + reference arm).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -rf > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu exit=$? =="; tail -n 25 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit=$? =="; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "== bench exit=$? =="; tail -c 3500 gpurun_out/bench_default.json; tail -n 5 gpurun_out/bench_default.err
if [ -n "$WITH_REF" ]; then
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== ref exit=$? =="; tail -c 1200 gpurun_out/bench_ref.json
fi
