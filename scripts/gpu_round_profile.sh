#!/bin/bash
# Round evidence in one call: GPU tests, smoke, GD bench line (default invocation), VQA inference bench line, ncu launch list of one
# GD step and one `ncu --set full` capture of the dominant GEMM (ViT fc1 / fc2 of the student's first layer).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -rf > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu exit=$? =="; tail -n 12 gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit=$? =="; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --gemm-breakdown > gpurun_out/bench_default.json 2> gpurun_out/gemm_breakdown.txt
echo "== bench exit=$? =="; tail -c 3000 gpurun_out/bench_default.json | cut -c1-3000; head -n 14 gpurun_out/gemm_breakdown.txt
timeout 600 python bench.py --workload vqa_infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vqa_infer.json 2> gpurun_out/bench_vqa_infer.err
echo "== bench vqa_infer exit=$? =="; cut -c1-330 gpurun_out/bench_vqa_infer.json; tail -n 4 gpurun_out/bench_vqa_infer.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches.csv python bench.py --profile-step --warmup 3 > gpurun_out/ncu_run.log 2>&1
echo "ncu exit=$?"; python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -24
bash scripts/gpu_ncu_gemm.sh 2>&1 | tail -8
