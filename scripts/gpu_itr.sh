#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q --no-header -rf -k "itr or vqa or materialised" > gpurun_out/test_itr.log 2>&1
echo "== itr/vqa tests exit=$? =="; tail -n 12 gpurun_out/test_itr.log
timeout 900 python bench.py --workload itr_step --steps 5 --warmup 3 > gpurun_out/bench_itr_step.json 2> gpurun_out/bench_itr_step.err
rc=$?; echo "== bench itr_step exit=$rc =="; cut -c1-2600 gpurun_out/bench_itr_step.json; tail -n 12 gpurun_out/bench_itr_step.err
if [ $rc -ne 0 ]; then
  timeout 900 python bench.py --workload itr_step --steps 3 --warmup 3 --eager --no-cpu-baseline > gpurun_out/bench_itr_step_eager.json 2> gpurun_out/bench_itr_step_eager.err
  echo "== bench itr_step eager exit=$? =="; cut -c1-1200 gpurun_out/bench_itr_step_eager.json; tail -n 12 gpurun_out/bench_itr_step_eager.err
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_itr_step.csv python bench.py --workload itr_step --profile-step --warmup 3 > gpurun_out/ncu_itr_step.log 2>&1
echo "ncu exit=$?"; python scripts/summarize_launches.py gpurun_out/launches_itr_step.csv | tee gpurun_out/launch_summary_itr_step.txt | head -24
