#!/bin/bash
# attention tests first (new kernels), then the whole GPU suite, then the VQA bench lines + ncu launch list of the pruning step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --no-header -rf -k attention > gpurun_out/test_attn.log 2>&1
rc=$?; echo "== attention tests exit=$rc =="; tail -n 25 gpurun_out/test_attn.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --no-header -rf > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu exit=$? =="; tail -n 12 gpurun_out/test_gpu_all.log
for wl in vqa_step vqa_infer; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "== bench $wl exit=$? =="; tail -c 2500 gpurun_out/bench_$wl.json | cut -c1-420; tail -n 6 gpurun_out/bench_$wl.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --profile-step --warmup 3 > gpurun_out/ncu_$wl.log 2>&1
  echo "ncu $wl exit=$?"; python scripts/summarize_launches.py gpurun_out/launches_$wl.csv | tee gpurun_out/launch_summary_$wl.txt | head -16
done
