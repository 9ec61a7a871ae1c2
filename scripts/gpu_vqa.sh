#!/bin/bash
# VQA workloads on the B200: parity tests, bench lines (graph, falling back to eager), A/B against the tiled attention, ncu launch lists.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -rf > gpurun_out/test_gpu_all.log 2>&1
echo "== pytest -m gpu exit=$? =="; tail -n 30 gpurun_out/test_gpu_all.log
for wl in vqa_step vqa_infer; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 ${BENCH_EXTRA} > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  rc=$?; echo "== bench $wl exit=$rc =="; tail -c 2500 gpurun_out/bench_$wl.json; tail -n 12 gpurun_out/bench_$wl.err
  if [ $rc -ne 0 ]; then
    timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --eager --no-cpu-baseline > gpurun_out/bench_${wl}_eager.json 2> gpurun_out/bench_${wl}_eager.err
    echo "== bench $wl eager exit=$? =="; tail -c 2500 gpurun_out/bench_${wl}_eager.json; tail -n 12 gpurun_out/bench_${wl}_eager.err
  fi
  EVLM_ATTN_NO_LONG=1 timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_tiled.json 2> gpurun_out/bench_${wl}_tiled.err
  echo "== bench $wl (tiled mma.sync attention for Lk>256) exit=$? =="; tail -c 2500 gpurun_out/bench_${wl}_tiled.json | cut -c1-260; tail -n 5 gpurun_out/bench_${wl}_tiled.err
  if [ -z "$NO_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --profile-step --warmup 3 > gpurun_out/ncu_$wl.log 2>&1
  echo "ncu $wl exit=$?"; python scripts/summarize_launches.py gpurun_out/launches_$wl.csv | tee gpurun_out/launch_summary_$wl.txt | head -32
  fi
done
