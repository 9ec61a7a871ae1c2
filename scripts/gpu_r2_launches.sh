#!/bin/bash
# ncu launch list of ONE steady-state GD step (eager launches between cudaProfilerStart/Stop), summarised per kernel
mkdir -p gpurun_out
TAG=${TAG:-v1}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_ncu_launches_$TAG.csv python bench.py --profile-step --warmup 3 ${BENCH_ARGS} > gpurun_out/ncu_run.log 2>&1
echo "ncu exit=$?"; python scripts/summarize_launches.py gpurun_out/r02_ncu_launches_$TAG.csv | tee gpurun_out/r02_launch_summary_$TAG.txt | head -40
python - <<PY
import csv,collections,re
rows=list(csv.reader(l for l in open('gpurun_out/r02_ncu_launches_$TAG.csv') if not l.startswith('==')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=vi or 'attn' not in r[ki]: continue
    short=re.sub(r'\(.*','',r[ki]).replace('void evlm::','').replace('evlm::','')
    a=agg.setdefault((short,r[gi]),[]); a.append(float(r[vi].replace(',',''))/1e3)
for k,v in agg.items(): print("%-36s grid %-16s n=%2d total %8.1f us : %s"%(k[0],k[1],len(v),sum(v)," ".join("%.0f"%x for x in v)))
PY
