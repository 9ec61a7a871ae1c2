#!/bin/bash
# gemm_test: every case once for correctness with the double-buffered (DB) epilogue, then A/B timing DB vs EVLM_GEMM_NO_DB=1 (interleaved)
mkdir -p gpurun_out
BIN=efficientvlm_b200/csrc/test/gemm_test
LOG=gpurun_out/gemm_db_ab.log
: > $LOG
for c in $($BIN); do
  out=$(timeout 90 $BIN $c 2>&1); rc=$?
  echo "$out" | grep -q PASS || { echo "FAIL[$rc] $c: $(echo "$out" | tail -3)"; }
done
CASES=${CASES:-"act_fwd_fc1 act_bwd_fc1 res_proj res_fc2 bert_out_drop bert_act_fc1 epi_fwd_vit"}
for c in $CASES; do
  for r in 1 2 3; do
    for v in db nodb; do
      if [ $v = nodb ]; then out=$(EVLM_GEMM_NO_DB=1 timeout 90 $BIN $c 2>&1); else out=$(timeout 90 $BIN $c 2>&1); fi
      t=$(echo "$out" | grep TFLOP | awk '{print $(NF-1)}')
      echo "$c $v $t" >> $LOG
    done
  done
done
python - <<'PY'
import collections, statistics
d = collections.OrderedDict()
for line in open("gpurun_out/gemm_db_ab.log"):
    p = line.split()
    if len(p) == 3:
        try: d.setdefault(p[0], collections.OrderedDict()).setdefault(p[1], []).append(float(p[2]))
        except ValueError: pass
for c, vs in d.items():
    print("%-16s" % c, "  ".join("%s %7.1f (%s)" % (v, statistics.median(x), " ".join("%.0f" % y for y in x)) for v, x in vs.items()))
PY
