#!/bin/bash
# A/B timing of gemm_test cases across library variants (variants/<name>/libevlm_b200.so), interleaved so that both see
# the same box state.  usage: VARIANTS="old new" CASES="fwd_qkv act_fwd_fc1" ROUNDS=3 bash scripts/gpu_gemm_ab.sh
mkdir -p gpurun_out
BIN=efficientvlm_b200/csrc/test/gemm_test
VARIANTS=${VARIANTS:-"old new"}
CASES=${CASES:-"fwd_qkv fwd_fc2 dgrad_fc2 wgrad_fc1 act_fwd_fc1 act_bwd_fc1 res_proj res_fc2 bert_proj bert_out_drop bert_act_fc1 itm_proj fwd_vocab"}
ROUNDS=${ROUNDS:-3}
LOG=gpurun_out/gemm_ab.log
: > $LOG
for c in $CASES; do
  for r in $(seq $ROUNDS); do
    for v in $VARIANTS; do
      out=$(LD_LIBRARY_PATH=$PWD/variants/$v timeout 90 $BIN $c 2>&1)
      echo "$out" | grep -q PASS || echo "  [$v] $c: $(echo "$out" | head -3)" >> $LOG
      t=$(echo "$out" | grep TFLOP | awk '{print $(NF-1)}')
      echo "$c $v $t" >> $LOG
    done
  done
done
python - <<'PY'
import collections, statistics
d = collections.OrderedDict()
for line in open("gpurun_out/gemm_ab.log"):
    p = line.split()
    if len(p) == 3 and not line.startswith(" "):
        try:
            d.setdefault(p[0], collections.OrderedDict()).setdefault(p[1], []).append(float(p[2]))
        except ValueError:
            pass
    elif line.startswith(" "):
        print(line.rstrip())
for c, vs in d.items():
    print("%-16s" % c, "  ".join("%s %7.1f (%s)" % (v, statistics.median(x), " ".join("%.0f" % y for y in x)) for v, x in vs.items()))
PY
