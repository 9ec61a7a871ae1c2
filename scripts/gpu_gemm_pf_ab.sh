#!/bin/bash
# A/B of the producer's L2 prefetch distance on gemm_test cases (same box, interleaved)
BIN=efficientvlm_b200/csrc/test/gemm_test
for c in ${CASES:-fwd_qkv fwd_fc1 fwd_fc2 dgrad_fc1 dgrad_fc2 wgrad_fc1 wgrad_proj act_fwd_fc1 act_bwd_fc1 res_proj res_fc2 bert_qkv}; do
  line="$c"
  for pf in ${PFS:-0 4 8 16}; do
    best=0
    for r in 1 2; do
      out=$(EVLM_GEMM_PF_AHEAD=$pf timeout 90 $BIN $c 2>&1)
      echo "$out" | grep -q PASS || line="$line FAIL($pf)"
      t=$(echo "$out" | grep TFLOP | awk '{print $(NF-1)}')
      best=$(python -c "print(max($best, float('${t:-0}')))")
    done
    line="$line  pf$pf=$best"
  done
  echo "$line"
done
