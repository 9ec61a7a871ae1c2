#!/bin/bash
# timing (and PASS/FAIL) of selected gemm_test cases, 3 runs each
BIN=efficientvlm_b200/csrc/test/gemm_test
for c in ${CASES:-act_fwd_fc1 act_bwd_fc1 epi_fwd_vit epi_bwd_pre bert_act_fc1}; do
  for r in 1 2 3; do
    out=$(timeout 90 $BIN $c 2>&1)
    echo "$c $(echo "$out" | grep -q PASS && echo PASS || echo FAIL) $(echo "$out" | grep TFLOP | awk '{print $(NF-1)}') $(echo "$out" | grep -i "err" | head -1)"
  done
done
