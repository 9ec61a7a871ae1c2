#!/bin/bash
# end-of-round GEMM evidence: isolated timings of the full-size gemm_test cases (plain and fused epilogues) and one
# `ncu --set full` capture each of the two epilogue-bound flavours the step spends most on: the fc2 dgrad with the activation-backward
# epilogue (act_bwd_fc1) and the N = K = 768 projection with the fp32 residual epilogue (res_proj)
mkdir -p gpurun_out
BIN=efficientvlm_b200/csrc/test/gemm_test
LOG=gpurun_out/gemm_cases_end.log
: > $LOG
for c in fwd_qkv fwd_fc1 fwd_fc2 dgrad_fc1 dgrad_fc2 wgrad_fc1 wgrad_proj act_fwd_fc1 act_bwd_fc1 res_proj res_fc2 bert_proj itm_proj bert_out_drop bert_act_fc1 fwd_vocab; do
  timeout 60 $BIN $c 2>&1 | grep TFLOP >> $LOG
done
cat $LOG
for CASE in act_bwd_fc1 res_proj; do
  CASE=$CASE bash scripts/gpu_ncu_gemmtest.sh > gpurun_out/ncu_end_$CASE.txt 2>&1
  head -40 gpurun_out/ncu_end_$CASE.txt
done
