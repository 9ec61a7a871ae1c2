#!/bin/bash
# A/B of the GEMM epilogue side-operand prefetch: variants/base/libevlm_b200.so (before) vs the in-tree library (after) on the
# full-size gemm_test cases (each case also verifies the result against its reference), then the GPU suite, smoke and the bench.
mkdir -p gpurun_out
BIN=efficientvlm_b200/csrc/test/gemm_test
LOG=gpurun_out/epilogue_prefetch_ab.log
: > $LOG
CASES="res_proj act_bwd_fc1 res_fc2 act_fwd_fc1 bert_out_drop bert_act_fc1 bert_proj itm_proj fwd_fc1 fwd_fc2 dgrad_fc1 wgrad_proj fwd_vocab"
for c in $CASES; do
  a=$(LD_LIBRARY_PATH=$PWD/variants/base timeout 60 $BIN $c 2>&1 | grep TFLOP)
  b=$(timeout 60 $BIN $c 2>&1 | grep -E "TFLOP|FAIL" | tr '\n' ' ')
  echo "base $a | new $b" >> $LOG
done
for c in fwd_small fwd_ragged epi_fwd_vit epi_fwd_bert epi_bwd_pre epi_bwd_post wgrad_split wgrad_ragged dgrad_small; do
  timeout 60 $BIN $c 2>&1 | grep -E "PASS|FAIL" >> $LOG
done
cat $LOG
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/final_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/final_smoke.log
timeout 120 python bench.py > gpurun_out/final_bench_v10.json 2> gpurun_out/final_bench_v10.err; echo bench rc=$?; cut -c1-330 gpurun_out/final_bench_v10.json
