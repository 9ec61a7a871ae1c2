#!/bin/bash
# Runs every gemm_test case in its own process (a trap in one case must not take the others down).
mkdir -p gpurun_out
LOG=gpurun_out/gemm_test.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
BIN=efficientvlm_b200/csrc/test/gemm_test
for c in $($BIN); do
  timeout 90 $BIN $c >> $LOG 2>&1
  echo "  exit=$? ($c)" >> $LOG
done
tail -n 80 $LOG
