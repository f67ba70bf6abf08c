#!/bin/bash
# Runs every GEMM self-test case in its own process (a trapped kernel kills only that case), then the perf probe.
mkdir -p gpurun_out
BIN=build/gemm_selftest
N=$($BIN)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/selftest.log 2>&1
for i in $(seq 0 $((N-1))); do
  timeout 60 $BIN $i >> gpurun_out/selftest.log 2>&1
  echo "case $i exit $?" >> gpurun_out/selftest.log
done
timeout 120 $BIN perf >> gpurun_out/selftest.log 2>&1
echo "perf exit $?" >> gpurun_out/selftest.log
cat gpurun_out/selftest.log
