#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/gram_tc_harness tests/cuda/gram_tc_harness.cu > gpurun_out/tc_build.log 2>&1
for d in 8 4 2 1; do
  echo "=== TC_DRAIN=$d"
  TC_DRAIN=$d timeout 300 /tmp/gram_tc_harness | grep "^case\|^timing\|HARNESS"
done > gpurun_out/tc_drain.log 2>&1
cat gpurun_out/tc_drain.log
for d in 8 2 1; do
  echo "=== IGV_TC_DRAIN=$d"
  IGV_TC_DRAIN=$d timeout 600 python -m pytest tests/test_gpu_precision.py -m gpu -q -s -k tf32 2>&1 | grep -E "sweep:|passed|failed"
done > gpurun_out/tc_drain_tol.log 2>&1
cat gpurun_out/tc_drain_tol.log
