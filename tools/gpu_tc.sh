#!/bin/bash
# tcgen05 Gram kernel bring-up: stand-alone harness (bounded waits inside the kernel; the whole run under a timeout)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/gram_tc_harness tests/cuda/gram_tc_harness.cu > gpurun_out/tc_build.log 2>&1
timeout 300 /tmp/gram_tc_harness > gpurun_out/tc_harness.log 2>&1
echo "harness rc=$?" >> gpurun_out/tc_harness.log
for v in $TC_VARIANTS; do
  echo "=== variant $v" >> gpurun_out/tc_harness.log
  timeout 60 /tmp/gram_tc_harness $v >> gpurun_out/tc_harness.log 2>&1
done
grep -v "^  b=" gpurun_out/tc_harness.log | tail -60
