#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/gram_tc_harness tests/cuda/gram_tc_harness.cu > gpurun_out/tc_build.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gram_tc --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02_k_gram_tc /tmp/gram_tc_harness > gpurun_out/tc_ncu.log 2>&1
ncu -i gpurun_out/r02_k_gram_tc.ncu-rep --page raw --csv > gpurun_out/r02_k_gram_tc_raw.csv 2>/dev/null
tail -5 gpurun_out/tc_ncu.log
