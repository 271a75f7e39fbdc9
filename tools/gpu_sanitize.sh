#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (landmarks, delayed init / new GNSS system, frame graph, FP32 / TF32 modes,
# the tcgen05 harness); slow, so small test selections only
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_landmarks.py tests/test_gpu_parity.py -m gpu -q -x -k "landmark or delayed or gnss or marginalize or all_obs_frames" > gpurun_out/san_1.log 2>&1
echo "rc=$?" >> gpurun_out/san_1.log; tail -4 gpurun_out/san_1.log
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "frame_step or a18 or small_angle or max_dim or triangulate" > gpurun_out/san_2.log 2>&1
echo "rc=$?" >> gpurun_out/san_2.log; tail -4 gpurun_out/san_2.log
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o /tmp/gram_tc_harness tests/cuda/gram_tc_harness.cu
timeout 900 $S --tool memcheck --error-exitcode 9 /tmp/gram_tc_harness > gpurun_out/san_3.log 2>&1
echo "rc=$?" >> gpurun_out/san_3.log; grep -E "ERROR SUMMARY|rc=|HARNESS" gpurun_out/san_3.log | tail -4
grep -E "ERROR SUMMARY" gpurun_out/san_1.log gpurun_out/san_2.log | tail -4
