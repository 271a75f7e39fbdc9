#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:k_ekf_update<.int.384>' --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02_k_ekf_visual python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/ncu_ekf.log 2>&1
ncu -i gpurun_out/r02_k_ekf_visual.ncu-rep --page raw --csv > gpurun_out/r02_k_ekf_visual_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_ekf.log; ls -la gpurun_out/r02_k_ekf_visual*
