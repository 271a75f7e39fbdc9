#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msckf_features -s 14 -c 1 -o gpurun_out/r02_c3_features2 -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2n_ncu_c3.log 2>&1
ncu -i gpurun_out/r02_c3_features2.ncu-rep --page raw --csv > gpurun_out/r02_c3_features2_raw.csv 2>/dev/null
ls -la gpurun_out/r02_c3_features2*
