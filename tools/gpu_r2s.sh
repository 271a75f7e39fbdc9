#!/bin/bash
mkdir -p gpurun_out
for p in fp64 tf32_gram; do
timeout 600 python bench.py --workload c5 --batch 148 --precision $p --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $p', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})"
done
timeout 600 python bench.py --workload c5 --batch 32 --precision tf32_gram --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 B=32 tf32', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_msckf_features --launch-skip 40 --launch-count 1 -f -o gpurun_out/r02_c5_features python bench.py --workload c5 --batch 148 --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2s_ncu.log 2>&1
ncu -i gpurun_out/r02_c5_features.ncu-rep --page raw --csv > gpurun_out/r02_c5_features_raw.csv 2>/dev/null
tail -3 gpurun_out/r2s_ncu.log
