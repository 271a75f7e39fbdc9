#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0.02})"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_propagate --launch-skip 12 --launch-count 1 -f -o gpurun_out/r02_k_propagate python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/ncu_prop.log 2>&1
ncu -i gpurun_out/r02_k_propagate.ncu-rep --page raw --csv > gpurun_out/r02_k_propagate_raw.csv 2>/dev/null
ls -la gpurun_out/r02_k_propagate_raw.csv
