#!/bin/bash
# A/B: triangulation kernel register caps; c5 FP64 Gram row split
mkdir -p gpurun_out
for mb in 3 4 5 6; do
  IGV_TRI_MINB=$mb timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_triangulate_grp -c 4 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | grep k_triangulate | awk -F'","' -v mb=$mb '{print "tri minb=" mb, $(NF-1), $NF}'
done > gpurun_out/r2r_tri.log 2>&1
cat gpurun_out/r2r_tri.log
for sp in 2 4 6 8; do
  IGV_QR_SPLIT=$sp timeout 600 python bench.py --workload c5 --batch 148 --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 fp64 split=$sp', round(d['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})"
done > gpurun_out/r2r_split.log 2>&1
cat gpurun_out/r2r_split.log
