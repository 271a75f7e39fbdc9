#!/bin/bash
# full ncu capture of one kernel inside the bench: bash tools/ncu_one.sh <tag> <kernel-regex> [skip] [count]
TAG=$1; K=$2; SKIP=${3:-14}; CNT=${4:-1}
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_ncu.log
