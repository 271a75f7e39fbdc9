#!/bin/bash
for cfg in "c2 1184" "c2 1" "c5 148"; do
set -- $cfg
timeout 600 python bench.py --workload $1 --batch $2 --steps 50 --warmup 5 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 B=$2', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0.02})"
done
