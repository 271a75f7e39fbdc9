#!/bin/bash
for fc in 0 1; do for b in 1 8 1184; do
IGV_FACTOR_CFG=$fc timeout 600 python bench.py --workload c2 --batch $b --steps 30 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('factor_cfg=$fc B=$b', round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0.02})"
done; done
