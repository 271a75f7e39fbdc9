#!/bin/bash
mkdir -p gpurun_out
for mw in 12 8; do
IGV_FUSE_MINW=$mw timeout 600 python bench.py --workload c3 --batch 1184 --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 minw=$mw', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})"
done
IGV_FUSE_MINW=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c3_stereo" 2>&1 | tail -2
