#!/bin/bash
# one bench line per BASELINE.json config (GPU arm only, short)
mkdir -p gpurun_out
TAG=${1:-cfg}
run() { echo "== $*" | tee -a gpurun_out/${TAG}_table.log; python bench.py --steps 10 --warmup 3 --no-latency "$@" 2>>gpurun_out/${TAG}_table.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(json.dumps({'workload': d['config']['workload'], 'value': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'kernels': d['kernel_ms_per_step'], 'cpu': d.get('cpu_baseline', {}).get('value'), 'path': d['roofline']['kernel'][:40]}))
" | tee -a gpurun_out/${TAG}_table.log; }
run --workload c1 --batch 1184 --cpu-sample-frames 400
run --workload c3 --batch 1184 --cpu-sample-frames 100
run --workload c2 --batch 8 --no-cpu-baseline
run --workload c2 --batch 64 --no-cpu-baseline
run --workload c5 --batch 32 --cpu-sample-frames 10
run --workload c5 --batch 148 --no-cpu-baseline
