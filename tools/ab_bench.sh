#!/bin/bash
# A/B of env-selected kernel variants: bash tools/ab_bench.sh <tag> "VAR=val" "VAR=val2" ...
TAG=$1; shift
mkdir -p gpurun_out
for kv in "$@"; do
  echo "== $kv" | tee -a gpurun_out/${TAG}_ab.log
  env $kv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'kernels': d['kernel_ms_per_step']}))
" | tee -a gpurun_out/${TAG}_ab.log
done
