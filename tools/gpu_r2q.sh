#!/bin/bash
# round 2: tcgen05 Gram mode through the library (tolerance tests), ill-conditioned stacks, bench of the modes at c5 / c3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_precision.py tests/test_gpu_illcond.py -m gpu -q -s > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
grep -E "sweep:|weak-pivot|passed|failed|rc=" gpurun_out/r2q_pytest.log | tail -30
: > gpurun_out/r2q_table.jsonl
for cfg in "c5 148 fp64" "c5 148 fp32_stack" "c5 148 tf32_gram" "c5 32 tf32_gram" "c3 1184 tf32_gram"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --batch $2 --precision $3 --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-c4 >> gpurun_out/r2q_table.jsonl 2>> gpurun_out/r2q_table.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2q_table.jsonl'):
    try:
        d=json.loads(l); print(d['config']['workload'][:40], d['dtype'][:40], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})
    except Exception as e: print('ERR',e)
PY
