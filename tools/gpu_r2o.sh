#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_gpu_qr_variants.py tests/test_golden.py tests/test_gpu_precision.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
: > gpurun_out/r2o_table.jsonl
for w in 7 8 10 12; do
  IGV_FEAT_WARPS=$w timeout 600 python bench.py --workload c3 --batch 1184 --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-c4 >> gpurun_out/r2o_table.jsonl 2>> gpurun_out/r2o_table.err
done
for cfg in "c5 148" "c2 1184" "c2 64" "c1 1184"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --batch $2 --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-c4 >> gpurun_out/r2o_table.jsonl 2>> gpurun_out/r2o_table.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2o_table.jsonl'):
    try:
        d=json.loads(l); print(d['config']['workload'][:60], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})
    except Exception as e: print('ERR',e)
PY
