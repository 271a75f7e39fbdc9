#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
IGV_TRI_CFG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2g_bench_tri1.json 2> gpurun_out/r2g_bench_tri1.err
python - <<'PY'
import json
for f in ('gpurun_out/r2g_bench.json','gpurun_out/r2g_bench_tri1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d.get('triangulation_extra'), d.get('single_sequence'), d.get('c4_sharded'))
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
