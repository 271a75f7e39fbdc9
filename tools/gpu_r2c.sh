#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_gpu_qr_variants.py tests/test_cpp_updaters.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
for c in 0 1; do
  IGV_FEAT_CONST=$c timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2c_bench_const$c.json 2> gpurun_out/r2c_bench_const$c.err
  IGV_FEAT_CONST=$c timeout 300 python bench.py --steps 20 --warmup 5 --workload c3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2c_bench_c3_const$c.json 2> gpurun_out/r2c_bench_c3_const$c.err
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}, d.get('c4_sharded'))
    except Exception as e: print(f,'ERR',e)
PY
