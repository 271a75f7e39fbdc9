#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_precision.py tests/test_gpu_landmarks.py -m gpu -q -s > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
grep -E "precision sweep|passed|failed" gpurun_out/r2f_pytest.log
for prec in fp64 fp32_stack; do
  timeout 400 python bench.py --steps 10 --warmup 3 --workload c5 --batch 148 --precision $prec --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2f_bench_c5_$prec.json 2> gpurun_out/r2f_bench_c5_$prec.err
  timeout 400 python bench.py --steps 20 --warmup 3 --workload c3 --precision $prec --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2f_bench_c3_$prec.json 2> gpurun_out/r2f_bench_c3_$prec.err
done
timeout 400 python bench.py --steps 20 --warmup 3 --workload c2 --batch 64 --precision fp32_stack --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2f_bench_c2b64_fp32.json 2> gpurun_out/r2f_bench_c2b64_fp32.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
