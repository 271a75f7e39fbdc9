#!/bin/bash
mkdir -p gpurun_out
for cfg in "384 384" "128 384" "256 384" "128 256" "256 256"; do
  set -- $cfg
  IGV_EKF_T_SMALL=$1 IGV_EKF_T_BIG=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2k_bench_$1_$2.json 2> gpurun_out/r2k_bench_$1_$2.err
  IGV_EKF_T_SMALL=$1 IGV_EKF_T_BIG=$2 timeout 300 python bench.py --steps 100 --warmup 10 --batch 8 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2k_bench_b8_$1_$2.json 2> gpurun_out/r2k_bench_b8_$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if k in ('ekf','features','propagate','qr')})
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
