#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_golden.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
for t in 0 1; do
  IGV_PROP_TMA=$t timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2i_bench_tma$t.json 2> gpurun_out/r2i_bench_tma$t.err
  IGV_PROP_TMA=$t timeout 300 python bench.py --steps 100 --warmup 10 --batch 8 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2i_bench_b8_tma$t.json 2> gpurun_out/r2i_bench_b8_tma$t.err
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_full.json 2> gpurun_out/r2i_bench_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items()})
        if 'single_sequence' in d: print('   single', d['single_sequence'])
        if 'c4_sharded' in d: print('   c4', {k:v for k,v in d['c4_sharded'].items() if k in ('value','ms_per_frame','wall_ms_per_frame','graph_replays','gpu_launches_per_frame','error')})
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
