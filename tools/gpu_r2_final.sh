#!/bin/bash
# round 2, final measurement visit: every GPU test, the default bench line + reference arm, the config table, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02f_smi.txt 2>&1
nproc > gpurun_out/r02f_nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r02f_nproc.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
tail -4 gpurun_out/r02f_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
: > gpurun_out/r02f_config_table.jsonl
for cfg in "c1 1184 fp64" "c3 1184 fp64" "c5 148 fp64" "c5 32 fp64" "c5 148 fp32_stack" "c5 148 tf32_gram" "c5 32 tf32_gram" "c2 64 fp64" "c2 8 fp64" "c2 1 fp64"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --batch $2 --precision $3 --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-c4 >> gpurun_out/r02f_config_table.jsonl 2>> gpurun_out/r02f_config_table.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r02f_ncu_launch.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_triangulate_grp --launch-count 1 -f -o gpurun_out/r02_k_triangulate_grp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r02f_ncu_tri.log 2>&1
ncu -i gpurun_out/r02_k_triangulate_grp.ncu-rep --page raw --csv > gpurun_out/r02_k_triangulate_grp_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/r02f_bench.json','gpurun_out/r02f_bench_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d.get('ms_per_step'), 'e2e', d.get('e2e',{}).get('value'), 'roofline', {k:d.get('roofline',{}).get(k) for k in ('bound','achieved','peak','frac','traffic')}, 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
for l in open('gpurun_out/r02f_config_table.jsonl'):
    try:
        d=json.loads(l); print(d['config']['workload'][:60], d['dtype'][:24], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.05})
    except Exception as e: print('ERR',e)
PY
