#!/bin/bash
# two GPUs: the one-process/two-devices test, then the bench at N=2 (c4 leg with the per-frame NCCL gather)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2t_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "two_devices" > gpurun_out/r2t_pytest.log 2>&1
tail -3 gpurun_out/r2t_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2t_bench_n2.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('n_gpus'), round(d['value']), d.get('ms_per_step'), 'e2e', d.get('e2e',{}).get('value'))
        if 'c4_sharded' in d: print('   c4', {k:v for k,v in d['c4_sharded'].items() if k in ('value','per_gpu','ms_per_frame','wall_ms_per_frame','graph_replays','sequences_per_gpu','gather','error')})
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
