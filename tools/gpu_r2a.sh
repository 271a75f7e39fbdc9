#!/bin/bash
# round 2, GPU visit a: tests after the device-guard / knob / chi2 changes + small-batch launch lists
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for B in 1 8 64; do
  timeout 200 python bench.py --steps 50 --warmup 5 --batch $B --no-cpu-baseline --no-latency > gpurun_out/r2a_bench_b$B.json 2> gpurun_out/r2a_bench_b$B.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_b1.csv python bench.py --steps 3 --warmup 3 --batch 1 --no-cpu-baseline --no-latency > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_b8.csv python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline --no-latency > /dev/null 2>&1
tail -3 gpurun_out/r2a_pytest.log
