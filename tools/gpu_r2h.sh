#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "frame_step" > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
# ncu: launch list of the default bench step, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2h_ncu_launch.log 2>&1
for k in k_msckf_features k_triangulate_grp k_ekf_update; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -o gpurun_out/r02_$k -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-c4 > gpurun_out/r2h_ncu_$k.log 2>&1
  ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02_* | head
