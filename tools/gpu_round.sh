#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, full captures of the top kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests|notests] [full|nofull]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-run}
TESTS=${2:-tests}
FULL=${3:-full}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/${TAG}_smi.txt

if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
  tail -1 $OUT/${TAG}_smoke.log
fi

timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json | cut -c1-400

NCU="ncu --clock-control none"
BARGS="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches.csv python $BARGS > $OUT/${TAG}_ncu_launch.log 2>&1
if [ "$FULL" = "full" ]; then
  # track-table kernels: launch list of the extra leg
  timeout 300 $NCU --metrics gpu__time_duration.sum -k regex:k_trk -c 200 --csv --log-file $OUT/${TAG}_trk_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_trk.log 2>&1
  for K in k_msckf_features k_ekf_update; do
    timeout 400 $NCU --set full --import-source on -k regex:$K -s 14 -c 2 -f -o $OUT/${TAG}_$K python $BARGS > $OUT/${TAG}_ncu_$K.log 2>&1
    ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_raw.csv 2>/dev/null
  done
fi
ls -la $OUT | tail -30
