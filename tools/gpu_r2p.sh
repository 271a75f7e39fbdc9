#!/bin/bash
# A/B of the stereo per-track kernel: 384-thread bound (168 registers) vs 512-thread bound (128 registers, spills)
mkdir -p gpurun_out
: > gpurun_out/r2p_table.txt
run() {  # label, env...
  label=$1; shift
  out=$(env "$@" timeout 600 python bench.py --workload c3 --batch 1184 --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-c4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['kernel_ms_per_step']['features'],3))")
  echo "$label $out" >> gpurun_out/r2p_table.txt
}
run "t384 W12 ps1" IGV_FEAT_WARPS=12
run "t384 W12 ps0" IGV_FEAT_WARPS=12 IGV_FEAT_PS=0
cp ingvio_b200/lib/libingvio_b200.so /tmp/keep.so
cp ingvio_b200/lib/libingvio_b200_s512.so ingvio_b200/lib/libingvio_b200.so
run "t512 W12 ps1" IGV_FEAT_WARPS=12
run "t512 W12 ps0" IGV_FEAT_WARPS=12 IGV_FEAT_PS=0
run "t512 W14 ps0" IGV_FEAT_WARPS=14 IGV_FEAT_PS=0
run "t512 W16 ps0" IGV_FEAT_WARPS=16 IGV_FEAT_PS=0
cp /tmp/keep.so ingvio_b200/lib/libingvio_b200.so
cat gpurun_out/r2p_table.txt
