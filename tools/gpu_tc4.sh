#!/bin/bash
mkdir -p gpurun_out
for tol in 1e-13 1e-6 1e-5 1e-4 1e-3; do
  echo "=== IGV_TC_PIVOT_TOL=$tol"
  IGV_TC_PIVOT_TOL=$tol timeout 600 python -m pytest tests/test_gpu_precision.py -m gpu -q -s -k "tf32_gram_tolerance" 2>&1 | grep -E "sweep:|passed|failed" | cut -c1-300
done > gpurun_out/tc_tol.log 2>&1
cat gpurun_out/tc_tol.log
