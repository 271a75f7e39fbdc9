#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py -m gpu -q -s 2>&1 | grep -E "sweep:|passed|failed|Error|error" | cut -c1-400 > gpurun_out/tc3_pytest.log
cat gpurun_out/tc3_pytest.log
bash tools/gpu_tc_ncu.sh > /dev/null 2>&1
grep "^case\|^timing\|HARNESS" gpurun_out/tc_ncu.log | tail -12
