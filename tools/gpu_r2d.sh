#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_r2.py tests/test_cpp_gnss.py tests/test_gpu_gnss_residuals.py tests/test_cpp_host_mirror.py tests/test_capi_cpu.py -m gpu -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
