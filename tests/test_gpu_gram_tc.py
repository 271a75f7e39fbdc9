"""The tcgen05 Gram kernel (ingvio_b200/csrc/k_gram_tc.cuh, IGV_PREC_TF32_GRAM) against a double-precision Gram matrix of the same
float stack: tests/cuda/gram_tc_harness.cu is compiled here with nvcc and run -- ragged and rejected tracks (their rows hold NaN),
the max_valid cap, 1-4 partial matrices, 31 / 67 / 128 / 150 / 181 columns (one to six operand atoms, with and without the second
output tile), then two timed cases. Bar: |dG| <= 1e-5 sqrt(G_rr G_cc) per entry (measured 1.5e-6 ... 3e-6: the unit's FP32
accumulation over 128 rows). The library-level tolerance of the mode is tests/test_gpu_precision.py."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gram_tc_harness(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "gram_tc_harness")
    src = os.path.join(ROOT, "tests", "cuda", "gram_tc_harness.cu")
    b = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe, src],
                       capture_output=True, text=True)
    assert b.returncode == 0, b.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
