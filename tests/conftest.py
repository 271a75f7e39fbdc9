import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


# IGV_TEST_LIB=emul: run against tests/emul/_build/libingvio_emul.so (the library's host code + the kernels the CPU execution
# model covers) instead of the CUDA library -- a development aid for machines without a GPU:
#   IGV_TEST_LIB=emul python -m pytest tests/test_gpu_tracks.py -m gpu
# The regular CPU suite uses the same library through tests/test_capi_on_cpu_model.py.
if os.environ.get("IGV_TEST_LIB") == "emul":
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_lib
    from ingvio_b200 import capi
    capi.LIB_PATH = build_lib.build()
