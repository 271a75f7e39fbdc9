"""The C-ABI's OWN host code (igv_api.cu: staging of host arguments, pointer modes, variable bookkeeping, window hooks of the
track table) together with the kernels the CPU execution model covers, built by tests/emul/build_lib.py into
libingvio_emul.so and driven through the regular binding (ingvio_b200.capi / BatchFilter) -- so the GPU tests of those
entry points also run, unchanged, in `pytest -m "not gpu"`. Calls that need a DMMA kernel must fail loudly, not fall back.

(IGV_TEST_LIB=emul python -m pytest tests/test_gpu_tracks.py -m gpu   runs whole GPU test files this way; slow.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emul"))

from ingvio_b200 import capi  # noqa: E402


@pytest.fixture(scope="module")
def cpu_model():
    import build_lib
    lib_path = build_lib.build()
    saved = (capi.LIB_PATH, capi._lib)
    capi.LIB_PATH, capi._lib = lib_path, None
    try:
        yield lib_path
    finally:
        capi.LIB_PATH, capi._lib = saved


def test_exports_every_header_symbol(cpu_model):
    lib = capi.load()
    for name in capi.header_symbols():
        assert hasattr(lib, name), name


def test_layout_and_init(cpu_model):
    import test_gpu_parity
    test_gpu_parity.test_layout_and_init()          # state init, add / marginalise GNSS variables, covariance read-back


def test_dmma_calls_fail_loudly(cpu_model):
    from ingvio_b200.filter import BatchFilter
    g = BatchFilter(1, 3, 4, 1)
    eye, z = np.eye(3).reshape(1, 9), np.zeros((1, 3))
    g.init_state_and_cov(eye, z, z, z, z, eye, z, np.full(21, 1e-2))
    with pytest.raises(capi.IgvError) as e:
        g.propagate_imu(np.zeros((1, 1, 3)), np.zeros((1, 1, 3)), np.full((1, 1), 0.01))
    assert e.value.status == capi.IGV_ERR_CUDA and "not available in the CPU model" in str(e.value)
    g.close()


@pytest.mark.parametrize("stereo", [False, True])
def test_track_table_reference_case(cpu_model, stereo):
    import test_gpu_tracks
    test_gpu_tracks.test_reference_map_server_test_on_device(stereo)


def test_track_table_depth_branches_and_errors(cpu_model):
    import test_gpu_tracks
    test_gpu_tracks.test_depth_branches_on_device()
    test_gpu_tracks.test_errors()


@pytest.mark.parametrize("mode,stereo,SW", [("sw_marg", False, 4), ("keyframe", True, 4)])
def test_track_table_scenario(cpu_model, mode, stereo, SW):
    import test_gpu_tracks
    from track_scenario import run_scenario
    B, F, T = 2, 16, 32
    cap = SW + 1 if mode == "sw_marg" else SW
    g = test_gpu_tracks._bare_filter(B, cap, F, T, stereo)
    augment, marg, clone_poses = test_gpu_tracks._window_ops(g)
    cov = run_scenario(g, augment, marg, clone_poses, mode, B, SW, stereo, frames=9, seed=31, F=F, meas_target=10, meas_stride=16)
    assert cov["lost"] > 0 and cov["seen"] > 0
    g.close()


def test_golden_replay(cpu_model):
    import test_golden_tracks
    test_golden_tracks.test_cuda_vs_golden("mono")


def test_cpp_map_server_mirror(cpu_model):
    """tests/cpp/test_map_server_mirror.cpp (TestMapServer.cpp:184-308 restated in C++) against the library's host code."""
    bdir = os.path.dirname(cpu_model)
    exe = os.path.join(bdir, "test_map_server_mirror_emul")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_map_server_mirror.cpp"), "-o", exe,
                        f"-L{bdir}", "-lingvio_emul", f"-Wl,-rpath,{bdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout, r.stdout + r.stderr
