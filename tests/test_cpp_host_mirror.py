"""Builds and runs the C++ host-mirror tests (tests/cpp/test_host_mirror.cpp): the reference's own gtest
cases restated against `ingvio::State` / `ingvio::StateManager` / `ingvio::UpdateBase` of
ingvio_b200/host/ingvio_host.hpp, which sit directly on the C-ABI -- the drop-in a C++ maintainer uses."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
LIBDIR = os.path.join(ROOT, "ingvio_b200", "lib")
LIBNAME = "ingvio_b200"
if os.environ.get("IGV_TEST_LIB") == "emul":      # development aid: the CPU model of the library (tests/conftest.py)
    LIBDIR, LIBNAME = os.path.join(ROOT, "tests", "emul", "_build"), "ingvio_emul"
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", SRC, "-o", EXE, f"-L{LIBDIR}", f"-l{LIBNAME}",
           f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libingvio_b200.so")), reason="library not built")
def test_host_mirror_compiles_and_links():
    """CPU: the header is plain C++17 and links against the C-ABI library only."""
    _build()


@pytest.mark.gpu
def test_host_mirror_reference_cases():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL TESTS PASSED" in r.stdout


# ---- map-server mirror (ingvio_b200/host/ingvio_map_server.hpp): TestMapServer.cpp:184-308 restated in C++ ----------
MS_SRC = os.path.join(ROOT, "tests", "cpp", "test_map_server_mirror.cpp")
EMUL = os.path.join(ROOT, "tests", "emul")


def _build_map_server_test(exe, libdir, libname):
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", MS_SRC, "-o", exe, f"-L{libdir}", f"-l{libname}", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_map_server_mirror_on_cpu_shim():
    """CPU: the mirror's own logic (views, slot <-> timestamp mapping, error paths) against tests/emul/igv_shim.cpp, i.e. the
    same C symbols backed by the track-table kernel source executed on the CPU."""
    bdir = os.path.join(EMUL, "_build")
    os.makedirs(bdir, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I" + cuda_inc, "-I" + EMUL,
                        os.path.join(EMUL, "igv_shim.cpp"), "-o", os.path.join(bdir, "libigv_shim.so")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = _build_map_server_test(os.path.join(bdir, "test_map_server_mirror_cpu"), bdir, "igv_shim")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libingvio_b200.so")), reason="library not built")
def test_map_server_mirror_links_against_the_library():
    _build_map_server_test(os.path.join(ROOT, "tests", "cpp", "_build", "test_map_server_mirror"), LIBDIR, LIBNAME)


@pytest.mark.gpu
def test_map_server_mirror_reference_cases():
    exe = _build_map_server_test(os.path.join(ROOT, "tests", "cpp", "_build", "test_map_server_mirror"), LIBDIR, LIBNAME)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout, r.stdout + r.stderr
