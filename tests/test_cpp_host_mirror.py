"""Builds and runs the C++ host-mirror tests (tests/cpp/test_host_mirror.cpp): the reference's own gtest
cases restated against `ingvio::State` / `ingvio::StateManager` / `ingvio::UpdateBase` of
ingvio_b200/host/ingvio_host.hpp, which sit directly on the C-ABI -- the drop-in a C++ maintainer uses."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
LIBDIR = os.path.join(ROOT, "ingvio_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", SRC, "-o", EXE, f"-L{LIBDIR}", "-lingvio_b200",
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
