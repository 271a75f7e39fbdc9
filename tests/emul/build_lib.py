"""TEST INFRASTRUCTURE: builds tests/emul/_build/libingvio_emul.so -- the library's OWN host code (igv_api.cu and every
launcher) plus the kernels the CPU execution model can run (state / layout operations, GNSS rows, GNSS front end,
triangulation, track table), compiled by g++ from sources rewritten by preprocess.py (three syntactic rewrites), on a fake
CUDA runtime with host-memory semantics. The DMMA kernels (propagation, EKF update, per-track MSCKF kernel, compression)
are NOT modelled: C-ABI calls that need them fail with IGV_ERR_CUDA. Loaded by tests through ingvio_b200.capi with
LIB_PATH redirected; never part of the product."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ingvio_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
GEN = os.path.join(BUILD, "gen")
OUT = os.path.join(BUILD, "libingvio_emul.so")
MODELLED = ["igv_api.cu", "igv_frame.cu", "k_state.cu", "k_misc.cu", "k_gnss.cu", "k_gnss_res.cu", "k_tri.cu", "k_tracks.cu",
            "k_propagate.cu", "k_ekf.cu", "k_gram.cu", "k_qr.cu", "k_msckf.cu", "k_landmark.cu"]

sys.path.insert(0, HERE)
import preprocess  # noqa: E402


def build():
    os.makedirs(GEN, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in
            ("cuda_emul.h", "cuda_fake_rt.cpp", "lib_stubs.cpp", "preprocess.py", "build_lib.py")] + \
           [os.path.join(ROOT, "include", "ingvio_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    for f in os.listdir(CSRC):            # headers are included by relative name from the generated sources
        if f.endswith((".h", ".cuh")):
            preprocess.emit(os.path.join(CSRC, f), os.path.join(GEN, f))
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    common = ["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-w", "-I" + cuda_inc, "-I" + HERE, "-I" + GEN,
              "-DIGV_EMULATE=1", "-DIGV_EMULATE_LAUNCHERS=1", "-include", os.path.join(HERE, "cuda_emul.h")]
    jobs, objs = [], []
    for f in MODELLED:
        gen = preprocess.emit(os.path.join(CSRC, f), os.path.join(GEN, f.replace(".cu", ".cpp")))
        obj = os.path.join(BUILD, "emul_" + f.replace(".cu", ".o"))
        objs.append(obj)
        jobs.append(common + ["-c", gen, "-o", obj])
    for f in ("cuda_fake_rt.cpp", "lib_stubs.cpp"):
        obj = os.path.join(BUILD, "emul_" + f.replace(".cpp", ".o"))
        objs.append(obj)
        jobs.append(["g++", "-std=c++20", "-O1", "-fPIC", "-w", "-I" + cuda_inc, "-c", os.path.join(HERE, f), "-o", obj])

    def run(cmd):
        return cmd, subprocess.run(cmd, capture_output=True, text=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if r.returncode != 0:
                raise RuntimeError("emulation build failed: " + " ".join(cmd[-3:]) + "\n" + (r.stdout + r.stderr)[:4000])
    # -Bsymbolic: the model's own cuda* / igv_* definitions must win over a CUDA runtime another library (the real .so, torch)
    # may already have brought into the process
    r = subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", "-o", OUT] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation link failed:\n" + (r.stdout + r.stderr)[:4000])
    return OUT


if __name__ == "__main__":
    print(build())
