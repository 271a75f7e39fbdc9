"""Builds libingvio_b200.so (sm_100a, -lineinfo) in-tree with nvcc. Used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "ingvio_b200", "csrc")
OUT = os.path.join(ROOT, "ingvio_b200", "lib")
LIB = os.path.join(OUT, "libingvio_b200.so")
SOURCES = ["igv_api.cu", "igv_frame.cu", "k_peak.cu", "k_tri.cu", "k_state.cu", "k_propagate.cu", "k_ekf.cu", "k_msckf.cu", "k_qr.cu", "k_gram.cu", "k_gnss.cu", "k_gnss_res.cu", "k_misc.cu", "k_tracks.cu", "k_landmark.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(verbose=False, ptxas_v=False):
    os.makedirs(OUT, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "ingvio_b200.h"))
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT, s.replace(".cu", ".o"))
        objs.append(obj)
        if _newer([src] + headers, obj):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0 or ptxas_v:
                sys.stderr.write(" ".join(cmd[-3:]) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, ptxas_v="-v" in sys.argv))
