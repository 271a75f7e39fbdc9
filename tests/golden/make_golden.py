"""Generates tests/golden/*.npz: seeded frame packets of the `tiny` / `tiny_stereo` workloads and the state
the ORACLE reaches after each frame (covariance, packed mean, chi^2 statistics, accept counts).

The reference ships no golden vectors and cannot be built here, so these fixtures pin
the oracle itself against drift (tests/test_golden.py::test_oracle_reproduces_golden, CPU) and give the GPU
tests an oracle-free comparison target (::test_cuda_matches_golden). Regenerate with
    python tests/golden/make_golden.py
only when the oracle is deliberately changed (and say so in the commit).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from helpers import filter_params, make_oracles, oracle_packed_state  # noqa: E402
from ingvio_b200.synth import WORKLOADS, SyntheticStream  # noqa: E402

FRAME_KEYS = ("gyro", "accel", "dt", "pf_w", "anchor_slot", "obs", "obs_mask", "obs_total")
GNSS_KEYS = ("unit", "res_pos", "res_vel", "sys", "ura", "psr_std", "dopp_std_mps", "el", "R_enu2ecef")


def generate(wname, n_frames=9, batch=2):
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st = SyntheticStream(wl, batch)
    orc = make_oracles(wl, st, fp)
    out = {"n_frames": n_frames, "batch": batch}
    ini = st.initial_state()
    for k in ("R", "p", "v", "bg", "ba"):
        out["init_" + k] = ini[k]
    for i in range(n_frames):
        fr = st.next_frame()
        for k in FRAME_KEYS:
            out[f"f{i}_{k}"] = getattr(fr, k)
        out[f"f{i}_t"] = fr.t
        out[f"f{i}_visual"] = fr.visual_mode is not None
        out[f"f{i}_marg"] = np.array(fr.marg_slots, dtype=np.int32)
        out[f"f{i}_max_valid"] = fr.max_valid
        for k in GNSS_KEYS:
            out[f"f{i}_gnss_{k}"] = getattr(fr.gnss, k)
        P, X, G, A = [], [], [], []
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
            P.append(f.cov())
            X.append(oracle_packed_state(f, wl.sw))
            gl = f.last.get("gammas", []) if fr.visual_mode is not None else []
            gam = np.full(wl.feats, np.nan)
            for fid, g, dof, ok in gl:
                gam[fid] = g
            G.append(gam)
            A.append(sum(1 for x in gl if x[3]))
        out[f"f{i}_P"] = np.stack(P)
        out[f"f{i}_X"] = np.stack(X)
        out[f"f{i}_gamma"] = np.stack(G)
        out[f"f{i}_accepted"] = np.array(A, dtype=np.int32)
    return out


if __name__ == "__main__":
    for w in ("tiny", "tiny_stereo"):
        d = generate(w)
        path = os.path.join(HERE, f"{w}_frames.npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path) // 1024, "KiB")
