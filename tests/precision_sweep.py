"""Reduced-precision sweep of the visual update at the ORACLE level (BASELINE.json configs[4]: "FP32 vs FP64 tol sweep").

TEST INFRASTRUCTURE / analysis: what tolerance on pose, trace(P) and the innovation an FP32 or TF32 variant of the device
path could promise, and how many chi^2-gate decisions it would flip, before any such kernel is written. The MSCKF update of
RemoveLostUpdate (RemoveLostUpdate.cpp:62-163: per-feature Jacobian, null space, gate, stacking, compression, EKF update) is
restated with an explicit arithmetic type per stage; everything else of the frame (propagation, augmentation,
marginalisation, GNSS update) stays the FP64 oracle, optionally with the state rounded to FP32 after every step.

Variants (stack = Jacobian + null space + gate + compression; ekf = S, K, P update, dx):
  fp64        stack f64 / ekf f64      must reproduce the oracle (tests/test_precision_sweep.py)
  stack32     stack f32 / ekf f64      FP32 (or tcgen05 kind::tf32-free FFMA) stack feeding the FP64 update
  stack_tf32  stack f32 with the stacked [H | r] rounded to TF32 (10-bit mantissa) before compression / ekf f64
  all32       everything f32, P and the mean stored in f32

    python tests/precision_sweep.py c2 12 c5 34      # workload, frames (pairs) -> profiles/r01_fp32_sweep.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ingvio_oracle as o  # noqa: E402
from ingvio_oracle import StateManager as SM  # noqa: E402
from ingvio_oracle.lie import skew  # noqa: E402

from helpers import filter_params, make_oracles  # noqa: E402
from ingvio_b200.synth import WORKLOADS, SyntheticStream  # noqa: E402


def tf32_round(x):
    """Round-to-nearest-even to a 10-bit mantissa (the operand format of kind::tf32 MMAs)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def tf32_split(x32):
    """x = hi + lo with both parts representable as TF32 operands (the "3xTF32" trick: hi*hi + hi*lo + lo*hi)."""
    hi = tf32_round(x32)
    lo = tf32_round((x32 - hi).astype(np.float32))
    return hi, lo


def gram_reduced(W, mode, chunk=128):
    """G = W^T W as a tcgen05 pipeline would form it: operands in the given format, FP32 accumulation inside a tile of
    `chunk` stack rows (the TMEM accumulator), tiles summed in FP64 (a fix-up pass).
      f32acc   FP32 operands, FP32 accumulation         (what a 3xTF32 split approximates)
      tf32     TF32 operands (kind::tf32), FP32 accumulation
      tf32x3   two-term TF32 split, three MMAs per tile (hi hi + hi lo + lo hi), FP32 accumulation"""
    W32 = W.astype(np.float32)
    n = W.shape[1]
    G = np.zeros((n, n), np.float64)
    for r0 in range(0, W.shape[0], chunk):
        T = W32[r0:r0 + chunk]
        if mode == "f32acc":
            acc = (T.T @ T).astype(np.float32)
        elif mode == "tf32":
            h = tf32_round(T)
            acc = (h.T @ h).astype(np.float32)
        else:
            h, l = tf32_split(T)
            acc = ((h.T @ h).astype(np.float32) + (h.T @ l).astype(np.float32) + (l.T @ h).astype(np.float32)).astype(np.float32)
        G += acc.astype(np.float64)
    return G


def compress_from_gram(G, n, tol=1e-13):
    """[Hc | rc] with Hc^T Hc = G[:n,:n], Hc^T rc = G[:n,n] (what the device's Gram path hands to the EKF update)."""
    lam, V = np.linalg.eigh(0.5 * (G[:n, :n] + G[:n, :n].T))
    keep = lam > tol * lam.max()
    Hc = np.sqrt(lam[keep])[:, None] * V[:, keep].T
    rc = (V[:, keep].T @ G[:n, n]) / np.sqrt(lam[keep])
    return Hc, rc


VARIANTS = {
    "fp64": dict(stack=np.float64, tf32=False, ekf=np.float64, store32=False),
    "stack32": dict(stack=np.float32, tf32=False, ekf=np.float64, store32=False),
    "stack_tf32": dict(stack=np.float32, tf32=True, ekf=np.float64, store32=False),
    "all32": dict(stack=np.float32, tf32=False, ekf=np.float32, store32=True),
    # the device's IGV_PREC_FP32_STACK mode (stack computed in f64, STORED as f32) with the Gram matrix of the stack formed
    # by a tensor-core pipeline in reduced precision (see gram_reduced)
    "store32_gram64": dict(stack=np.float64, tf32=False, ekf=np.float64, store32=False, gram="f64"),
    "gram_f32acc": dict(stack=np.float64, tf32=False, ekf=np.float64, store32=False, gram="f32acc"),
    "gram_tf32": dict(stack=np.float64, tf32=False, ekf=np.float64, store32=False, gram="tf32"),
    "gram_tf32x3": dict(stack=np.float64, tf32=False, ekf=np.float64, store32=False, gram="tf32x3"),
}


def _round_state(state, dt):
    state.cov = state.cov.astype(dt).astype(np.float64)
    e = state.extended_pose
    e.rot, e.vec1, e.vec2 = (a.astype(dt).astype(np.float64) for a in (e.rot, e.vec1, e.vec2))
    for c in state.sw_camleft_poses.values():
        c.rot, c.vec = c.rot.astype(dt).astype(np.float64), c.vec.astype(dt).astype(np.float64)


def visual_update(state, frame, noise, chi2_table, var, max_valid):
    """RemoveLostUpdate::updateStateMono after the track selection, mono, every stage in the arithmetic of `var`."""
    dt = var["stack"]
    times = state.sw_times()
    clones = [state.sw_camleft_poses[t] for t in times]
    ncl = len(clones)
    Rs = [c.rot.astype(dt) for c in clones]
    ps = [c.vec.astype(dt) for c in clones]
    idx_cols = np.concatenate([np.arange(c.idx(), c.idx() + 6) for c in clones])
    P_s = state.cov[np.ix_(idx_cols, idx_cols)].astype(dt)
    n = 6 * ncl
    sig2 = dt(noise) ** 2
    blocks, accepted, gammas = [], [], []
    F = frame.pf_w.shape[0]
    for f in range(F):
        slots = [s for s in range(ncl) if frame.obs_mask[f, s]]
        if len(slots) < 2:
            continue
        M = 2 * len(slots)
        pf = frame.pf_w[f].astype(dt)
        a = int(frame.anchor_slot[f])
        H = np.zeros((M, n), dt)
        Hf = np.zeros((M, 3), dt)
        r = np.zeros(M, dt)
        for k, s in enumerate(slots):
            R, p = Rs[s], ps[s]
            pc = R.T @ (pf - p)                                      # RemoveLostUpdate.cpp:209
            Hp = np.array([[1 / pc[2], 0, -pc[0] / pc[2] ** 2], [0, 1 / pc[2], -pc[1] / pc[2] ** 2]], dt)   # :211-215
            J = np.zeros((3, n), dt)
            if s != a:
                J[:, 6 * s:6 * s + 3] = R.T @ skew(pf).astype(dt)    # :231
                J[:, 6 * a:6 * a + 3] = -J[:, 6 * s:6 * s + 3]       # :232
            J[:, 6 * s + 3:6 * s + 6] = -R.T                         # :235
            H[2 * k:2 * k + 2] = Hp @ J
            Hf[2 * k:2 * k + 2] = Hp @ R.T                           # :237
            r[2 * k:2 * k + 2] = frame.obs[f, s, :2].astype(dt) - np.array([pc[0] / pc[2], pc[1] / pc[2]], dt)   # :253
        Q, _ = np.linalg.qr(Hf, mode="complete")                    # left null space (:268-272; any orthonormal basis)
        V = Q[:, 3:]
        Hn, rn = V.T @ H, V.T @ r
        S = Hn @ P_s @ Hn.T + sig2 * np.eye(M - 3, dtype=dt)         # Update.cpp:45-58
        gamma = float(rn @ np.linalg.solve(S, rn))
        dof = int(frame.obs_total[f]) - 1
        ok = gamma < chi2_table[dof - 1]
        gammas.append(gamma)
        if not ok:
            continue
        blocks.append(np.hstack([Hn, rn[:, None]]))
        accepted.append(f)
        if max_valid > 0 and len(accepted) >= max_valid:
            break
    if not blocks:
        return accepted, gammas, 0.0
    W = np.vstack(blocks)
    if var["tf32"]:
        W = tf32_round(W)
    de = var["ekf"]
    if var.get("gram"):
        W = W.astype(np.float32).astype(np.float64)                  # the stack is stored in single precision
        G = W.T @ W if var["gram"] == "f64" else gram_reduced(W, var["gram"])
        Hc, rc = compress_from_gram(G, n)
    else:
        if W.shape[0] > n:                                           # compression (:139-155), Q^T [H r]
            W = np.linalg.qr(W, mode="r")[:min(n + 1, W.shape[0])]
        Hc, rc = W[:, :n].astype(de), W[:, n].astype(de)
    P = state.cov.astype(de)
    PHt = P[:, idx_cols] @ Hc.T                                      # StateManager.cpp:381-397
    S = Hc @ PHt[idx_cols, :] + de(noise) ** 2 * np.eye(Hc.shape[0], dtype=de)
    K = np.linalg.solve(S.T, PHt.T).T                                # K = P H^T S^-1 (:405)
    Pn = P - K @ PHt.T
    state.cov = (0.5 * (Pn + Pn.T)).astype(np.float64)              # :407-411
    dx = (K @ rc).astype(np.float64)
    SM.box_plus(state, dx)
    return accepted, gammas, float(np.linalg.norm(rc))


class SweepFilter:
    def __init__(self, oracle_filter, variant, fp, max_valid):
        self.f, self.var, self.fp, self.max_valid = oracle_filter, VARIANTS[variant], fp, max_valid
        self.table = np.array([oracle_filter.remove_lost._table(d) for d in range(1, 200)])

    def step(self, fr):
        f, st = self.f, self.f.state
        f.propagate_augment(fr)
        if self.var["store32"]:
            _round_state(st, np.float32)
        acc, gam, rn = ([], [], 0.0)
        if fr.visual_mode is not None and fr.pf_w.shape[0] > 0:
            acc, gam, rn = visual_update(st, fr, self.fp.visual_noise, self.table, self.var, self.max_valid)
        f.marginalize(fr)
        f.gnss_update(fr)
        if self.var["store32"]:
            _round_state(st, np.float32)
        return set(acc), rn


def _rot_angle(Ra, Rb):
    """Angle of Ra^T Rb from its skew part (arccos of the trace loses half the digits near zero)."""
    M = Ra.T @ Rb
    w = 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def sweep(workload, n_frames, variants=("stack32", "stack_tf32", "all32"), seq0=0):
    wl = WORKLOADS[workload]
    fp = filter_params(wl)
    names = ("fp64",) + tuple(variants)
    runs = {}
    for v in names:
        st = SyntheticStream(wl, 1, seq0=seq0)
        runs[v] = (st, SweepFilter(make_oracles(wl, st, fp)[0], v, fp, wl.feats))
    out = {v: dict(pos=0.0, rot=0.0, trace_rel=0.0, P_rel=0.0, res_rel=0.0, flips=0, gated=0, diverged=False) for v in variants}
    for k in range(n_frames):
        res = {}
        for v in names:
            st, sf = runs[v]
            res[v] = sf.step(st.next_frame().seq(0))
        ref = runs["fp64"][1].f
        for v in variants:
            g = runs[v][1].f
            m = out[v]
            P0, P1 = ref.cov(), g.cov()
            if not np.all(np.isfinite(P1)) or P1.shape != P0.shape:
                m["diverged"] = True
                continue
            m["pos"] = max(m["pos"], float(np.linalg.norm(g.state.extended_pose.vec1 - ref.state.extended_pose.vec1)))
            m["rot"] = max(m["rot"], _rot_angle(g.state.extended_pose.rot, ref.state.extended_pose.rot))
            m["trace_rel"] = max(m["trace_rel"], abs(np.trace(P1) - np.trace(P0)) / np.trace(P0))
            m["P_rel"] = max(m["P_rel"], float(np.linalg.norm(P1 - P0) / max(1.0, np.linalg.norm(P0))))
            if res["fp64"][1] > 0:
                m["res_rel"] = max(m["res_rel"], abs(res[v][1] - res["fp64"][1]) / res["fp64"][1])
            m["flips"] += len(res[v][0] ^ res["fp64"][0])
            m["gated"] += len(res["fp64"][0])
    return dict(workload=workload, frames=n_frames, N=wl.dim, feats=wl.feats, results=out)


if __name__ == "__main__":
    args = sys.argv[1:] or ["c2", "12"]
    vs = tuple(os.environ["IGV_SWEEP_VARIANTS"].split(",")) if os.environ.get("IGV_SWEEP_VARIANTS") else ("stack32", "stack_tf32", "all32")
    table = [sweep(args[i], int(args[i + 1]), variants=vs) for i in range(0, len(args), 2)]
    path = os.path.join(ROOT, "profiles", os.environ.get("IGV_SWEEP_OUT", "r01_fp32_sweep.json"))
    with open(path, "w") as fh:
        json.dump(table, fh, indent=1)
    print(json.dumps(table, indent=1))
