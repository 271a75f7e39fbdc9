"""Per-feature inverse-depth Levenberg-Marquardt triangulation with Huber weights.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/Triangulator.cpp:30-359 (+ the anchor-depth check of
MapServerManager.cpp:275-307). SURVEY.md §8f rank 1 ("next" row): the step right before the MSCKF
Jacobians. Pinned by the reference's own test (tests/test_oracle_triangulator.py restates
test/TestTriangulator.cpp:133-177).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class TriParams:
    """Triangulator.h:38-47 defaults (IngvioParams values: config/*/ingvio_*.yaml:35-45)."""
    trans_thres: float = 0.1
    huber_epsilon: float = 0.01
    conv_precision: float = 5e-7
    init_damping: float = 1e-3
    outer_loop_max_iter: int = 10
    inner_loop_max_iter: int = 10
    max_depth: float = 60.0
    min_depth: float = 0.2


class Triangulator:
    def __init__(self, prm: TriParams = None):
        self.p = prm or TriParams()

    # poses: list of (R c2w, p) in time order; obs: list of (2,) in the same order
    def find_longest_trans(self, poses, obs):
        """Triangulator.cpp:30-66. Returns (index of the view with the longest orthogonal baseline, length)."""
        R_last, p_last = poses[-1]
        u = np.array([obs[-1][0], obs[-1][1], 1.0])
        u /= np.linalg.norm(u)
        uw = R_last @ u
        proj = np.eye(3) - np.outer(uw, uw)
        max_len, max_i = -np.inf, len(poses) - 1
        for i, (R, p) in enumerate(poses[:-1]):
            t = abs(np.linalg.norm(proj @ (p - p_last)))
            if t > max_len:
                max_len, max_i = t, i
        return max_i, max_len

    @staticmethod
    def rel_poses(poses):
        """Triangulator.cpp:68-87: T_rel_i = T_i^-1 T_last (identity for the last)."""
        R_last, p_last = poses[-1]
        out = []
        for i, (R, p) in enumerate(poses):
            if i == len(poses) - 1:
                out.append((np.eye(3), np.zeros(3)))
            else:
                out.append((R.T @ R_last, R.T @ (p_last - p)))
        return out

    @staticmethod
    def init_depth(m1, m2, T12):
        """Triangulator.cpp:89-105."""
        R, t = T12
        tm1 = R @ np.array([m1[0], m1[1], 1.0])
        A = np.array([tm1[0] - m2[0] * tm1[2], tm1[1] - m2[1] * tm1[2]])
        b = np.array([m2[0] * t[2] - t[0], m2[1] * t[2] - t[1]])
        return float((A @ b) / (A @ A))

    @staticmethod
    def unit_cost(meas, rel, sol):
        """Triangulator.cpp:107-124."""
        with np.errstate(divide="ignore", invalid="ignore"):
            z = 1.0 / sol[2]
            pf0 = np.array([sol[0] * z, sol[1] * z, z])
            pf = rel[0] @ pf0 + rel[1]
            mh = pf[:2] / pf[2]
        d = np.asarray(meas) - mh
        return float(d @ d)

    def total_cost(self, obs, rels, sol):
        return sum(self.unit_cost(m, r, sol) for m, r in zip(obs, rels))

    def res_jacobian(self, meas, rel, sol):
        """Triangulator.cpp:138-171."""
        R, t = rel
        tp = R @ np.array([sol[0], sol[1], 1.0]) + t * sol[2]
        mh = tp[:2] / tp[2]
        res = mh - np.asarray(meas)
        W = np.array([[1.0 / tp[2], 0.0, -tp[0] / tp[2] ** 2], [0.0, 1.0 / tp[2], -tp[1] / tp[2] ** 2]])
        U = np.column_stack([R[:, 0], R[:, 1], t])
        J = W @ U
        e = np.linalg.norm(res)
        w = 1.0 if e <= self.p.huber_epsilon else np.sqrt(2.0 * self.p.huber_epsilon / e)
        return res, J, w

    def triangulate_mono(self, obs, poses):
        """Triangulator.cpp:173-318. obs/poses already filtered to common timestamps, time order.
        Returns (ok, pf_world)."""
        P = self.p
        if len(obs) <= 4:
            return False, np.zeros(3)
        max_i, max_trans = self.find_longest_trans(poses, obs)
        if max_trans < P.trans_thres:
            return False, np.zeros(3)
        rels = self.rel_poses(poses)
        with np.errstate(divide="ignore", invalid="ignore"):
            sol = np.array([obs[-1][0], obs[-1][1], 1.0 / self.init_depth(obs[-1], obs[max_i], rels[max_i])])
            total = self.total_cost(obs, rels, sol)
        lam = P.init_damping
        inner = outer = 0
        reduced = False
        delta_norm = np.inf
        while True:
            A = np.zeros((3, 3))
            b = np.zeros(3)
            for m, r in zip(obs, rels):
                res, J, w = self.res_jacobian(m, r, sol)
                if w == 1.0:
                    A += J.T @ J
                    b -= J.T @ res
                else:
                    A += w ** 2 * J.T @ J
                    b -= w ** 2 * J.T @ res
            while True:
                try:
                    delta = np.linalg.solve(A + lam * np.eye(3), b)   # Eigen ldlt().solve
                except np.linalg.LinAlgError:
                    delta = np.full(3, np.nan)
                new_sol = sol + delta
                delta_norm = float(np.linalg.norm(delta))
                with np.errstate(divide="ignore", invalid="ignore"):
                    new_total = self.total_cost(obs, rels, new_sol)
                if new_total < total:
                    total, sol, reduced = new_total, new_sol, True
                    lam = lam / 10.0 if lam / 10.0 > 1e-10 else 1e-10
                else:
                    reduced = False
                    lam = lam * 10 if lam * 10 < 1e12 else 1e12
                cont = inner < P.inner_loop_max_iter and not reduced
                inner += 1
                if not cont:
                    break
            inner = 0
            cont = outer < P.outer_loop_max_iter and delta_norm > P.conv_precision
            outer += 1
            if not cont:
                break
        with np.errstate(divide="ignore", invalid="ignore"):
            z = 1.0 / sol[2]
            pf_last = np.array([sol[0] * z, sol[1] * z, z])
        if (outer >= P.outer_loop_max_iter and inner >= P.inner_loop_max_iter) or delta_norm > P.conv_precision:
            return False, np.zeros(3)
        for R, t in rels:
            if (R @ pf_last + t)[2] <= P.min_depth:
                return False, np.zeros(3)
        if pf_last[2] < P.min_depth or pf_last[2] > P.max_depth:
            return False, np.zeros(3)
        pf = poses[-1][0] @ pf_last + poses[-1][1]
        if np.isnan(pf).any():
            return False, np.zeros(3)
        return True, pf

    def triangulate_stereo(self, sobs, poses, T_cl2cr):
        """Triangulator.cpp:320-359: each stereo view becomes two mono views (left, then right camera with
        pose T_left * T_cl2cr^-1), interleaved in time order."""
        Rc, pc = T_cl2cr
        obs, ps = [], []
        for z, (R, p) in zip(sobs, poses):
            obs.append(np.array(z[0:2]))
            ps.append((R, p))
            obs.append(np.array(z[2:4]))
            Rr = R @ Rc.T
            ps.append((Rr, p - Rr @ pc))
        return self.triangulate_mono(obs, ps)

    def triangulate_feature(self, obs_by_slot, mask, clone_poses, anchor_slot=None, stereo=False, T_cl2cr=None):
        """filterCommonTimestamp + triangulate + the anchor-depth check of
        FeatureInfoManager::triangulateFeatureInfo{Mono,Stereo} (MapServerManager.cpp:275-341)."""
        idx = [s for s in range(len(clone_poses)) if mask[s]]
        obs = [np.asarray(obs_by_slot[s], dtype=np.float64) for s in idx]
        poses = [clone_poses[s] for s in idx]
        ok, pf = (self.triangulate_stereo(obs, poses, T_cl2cr) if stereo else self.triangulate_mono(obs, poses))
        if ok and anchor_slot is not None:
            Ra, pa = clone_poses[anchor_slot]
            if (Ra.T @ (pf - pa))[2] <= 0:
                return False, pf
        return ok, pf
