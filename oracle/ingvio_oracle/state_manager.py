"""Covariance algebra of the filter.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/StateManager.cpp:27-694, function by function,
same block order and same arithmetic order where it matters for rounding.
"""
import numpy as np
from scipy.stats import chi2 as _chi2

from .state import BDS, FS, GPS, State
from .types import SE3, Scalar


class StateManager:
    @staticmethod
    def check_state_continuity(state: State):
        """StateManager.cpp:27-40."""
        idx = 0
        for v in state.err_variables:
            if v.idx() != idx:
                return False
            idx += v.size()
        return state.cov.shape == (idx, idx)

    @staticmethod
    def propagate_state_cov(state: State, Phi_imu, G_imu, dt):
        """StateManager.cpp:42-119."""
        sp = state.state_params
        P = state.cov
        n = P.shape[0]
        cov_tmp = np.empty((n, n))
        cov_tmp[:15, :15] = Phi_imu @ P[:15, :15] @ Phi_imu.T
        cov21 = P[15:, :15] @ Phi_imu.T
        cov22 = P[15:, 15:].copy()
        if sp.enable_gnss and FS in state.gnss:
            cov21_tmp = cov21.copy()
            cov22_tmp = cov22.copy()
            lc = state.gnss[FS].idx() - 15
            for i in range(GPS, BDS + 1):
                if i in state.gnss:
                    lr = state.gnss[i].idx() - 15
                    cov21_tmp[lr, :] += dt * cov21[lc, :]
                    cov22_tmp[lr, :] += dt * cov22[lc, :]
            cov21 = cov21_tmp
            cov22 = cov22_tmp.copy()
            for i in range(GPS, BDS + 1):
                if i in state.gnss:
                    lr = state.gnss[i].idx() - 15
                    cov22_tmp[:, lr] += dt * cov22[:, lc]
            cov22 = cov22_tmp
        cov_tmp[15:, :15] = cov21
        cov_tmp[:15, 15:] = cov21.T
        cov_tmp[15:, 15:] = cov22
        G_tmp = np.array(G_imu, dtype=np.float64).copy()
        G_tmp[:, 0:3] *= sp.noise_g
        G_tmp[:, 3:6] *= sp.noise_a
        G_tmp[:, 6:9] *= sp.noise_bg
        G_tmp[:, 9:12] *= sp.noise_ba
        cov_tmp[:15, :15] += dt * Phi_imu @ G_tmp @ G_tmp.T @ Phi_imu.T
        if sp.enable_gnss:
            for i in range(5):
                if i not in state.gnss:
                    continue
                for j in range(5):
                    if j not in state.gnss:
                        continue
                    a, b = state.gnss[i].idx(), state.gnss[j].idx()
                    if i != FS and j != FS:
                        cov_tmp[a, b] += dt * sp.noise_clockbias ** 2 + dt ** 3 * sp.noise_cb_rw ** 2
                    elif i == FS and j == FS:
                        cov_tmp[a, b] += dt * sp.noise_cb_rw ** 2
                    else:
                        cov_tmp[a, b] += dt ** 2 * sp.noise_cb_rw ** 2
        state.cov = 0.5 * (cov_tmp + cov_tmp.T)

    @staticmethod
    def get_full_cov(state: State):
        return state.cov.copy()

    @staticmethod
    def get_marginal_cov(state: State, small_variables):
        """StateManager.cpp:128-153."""
        sel = np.concatenate([np.arange(v.idx(), v.idx() + v.size()) for v in small_variables]) \
            if small_variables else np.zeros(0, dtype=int)
        return state.cov[np.ix_(sel, sel)].copy()

    @staticmethod
    def marginalize(state: State, marg):
        """StateManager.cpp:155-192."""
        if not any(v is marg for v in state.err_variables):
            raise RuntimeError("[StateManager]: Marg is not in the current state!")
        s, sz = marg.idx(), marg.size()
        keep = np.r_[0:s, s + sz:state.cov.shape[0]]
        state.cov = state.cov[np.ix_(keep, keep)].copy()
        remaining = []
        for v in state.err_variables:
            if v is not marg:
                if v.idx() > s:
                    v.set_cov_idx(v.idx() - sz)
                remaining.append(v)
        marg.set_cov_idx(-1)
        state.err_variables = remaining

    @staticmethod
    def add_variable_independent(state: State, new_state, cov_block):
        """StateManager.cpp:194-214."""
        cov_block = np.atleast_2d(np.asarray(cov_block, dtype=np.float64))
        assert new_state.size() == cov_block.shape[0]
        old = state.curr_cov_size()
        new_cov = np.zeros((old + new_state.size(),) * 2)
        new_cov[:old, :old] = state.cov
        new_cov[old:, old:] = cov_block
        new_state.set_cov_idx(old)
        state.cov = new_cov
        state.err_variables.append(new_state)

    @staticmethod
    def add_gnss_variable(state: State, gtype, value, cov):
        """StateManager.cpp:216-231."""
        s = Scalar()
        s.set_value(value)
        state.gnss[gtype] = s
        StateManager.add_variable_independent(state, s, [[cov]])

    @staticmethod
    def marg_gnss_variable(state: State, gtype):
        """StateManager.cpp:233-242."""
        StateManager.marginalize(state, state.gnss[gtype])
        del state.gnss[gtype]

    @staticmethod
    def box_plus(state: State, dx):
        """StateManager.cpp:244-251."""
        assert dx.shape[0] == state.curr_cov_size()
        for v in state.err_variables:
            v.update(dx)

    @staticmethod
    def augment_sliding_window_pose(state: State):
        """StateManager.cpp:253-296."""
        if state.timestamp in state.sw_camleft_poses:
            return
        clone = SE3()
        R_i2w = state.extended_pose.value_linear()
        p_i2w = state.extended_pose.value_trans1()
        R_c2i = state.camleft_imu_extrinsics.value_linear()
        p_c2i = state.camleft_imu_extrinsics.value_trans()
        clone.set_value(R_i2w @ R_c2i, R_i2w @ p_c2i + p_i2w)
        n = state.curr_cov_size()
        clone.set_cov_idx(n)
        state.sw_camleft_poses[state.timestamp] = clone
        state.err_variables.append(clone)
        J = np.zeros((6, 21))
        J[:6, :6] = np.eye(6)
        J[0:3, 15:18] = R_i2w
        J[3:6, 18:21] = R_i2w
        cov_new = np.zeros((n + 6, n + 6))
        cov_new[:n, :n] = state.cov
        cov_new[n:, n:] = J @ state.cov[:21, :21] @ J.T
        cov_new[n:, :n] = J @ state.cov[:21, :n]
        cov_new[:n, n:] = cov_new[n:, :n].T
        state.cov = 0.5 * (cov_new + cov_new.T)

    @staticmethod
    def marg_sliding_window_pose(state: State, marg_time=None):
        """StateManager.cpp:316-338."""
        if marg_time is None:
            marg_time = state.next_marg_time()
            if marg_time == float("inf"):
                return
        StateManager.marginalize(state, state.sw_camleft_poses[marg_time])
        del state.sw_camleft_poses[marg_time]

    @staticmethod
    def add_anchored_landmark_in_state(state: State, lm, lm_id, cov):
        """StateManager.cpp:298-314."""
        if lm_id in state.anchored_landmarks:
            return
        state.anchored_landmarks[lm_id] = lm
        StateManager.add_variable_independent(state, lm, cov)

    @staticmethod
    def _ph_t(state: State, var_order, H):
        """The type-indexed block loop of StateManager.cpp:381-397 / :497-511 / :666-682."""
        H = np.atleast_2d(H)
        PH_T = np.zeros((state.cov.shape[0], H.shape[0]))
        off = 0
        h_idx = []
        for v in var_order:
            h_idx.append(off)
            off += v.size()
        for tv in state.err_variables:
            acc = np.zeros((tv.size(), H.shape[0]))
            for v, h0 in zip(var_order, h_idx):
                acc += state.cov[tv.idx():tv.idx() + tv.size(), v.idx():v.idx() + v.size()] @ \
                    H[:, h0:h0 + v.size()].T
            PH_T[tv.idx():tv.idx() + tv.size(), :] = acc
        return PH_T

    @staticmethod
    def ekf_update(state: State, var_order, H, res, R, return_dx=False):
        """StateManager.cpp:359-426."""
        H = np.atleast_2d(np.asarray(H, dtype=np.float64))
        res = np.asarray(res, dtype=np.float64).reshape(-1)
        R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        assert res.shape[0] == R.shape[0] == H.shape[0] and R.shape[0] == R.shape[1]
        assert StateManager.check_sub_order(state, var_order)
        PH_T = StateManager._ph_t(state, var_order, H)
        small_cov = StateManager.get_marginal_cov(state, var_order)
        S = H @ small_cov @ H.T + R
        K = PH_T @ np.linalg.inv(S)
        cov_tmp = state.cov - K @ PH_T.T
        state.cov = 0.5 * (cov_tmp + cov_tmp.T)
        neg = bool(np.any(np.diag(state.cov) < 0.0))
        dx = K @ res
        StateManager.box_plus(state, dx)
        if return_dx:
            return dx, neg
        return None

    @staticmethod
    def check_sub_order(state: State, sub_order):
        return all(any(v is w for w in state.err_variables) for v in sub_order)

    @staticmethod
    def calc_sub_var_size(sub_var):
        return sum(v.size() for v in sub_var if v is not None)

    @staticmethod
    def add_variable_delayed_invertible(state: State, var_new, var_old_order, H_old, H_new, res, noise_iso_meas):
        """StateManager.cpp:462-541."""
        if any(v is var_new for v in state.err_variables):
            return
        H_old = np.atleast_2d(H_old)
        H_new = np.atleast_2d(H_new)
        PH_T = StateManager._ph_t(state, var_old_order, H_old)
        small_cov = StateManager.get_marginal_cov(state, var_old_order)
        S = H_old @ small_cov @ H_old.T
        S[np.diag_indices_from(S)] += noise_iso_meas ** 2.0
        H_new_inv = np.linalg.inv(H_new)
        cov_newnew = H_new_inv @ S @ H_new_inv.T
        n, k = state.cov.shape[0], var_new.size()
        cov_tmp = np.zeros((n + k, n + k))
        cov_tmp[:n, :n] = state.cov
        cov_tmp[n:, n:] = cov_newnew
        cov_tmp[:n, n:] = -PH_T @ H_new_inv.T
        cov_tmp[n:, :n] = cov_tmp[:n, n:].T
        var_new.set_cov_idx(n)
        state.err_variables.append(var_new)
        state.cov = 0.5 * (cov_tmp + cov_tmp.T)

    @staticmethod
    def add_variable_delayed(state: State, var_new, var_old_order, H_old, H_new, res,
                             noise_iso_meas, chi2_mult_factor, do_chi2=True):
        """StateManager.cpp:547-630. H_old, H_new, res are modified in place like the reference."""
        if any(v is var_new for v in state.err_variables):
            return False
        new_sz = var_new.size()
        if H_new.shape[0] <= H_new.shape[1]:
            return False
        # Givens sweep, StateManager.cpp:580-592 (Eigen makeGivens + applyOnTheLeft(adjoint))
        for n in range(H_new.shape[1]):
            for m in range(H_new.shape[0] - 1, n, -1):
                p, q = H_new[m - 1, n], H_new[m, n]
                c, s = _make_givens(p, q)
                # adjoint of J=[c s; -s c] applied on the left of rows (m-1, m): [c -s; s c]
                for M in (H_new[:, n:], None, H_old):
                    if M is None:
                        a, b = res[m - 1], res[m]
                        res[m - 1], res[m] = c * a - s * b, s * a + c * b
                        continue
                    a = M[m - 1, :].copy()
                    b = M[m, :].copy()
                    M[m - 1, :] = c * a - s * b
                    M[m, :] = s * a + c * b
        Hxinit = H_old[:new_sz, :].copy()
        Hfinit = H_new[:new_sz, :new_sz].copy()
        resinit = res[:new_sz].copy()
        Hup = H_old[new_sz:, :].copy()
        resup = res[new_sz:].copy()
        small_cov = StateManager.get_marginal_cov(state, var_old_order)
        S = Hup @ small_cov @ Hup.T
        S[np.diag_indices_from(S)] += noise_iso_meas ** 2.0
        chi2 = float(resup @ np.linalg.solve(S, resup)) if resup.size else 0.0
        chi2_check = float(_chi2.ppf(0.95, res.shape[0]))
        if chi2 > chi2_mult_factor * chi2_check and do_chi2:
            return False
        StateManager.add_variable_delayed_invertible(state, var_new, var_old_order, Hxinit, Hfinit,
                                                     resinit, noise_iso_meas)
        if Hup.shape[0] > 0:
            StateManager.ekf_update(state, var_old_order, Hup, resup,
                                    noise_iso_meas ** 2.0 * np.eye(resup.shape[0]))
        return True

    @staticmethod
    def replace_var_linear(state: State, target_var, dependence_order, H):
        """StateManager.cpp:632-693."""
        if not any(v is target_var for v in state.err_variables):
            return
        H = np.atleast_2d(H)
        assert target_var.size() == H.shape[0]
        PH_T = StateManager._ph_t(state, dependence_order, H)
        small_cov = StateManager.get_marginal_cov(state, dependence_order)
        HPH_T = H @ small_cov @ H.T
        t0, ts = target_var.idx(), target_var.size()
        state.cov[:, t0:t0 + ts] = PH_T
        state.cov[t0:t0 + ts, :] = PH_T.T
        state.cov[t0:t0 + ts, t0:t0 + ts] = HPH_T


def _make_givens(p, q):
    """Eigen::JacobiRotation::makeGivens (real case, Eigen/src/Jacobi/Jacobi.h, 3.3.x):
    returns (c, s) such that  [c s; -s c]^T [p; q] = [r; 0]. Third-party (Eigen, unpinned
    system version, README.md:34-36); published algorithm restated."""
    if q == 0.0:
        return (1.0 if p >= 0 else -1.0), 0.0
    if p == 0.0:
        return 0.0, (1.0 if q < 0 else -1.0)
    if abs(p) > abs(q):
        t = q / p
        u = np.sqrt(1.0 + t * t)
        if p < 0:
            u = -u
        c = 1.0 / u
        s = -t * c
        return c, s
    t = p / q
    u = np.sqrt(1.0 + t * t)
    if q < 0:
        u = -u
    s = -1.0 / u
    c = -t * s
    return c, s
