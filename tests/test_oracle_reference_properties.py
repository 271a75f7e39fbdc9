"""Pins the oracle against the reference's OWN property tests (no golden vectors exist upstream).

Each test restates one gtest of /root/reference/ingvio_estimator/test/ 1:1 (file:line in the
docstring), with the reference's tolerance. Random inputs are seeded here (upstream uses unseeded
Eigen::Random()).
"""
import math

import numpy as np
import pytest

import ingvio_oracle as o
from ingvio_oracle import BDS, FS, GLO, GPS, YOF, StateManager as SM


def rand_rot(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def expm_so3(v):
    th = np.linalg.norm(v)
    K = o.skew(v)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * K @ K


def test_basic_funcs():
    """TestStateManager.cpp:31-51 (1e-8)."""
    rng = np.random.default_rng(1)
    v = rng.uniform(-1, 1, 3)
    assert np.linalg.norm(o.vee(o.skew(v)) - v) < 1e-8
    assert np.linalg.norm(o.gamma_func(v, 0) - expm_so3(v)) < 1e-8
    z = np.zeros(3)
    assert np.linalg.norm(o.gamma_func(z, 1) - np.eye(3)) < 1e-8
    assert np.linalg.norm(o.gamma_func(z, 2) - 0.5 * np.eye(3)) < 1e-8
    assert np.linalg.norm(o.gamma_func(z, 3) - np.eye(3) / 6) < 1e-8


def test_gamma_series():
    """Gamma_m(phi) = sum_k phi^k/(k+m)!  (definition, AuxGammaFunc.h) for m=0..3."""
    rng = np.random.default_rng(2)
    for _ in range(5):
        v = rng.uniform(-1.5, 1.5, 3)
        K = o.skew(v)
        for m in range(4):
            acc = np.zeros((3, 3))
            term = np.eye(3)
            for k in range(40):
                acc += term / math.factorial(k + m)
                term = term @ K
            assert np.linalg.norm(o.gamma_func(v, m) - acc) < 1e-10


def _state_with_gnss(fp=None):
    fp = fp or o.FilterParams(enable_gnss=1)
    st = o.State(fp)
    return fp, st


def test_state_add_marg_prop():
    """TestStateManager.cpp:53-156: dims, add/marg of GNSS scalars, propagateStateCov closed form (1e-10)."""
    rng = np.random.default_rng(3)
    fp, st = _state_with_gnss()
    assert SM.check_state_continuity(st) and st.curr_cov_size() == 21
    SM.add_gnss_variable(st, GPS, 20.0, 4.0)
    assert st.curr_cov_size() == 22 and st.curr_err_variable_size() == 5
    SM.add_gnss_variable(st, BDS, 16.0, 4.0)
    SM.add_gnss_variable(st, YOF, 123.0, 1.0)
    SM.add_gnss_variable(st, FS, 2.0, 1.0)
    assert st.curr_cov_size() == 25 and st.curr_err_variable_size() == 8
    SM.marg_gnss_variable(st, GPS)
    assert st.curr_cov_size() == 24 and st.curr_err_variable_size() == 7
    SM.add_gnss_variable(st, GLO, 3.0, 6.0)
    assert st.curr_cov_size() == 25 and st.curr_err_variable_size() == 8
    # make P a generic SPD matrix so the closed form is exercised on dense cross terms too
    A = rng.standard_normal((25, 25))
    st.cov = A @ A.T / 25 + np.eye(25) * 0.1
    cov = SM.get_full_cov(st)
    Phi_imu = rng.uniform(-1, 1, (15, 15))
    G_imu = rng.uniform(-1, 1, (15, 12))
    dt = 1.5
    sp = st.state_params
    Q = np.zeros((14, 14))
    Q[0:3, 0:3] = np.eye(3) * sp.noise_g ** 2
    Q[3:6, 3:6] = np.eye(3) * sp.noise_a ** 2
    Q[6:9, 6:9] = np.eye(3) * sp.noise_bg ** 2
    Q[9:12, 9:12] = np.eye(3) * sp.noise_ba ** 2
    Q[12, 12] = sp.noise_clockbias ** 2
    Q[13, 13] = sp.noise_cb_rw ** 2
    # order after the ops above: BDS(21) YOF(22) FS(23) GLO(24)
    assert [st.gnss[k].idx() for k in (BDS, YOF, FS, GLO)] == [21, 22, 23, 24]
    Phi_g = np.eye(4)
    Phi_g[0, 2] = dt
    Phi_g[3, 2] = dt
    Phi = np.eye(25)
    Phi[:15, :15] = Phi_imu
    Phi[21:, 21:] = Phi_g
    G = np.zeros((25, 14))
    G[:15, :12] = G_imu
    Gg = np.zeros((4, 2))
    Gg[0, 0] = 1
    Gg[2, 1] = 1
    Gg[3, 0] = 1
    G[21:, 12:] = Gg
    ref = Phi @ cov @ Phi.T + dt * Phi @ G @ Q @ G.T @ Phi.T
    SM.propagate_state_cov(st, Phi_imu, G_imu, dt)
    assert np.linalg.norm(SM.get_full_cov(st) - ref) < 1e-10 * max(1.0, np.linalg.norm(ref))
    small = SM.get_marginal_cov(st, [st.extended_pose, st.ba, st.gnss[YOF], st.gnss[GLO]])
    full = SM.get_full_cov(st)
    assert np.linalg.norm(small[:9, :9] - full[:9, :9]) < 1e-6
    assert np.linalg.norm(small[9:12, 9:12] - full[12:15, 12:15]) < 1e-6
    assert abs(small[13, 12] - full[24, 22]) < 1e-12


def _fixture_state(rng):
    """TestStateManager.cpp:160-187 fixture."""
    fp = o.FilterParams(enable_gnss=1)
    st = o.State(fp)
    st.extended_pose.rot = rand_rot(rng)
    st.extended_pose.vec1 = rng.uniform(-1, 1, 3)
    st.extended_pose.vec2 = rng.uniform(-1, 1, 3)
    st.bg.set_value(rng.uniform(-1, 1, 3))
    st.ba.set_value(rng.uniform(-1, 1, 3))
    st.camleft_imu_extrinsics.set_value(rand_rot(rng), rng.uniform(-1, 1, 3))
    SM.add_gnss_variable(st, GPS, 20.0, 4.0)
    SM.add_gnss_variable(st, YOF, 123.0, 1.0)
    SM.add_gnss_variable(st, FS, 2.0, 1.0)
    SM.add_gnss_variable(st, BDS, 16.0, 4.0)
    return fp, st, rng.uniform(-1, 1, (15, 15)), rng.uniform(-1, 1, (15, 12))


def test_augment_pose():
    """TestStateManager.cpp:195-255 (1e-8)."""
    rng = np.random.default_rng(4)
    fp, st, Phi, G = _fixture_state(rng)

    def largeJ(n, C):
        J = np.zeros((n + 6, n))
        J[:n, :n] = np.eye(n)
        J[n:n + 6, :6] = np.eye(6)
        J[n:n + 3, 15:18] = C
        J[n + 3:n + 6, 18:21] = C
        return J

    st.timestamp = 1.0
    for t_aug, dt in ((2.5, 1.5), (3.0, 0.5)):
        SM.propagate_state_cov(st, Phi, G, dt)
        st.timestamp = t_aug
        cov1 = SM.get_full_cov(st)
        J1 = largeJ(cov1.shape[0], st.extended_pose.rot)
        SM.augment_sliding_window_pose(st)
        ref = J1 @ cov1 @ J1.T
        assert np.linalg.norm(SM.get_full_cov(st) - ref) < 1e-8 * max(1, np.linalg.norm(ref))
        assert st.curr_cov_size() == cov1.shape[0] + 6
        cl = st.sw_camleft_poses[t_aug]
        assert np.linalg.norm(st.extended_pose.rot @ st.camleft_imu_extrinsics.rot - cl.rot) < 1e-8
        assert np.linalg.norm(st.extended_pose.vec1 + st.extended_pose.rot @ st.camleft_imu_extrinsics.vec
                              - cl.vec) < 1e-8


def test_state_box_plus():
    """TestStateManager.cpp:257-476, box-plus part (:396-455) (1e-8)."""
    rng = np.random.default_rng(5)
    fp, st, Phi, G = _fixture_state(rng)
    st.timestamp = 2.5
    SM.augment_sliding_window_pose(st)
    st.timestamp = 3.0
    SM.augment_sliding_window_pose(st)
    lm = o.AnchoredLandmark()
    lm.reset_anchored_pose(st.sw_camleft_poses[2.5])
    lm.set_value_pos_xyz(rng.uniform(-1, 1, 3))
    SM.add_anchored_landmark_in_state(st, lm, 5, 10.0 * np.eye(3))
    n = st.curr_cov_size()
    assert n == 21 + 4 + 12 + 3
    dx = rng.uniform(-1, 1, n)
    R0, p0, v0 = st.extended_pose.rot.copy(), st.extended_pose.vec1.copy(), st.extended_pose.vec2.copy()
    bg0, ba0 = st.bg.value().copy(), st.ba.value().copy()
    Re0, pe0 = st.camleft_imu_extrinsics.rot.copy(), st.camleft_imu_extrinsics.vec.copy()
    c0 = st.sw_camleft_poses[2.5]
    Rc0, pc0 = c0.rot.copy(), c0.vec.copy()
    gps0 = st.gnss[GPS].value()
    pf0 = lm.value_pos_xyz().copy()
    SM.box_plus(st, dx)
    G0, G1 = o.gamma_func(dx[0:3], 0), o.gamma_func(dx[0:3], 1)
    assert np.linalg.norm(st.extended_pose.rot - G0 @ R0) < 1e-8
    assert np.linalg.norm(st.extended_pose.vec1 - (G0 @ p0 + G1 @ dx[3:6])) < 1e-8
    assert np.linalg.norm(st.extended_pose.vec2 - (G0 @ v0 + G1 @ dx[6:9])) < 1e-8
    assert np.linalg.norm(st.bg.value() - (bg0 + dx[9:12])) < 1e-8
    assert np.linalg.norm(st.ba.value() - (ba0 + dx[12:15])) < 1e-8
    Ge0, Ge1 = o.gamma_func(dx[15:18], 0), o.gamma_func(dx[15:18], 1)
    assert np.linalg.norm(st.camleft_imu_extrinsics.rot - Ge0 @ Re0) < 1e-8
    assert np.linalg.norm(st.camleft_imu_extrinsics.vec - (Ge0 @ pe0 + Ge1 @ dx[18:21])) < 1e-8
    assert abs(st.gnss[GPS].value() - (gps0 + dx[st.gnss[GPS].idx()])) < 1e-8
    i = c0.idx()
    Gc0, Gc1 = o.gamma_func(dx[i:i + 3], 0), o.gamma_func(dx[i:i + 3], 1)
    assert np.linalg.norm(c0.rot - Gc0 @ Rc0) < 1e-8
    assert np.linalg.norm(c0.vec - (Gc0 @ pc0 + Gc1 @ dx[i + 3:i + 6])) < 1e-8
    j = lm.idx()
    assert np.linalg.norm(lm.value_pos_xyz() - (Gc0 @ pf0 + Gc1 @ dx[j:j + 3])) < 1e-8
    # marg order restores dim 25 (TestStateManager.cpp:457-475)
    SM.marginalize(st, lm)
    SM.marg_sliding_window_pose(st, 2.5)
    SM.marg_sliding_window_pose(st, 3.0)
    assert st.curr_cov_size() == 25 and SM.check_state_continuity(st)


def test_state_cov_update():
    """TestStateManager.cpp:478-557: ekfUpdate covariance == (I-KH)P for a sparse var_order (1e-8)."""
    rng = np.random.default_rng(6)
    fp, st, Phi, G = _fixture_state(rng)
    st.timestamp = 1.0
    SM.propagate_state_cov(st, Phi, G, 1.5)
    st.timestamp = 2.5
    SM.augment_sliding_window_pose(st)
    lm = o.AnchoredLandmark()
    lm.reset_anchored_pose(st.sw_camleft_poses[2.5])
    lm.set_value_pos_xyz(rng.uniform(-1, 1, 3))
    SM.add_anchored_landmark_in_state(st, lm, 5, 10.0 * np.eye(3))
    var_order = [st.extended_pose, st.gnss[GPS], st.gnss[BDS], st.gnss[FS]]
    assert SM.calc_sub_var_size(var_order) == 12
    res = rng.uniform(-1, 1, 6)
    H = np.zeros((6, 12))
    H[:, 0:3] = rng.uniform(-1, 1, (6, 3))
    H[0:3, 3:6] = rng.uniform(-1, 1, (3, 3))
    H[3:6, 6:9] = rng.uniform(-1, 1, (3, 3))
    H[0, 9] = H[1, 9] = H[2, 10] = 1.0
    H[3:6, 11] = 1.0
    n = st.curr_cov_size()
    HL = np.zeros((6, n))
    HL[:, :9] = H[:, :9]
    HL[0, st.gnss[GPS].idx()] = HL[1, st.gnss[GPS].idx()] = 1.0
    HL[2, st.gnss[BDS].idx()] = 1.0
    HL[3:6, st.gnss[FS].idx()] = 1.0
    P0 = SM.get_full_cov(st)
    R = 0.5 * np.eye(6)
    SM.ekf_update(st, var_order, H, res, R)
    K = P0 @ HL.T @ np.linalg.inv(HL @ P0 @ HL.T + R)
    ref = (np.eye(n) - K @ HL) @ P0
    assert np.linalg.norm(ref - SM.get_full_cov(st)) < 1e-8 * max(1, np.linalg.norm(ref))


def test_add_var_delayed():
    """TestStateManager.cpp:594-725: addVariableDelayedInvertible formulas and addVariableDelayed =
    Givens split + invertible init + residual EKF (1e-8)."""
    rng = np.random.default_rng(7)
    fp, st, Phi, G = _fixture_state(rng)
    st.timestamp = 1.0
    SM.propagate_state_cov(st, Phi, G, 1.5)
    st.timestamp = 2.5
    SM.augment_sliding_window_pose(st)
    order = [st.extended_pose, st.gnss[YOF]]
    # --- invertible
    P0 = SM.get_full_cov(st)
    n = P0.shape[0]
    Hx = rng.uniform(-1, 1, (1, 10))
    Hf = np.array([[1.7]])
    res = rng.uniform(-1, 1, 1)
    noise = 0.3
    var = o.Scalar()
    st2 = st
    SM.add_variable_delayed_invertible(st2, var, order, Hx, Hf, res, noise)
    HL = np.zeros((1, n))
    HL[:, :9] = Hx[:, :9]
    HL[0, st.gnss[YOF].idx()] = Hx[0, 9]
    P1 = SM.get_full_cov(st2)
    Hfi = np.linalg.inv(Hf)
    assert np.linalg.norm(P1[:n, :n] - P0) < 1e-8
    assert np.linalg.norm(P1[:n, n:] + P0 @ HL.T @ Hfi.T) < 1e-8
    assert np.linalg.norm(P1[n:, n:] - Hfi @ (HL @ P0 @ HL.T + noise ** 2 * np.eye(1)) @ Hfi.T) < 1e-8
    assert var.idx() == n
    # --- general (tall H_new): any orthogonal split must give the same posterior
    SM.marginalize(st, var)
    P0 = SM.get_full_cov(st)
    m = 5
    Hx = rng.uniform(-1, 1, (m, 10))
    Hf = np.ones((m, 1))
    res = 0.01 * rng.uniform(-1, 1, m)
    var = o.Scalar()
    Hx_in, Hf_in, res_in = Hx.copy(), Hf.copy(), res.copy()
    ok = SM.add_variable_delayed(st, var, order, Hx_in, Hf_in, res_in, noise, 1e6, True)
    assert ok
    # reference check: QR of Hf by hand
    Q, _ = np.linalg.qr(Hf, mode="complete")
    HxQ, HfQ, rQ = Q.T @ Hx, Q.T @ Hf, Q.T @ res
    HL = np.zeros((m, n))
    HL[:, :9] = HxQ[:, :9]
    HL[:, st.gnss[YOF].idx()] = HxQ[:, 9]
    Hfi = np.linalg.inv(HfQ[:1, :1])
    Pa = np.zeros((n + 1, n + 1))
    Pa[:n, :n] = P0
    Pa[:n, n:] = -P0 @ HL[:1].T @ Hfi.T
    Pa[n:, :n] = Pa[:n, n:].T
    Pa[n:, n:] = Hfi @ (HL[:1] @ P0 @ HL[:1].T + noise ** 2) @ Hfi.T
    Hup = np.zeros((m - 1, n + 1))
    Hup[:, :n] = HL[1:]
    K = Pa @ Hup.T @ np.linalg.inv(Hup @ Pa @ Hup.T + noise ** 2 * np.eye(m - 1))
    ref = (np.eye(n + 1) - K @ Hup) @ Pa
    assert np.linalg.norm(ref - SM.get_full_cov(st)) < 1e-8 * max(1, np.linalg.norm(ref))


def test_replace_var_linear():
    """TestMapServer.cpp:527-551: replaceVarLinear == H P H^T block replacement (1e-10)."""
    rng = np.random.default_rng(8)
    fp, st, Phi, G = _fixture_state(rng)
    st.timestamp = 1.0
    SM.propagate_state_cov(st, Phi, G, 1.5)
    st.timestamp = 2.5
    SM.augment_sliding_window_pose(st)
    lm = o.AnchoredLandmark()
    lm.reset_anchored_pose(st.sw_camleft_poses[2.5])
    SM.add_anchored_landmark_in_state(st, lm, 1, 3.0 * np.eye(3))
    dep = [lm, st.sw_camleft_poses[2.5]]
    H = rng.uniform(-1, 1, (3, 9))
    P0 = SM.get_full_cov(st)
    n = P0.shape[0]
    HL = np.zeros((3, n))
    HL[:, lm.idx():lm.idx() + 3] = H[:, :3]
    c = st.sw_camleft_poses[2.5].idx()
    HL[:, c:c + 6] = H[:, 3:]
    SM.replace_var_linear(st, lm, dep, H)
    T = np.eye(n)
    T[lm.idx():lm.idx() + 3, :] = HL
    ref = T @ P0 @ T.T
    assert np.linalg.norm(ref - SM.get_full_cov(st)) < 1e-10 * max(1, np.linalg.norm(ref))


def test_one_step_prop_convergence():
    """TestPropagator.cpp:114-189: analytic vs RK4/Taylor transition error decreases monotonically
    as dt = 1, 0.1, ..., 1e-4 (ordering only)."""
    rng = np.random.default_rng(9)
    errs_state, errs_phi = [], []
    w = rng.uniform(-1, 1, 3)
    a = rng.uniform(-1, 1, 3) + np.array([0, 0, 9.8])
    R0 = rand_rot(rng)
    p0, v0 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    for dt in (1.0, 0.1, 0.01, 1e-3, 1e-4):
        out = []
        for analytic in (True, False):
            st = o.State(o.FilterParams(enable_gnss=0))
            st.init_state_and_cov(0.0, R0, p0, v0, np.zeros(3), np.zeros(3))
            prop = o.ImuPropagator(9.8)
            Phi, G = prop.state_and_cov_transition(st, o.ImuCtrl(0.0, w, a), dt, analytic)
            out.append((st.extended_pose.rot.copy(), st.extended_pose.vec1.copy(),
                        st.extended_pose.vec2.copy(), Phi))
        (Ra, pa, va, Pa), (Rn, pn, vn, Pn) = out
        errs_state.append(np.linalg.norm(Ra - Rn) + np.linalg.norm(pa - pn) + np.linalg.norm(va - vn))
        errs_phi.append(np.linalg.norm(Pa - Pn))
    assert all(errs_state[i] > errs_state[i + 1] for i in range(len(errs_state) - 1))
    assert all(errs_phi[i] > errs_phi[i + 1] for i in range(len(errs_phi) - 1))


def test_phi_matches_numerical_jacobian():
    """Independent pin of ImuPropagator.cpp:150-161: Phi is the Jacobian of the left-invariant error
    of the analytic mean propagation (finite differences)."""
    rng = np.random.default_rng(10)
    w = rng.uniform(-0.5, 0.5, 3)
    a = rng.uniform(-1, 1, 3) + np.array([0, 0, 9.8])
    dt = 0.05
    R0 = rand_rot(rng)
    p0, v0 = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    bg0, ba0 = 0.01 * rng.uniform(-1, 1, 3), 0.1 * rng.uniform(-1, 1, 3)

    def run(dx):
        st = o.State(o.FilterParams(enable_gnss=0))
        st.init_state_and_cov(0.0, R0, p0, v0, bg0, ba0)
        full = np.zeros(21)
        full[:15] = dx
        SM.box_plus(st, full)
        prop = o.ImuPropagator(9.8)
        Phi, _ = prop.state_and_cov_transition(st, o.ImuCtrl(0.0, w, a), dt, True)
        e = st.extended_pose
        return e.rot.copy(), e.vec1.copy(), e.vec2.copy(), st.bg.value().copy(), st.ba.value().copy(), Phi

    Rn, pn, vn, bgn, ban, Phi = run(np.zeros(15))
    eps = 1e-6
    J = np.zeros((15, 15))
    for k in range(15):
        d = np.zeros(15)
        d[k] = eps
        R, p, v, bg, ba, _ = run(d)
        # left-invariant error: X_pert = Exp(xi) X_nom  ->  dR = R Rn^T, dp = p - dR pn, dv = v - dR vn
        dR = R @ Rn.T
        th = o.vee(dR - dR.T) / 2.0
        J[:, k] = np.r_[th, p - dR @ pn, v - dR @ vn, bg - bgn, ba - ban] / eps
    D = J - Phi
    # The reference evaluates Psi1/Psi2 as M1*(c1*WA+...) (AuxGammaFunc.cpp:163,221), which leaves an
    # O(|a| dt^2) gap to the true bias-gyro -> velocity/position sensitivity. The oracle reproduces
    # the reference (that is what parity means); those two blocks are therefore only bounded here.
    bg_gap = D[3:9, 9:12].copy()
    D[3:9, 9:12] = 0.0
    assert np.linalg.norm(D) < 5e-5 * max(1, np.linalg.norm(Phi))
    assert np.linalg.norm(bg_gap) < 2.0 * np.linalg.norm(a) * dt ** 2
