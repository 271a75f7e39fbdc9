"""GPU parity: CUDA path (through the C-ABI) vs the numpy oracle on the same seeded inputs.

Bars (the reference's own unit-test tolerances): per operation
|dP|_F <= 1e-8 max(1,|P|_F), |d dx| <= 1e-8, state <= 1e-9 relative. Integer outputs (accept counts,
layout indices) are exact.
"""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ingvio_oracle as o
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF, StateManager as SM

from helpers import (GNSS_INIT, assert_state_close, filter_params, gstep, make_gpu, make_oracles,
                     oracle_packed_state, rand_rot)
from ingvio_b200 import capi
from ingvio_b200.synth import WORKLOADS, SyntheticStream


def _warm(wl, B, n_frames, fp=None, seq0=0, **gpu_kw):
    fp = fp or filter_params(wl)
    st = SyntheticStream(wl, B, seq0=seq0)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp, **gpu_kw)
    for _ in range(n_frames):
        fr = st.next_frame()
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    return fp, st, orc, g


def test_layout_and_init():
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 3)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    assert g.curr_cov_size() == 27 and g.curr_err_variable_size() == 10
    for gt in range(6):
        assert g.gnss_idx(gt) == orc[0].state.gnss[gt].idx()
    assert_state_close(g, orc, wl.sw, what="init")
    # add/marg order semantics of TestStateManager.cpp:67-91
    g.marg_gnss_variable(GPS)
    for f in orc:
        SM.marg_gnss_variable(f.state, GPS)
    assert g.curr_cov_size() == 26
    g.add_gnss_variable(GPS, 7.0, 9.0)
    for f in orc:
        SM.add_gnss_variable(f.state, GPS, 7.0, 9.0)
    for gt in range(6):
        assert g.gnss_idx(gt) == orc[0].state.gnss[gt].idx()
    assert_state_close(g, orc, wl.sw, what="add/marg gnss")


def test_propagate_cov_random_phi():
    """StateManager::propagateStateCov with random Phi/G (the reference's own test shape,
    TestStateManager.cpp:93-137), GNSS scalars in a scrambled order, dense prior."""
    rng = np.random.default_rng(11)
    wl = WORKLOADS["tiny"]
    fp, st, orc, g = _warm(wl, 2, 5)
    B = 2
    Phi = rng.uniform(-1, 1, (B, 15, 15))
    G = rng.uniform(-1, 1, (B, 15, 12))
    dt = np.array([1.5, 0.01])
    g.propagate_state_cov(Phi, G, dt)
    for b, f in enumerate(orc):
        SM.propagate_state_cov(f.state, Phi[b], G[b], dt[b])
    assert_state_close(g, orc, wl.sw, tol_P=1e-10, what="propagate_cov")


def test_propagate_without_fs_and_without_gnss():
    rng = np.random.default_rng(12)
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp, with_gnss=False)
    g = make_gpu(wl, st, fp, with_gnss=False)
    for gt, val, cov in ((BDS, 1.0, 4.0), (GPS, 2.0, 4.0)):   # clock biases but no FS
        g.add_gnss_variable(gt, val, cov)
        SM.add_gnss_variable(orc[0].state, gt, val, cov)
    Phi = rng.uniform(-1, 1, (1, 15, 15))
    G = rng.uniform(-1, 1, (1, 15, 12))
    g.propagate_state_cov(Phi, G, np.array([0.3]))
    SM.propagate_state_cov(orc[0].state, Phi[0], G[0], 0.3)
    assert_state_close(g, orc, wl.sw, tol_P=1e-10, what="propagate no FS")


def test_imu_propagate_and_augment():
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    fr = st.next_frame(with_visual=False, with_gnss=False, marg_oldest=False)
    dt = fr.dt.copy()
    dt[:, 3] = 1e-7   # skipped step (ImuPropagator.cpp:262)
    g.propagate_imu(fr.gyro, fr.accel, dt)
    for b, f in enumerate(orc):
        f.prop.propagate_steps(f.state, fr.gyro[b], fr.accel[b], dt[b])
    assert_state_close(g, orc, wl.sw, tol_P=1e-10, what="imu propagate")
    g.augment_sliding_window_pose()
    for f in orc:
        f.state.timestamp = fr.t
        SM.augment_sliding_window_pose(f.state)
    assert g.num_clones() == 1 and g.clone_idx(0) == 27
    assert_state_close(g, orc, wl.sw, tol_P=1e-10, what="augment")


def test_marginalize_clone_and_gnss():
    wl = WORKLOADS["tiny"]
    fp, st, orc, g = _warm(wl, 2, 3)   # 3 clones, no marg yet (sw=4)
    assert g.num_clones() == 3
    g.marg_sliding_window_pose(1)
    for f in orc:
        SM.marg_sliding_window_pose(f.state, f.state.sw_times()[1])
    g.marg_gnss_variable(GLO)
    for f in orc:
        SM.marg_gnss_variable(f.state, GLO)
    assert_state_close(g, orc, wl.sw, tol_P=1e-12, tol_x=1e-12, what="marginalize")
    assert g.clone_idx(1) == orc[0].state.sw_camleft_poses[orc[0].state.sw_times()[1]].idx()


@pytest.mark.parametrize("r_kind", ["iso", "diag", "full"])
def test_ekf_update_sparse_var_order(r_kind):
    """StateManager::ekfUpdate against the oracle AND the dense textbook form (TestStateManager.cpp:478-557)."""
    rng = np.random.default_rng(13)
    wl = WORKLOADS["tiny"]
    fp, st, orc, g = _warm(wl, 2, 4)
    s0 = orc[0].state
    var_objs = lambda s: [s.extended_pose, s.gnss[GPS], s.sw_camleft_poses[s.sw_times()[1]], s.gnss[FS], s.ba]
    order = [(v.idx(), v.size()) for v in var_objs(s0)]
    n = sum(sz for _, sz in order)
    rows = 7
    H = rng.uniform(-1, 1, (2, rows, n))
    res = 0.01 * rng.uniform(-1, 1, (2, rows))
    if r_kind == "iso":
        R_gpu, R_or = 0.25, lambda b: 0.25 * np.eye(rows)
    elif r_kind == "diag":
        d = rng.uniform(0.1, 1.0, (2, rows))
        R_gpu, R_or = d, lambda b: np.diag(d[b])
    else:
        A = rng.uniform(-1, 1, (2, rows, rows))
        Rf = A @ np.swapaxes(A, -1, -2) + 0.5 * np.eye(rows)
        R_gpu, R_or = Rf, lambda b: Rf[b]
    P0 = g.get_full_cov()
    gam = g.whiten_residual(order, H, res, R_gpu)
    dx = g.ekf_update(order, H, res, R_gpu)
    for b, f in enumerate(orc):
        N = f.state.curr_cov_size()
        HL = np.zeros((rows, N))
        c = 0
        for i0, sz in order:
            HL[:, i0:i0 + sz] = H[b][:, c:c + sz]
            c += sz
        S = HL @ P0[b] @ HL.T + R_or(b)
        K = P0[b] @ HL.T @ np.linalg.inv(S)
        ref = (np.eye(N) - K @ HL) @ P0[b]
        gam_ref = float(res[b] @ np.linalg.solve(S, res[b]))
        dxo, _ = SM.ekf_update(f.state, var_objs(f.state), H[b], res[b], R_or(b), return_dx=True)
        assert np.linalg.norm(g.get_full_cov()[b] - ref) <= 1e-8 * max(1, np.linalg.norm(ref))
        assert np.linalg.norm(dx[b] - dxo) <= 1e-8
        assert abs(gam[b] - gam_ref) <= 1e-9 * max(1, abs(gam_ref))
    assert_state_close(g, orc, wl.sw, what=f"ekf_update {r_kind}")
    Pn = g.get_full_cov()
    assert np.array_equal(Pn, np.swapaxes(Pn, -1, -2)), "posterior must be exactly symmetric"


def _compare_visual(g, orc, fr, fp, want_cap=None):
    out = gstep(g, fr, fp, want=True)
    gam_gpu = out["visual"]["gamma"]
    acc_gpu = out["visual"]["accepted"]
    for b, f in enumerate(orc):
        dxv, dxg = f.step(fr.seq(b))
        gl = f.last["gammas"]
        n_acc = sum(1 for x in gl if x[3])
        if want_cap:
            n_acc = min(n_acc, want_cap)
        assert acc_gpu[b] == n_acc, (b, acc_gpu[b], n_acc)
        for fid, gam, dof, ok in gl:
            assert abs(gam_gpu[b, fid] - gam) <= 1e-8 * max(1.0, abs(gam)), (fid, gam_gpu[b, fid], gam)
        if dxv is not None:
            assert np.linalg.norm(out["visual"]["dx"][b][:len(dxv[0])] - dxv[0]) <= 1e-8
    return out


@pytest.mark.parametrize("wname", ["tiny", "tiny_stereo"])
def test_msckf_all_obs_frames(wname):
    """RemoveLost-style update over whole frames (propagate, augment, visual, marg, GNSS)."""
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(8):
        fr = st.next_frame()
        if fr.visual_mode is None:
            g.step(fr, noise=fp.visual_noise)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
        else:
            _compare_visual(g, orc, fr, fp)
        assert_state_close(g, orc, wl.sw, what=f"{wname} frame {i}")
    # the joint GNSS gate may legitimately fire on 10 rows (GnssUpdate.cpp:286); nothing else may
    assert np.all((g.flags() & ~(capi.FLAG_GNSS_REJECTED | capi.FLAG_WEAK_PIVOT)) == 0)


def test_msckf_ragged_outliers_and_cap():
    """Ragged observation masks, gross outliers (chi^2 rejections), a track with < 2 obs, dof from a
    longer track history, and the max_valid_ids cap (RemoveLostUpdate.h:38)."""
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    rng = np.random.default_rng(5)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(6):
        fr = st.next_frame()
        if fr.visual_mode is not None:
            ncl = int(fr.obs_mask[0, 0].sum())
            m = fr.obs_mask.copy()
            drop = rng.random(m.shape) < 0.25
            m[drop] = 0
            m[:, 0, :] = 0
            m[:, 0, 0] = 1                      # one observation only -> M=2 <= 3 -> skipped
            m[:, 1, :ncl] = 1
            fr.obs_mask = m
            # gross outlier + track history longer than the window (dof rule uses the map size)
            fr.obs[:, 2, :, :] += 0.9
            fr.obs_total = m.sum(-1).astype(np.int32) + 2
            # tracks need >= 2 usable obs for the oracle's SVD path; emulate the reference's >=4 obs rule
            few = m.sum(-1) < 2
            fr.obs_mask[few] = 0
            fr.obs_mask[:, 0, 0] = 1
            fr.max_valid = 5
            # oracle cannot digest a 1-observation track (reference never feeds one): drop it there
            out = g.step(fr, noise=fp.visual_noise, want=True)
            for b, f in enumerate(orc):
                frb = fr.seq(b)
                frb.obs_mask = frb.obs_mask.copy()
                frb.obs_mask[0, :] = 0
                f.propagate_augment(frb)
                ms = f.build_map_server(frb)
                ids = [k for k in sorted(ms) if len(ms[k].mono_obs) - 2 >= 2]
                f.remove_lost.last_gammas = []
                f.remove_lost.max_valid_ids = 5
                f.remove_lost.update_with_ids(f.state, ms, ids, False, keep="cols")
                f.marginalize(frb)
                f.gnss_update(frb)
                gl = f.remove_lost.last_gammas
                # the oracle stops at the cap; the GPU gates every track first
                for fid, gam, dof, ok in gl:
                    assert abs(out["visual"]["gamma"][b, fid] - gam) <= 1e-8 * max(1, abs(gam))
                assert out["visual"]["accepted"][b] == sum(1 for x in gl if x[3])
        else:
            g.step(fr, noise=fp.visual_noise)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
        assert_state_close(g, orc, wl.sw, what=f"ragged frame {i}")


@pytest.mark.parametrize("mode,stereo", [("keyframe", False), ("sw_marg", False), ("keyframe", True)])
def test_msckf_selected_modes(mode, stereo):
    """KeyframeUpdate / SwMargUpdate: selected clones, anchors inside and outside the selection
    (including the reference's anchor-column overwrite)."""
    wl = WORKLOADS["tiny_stereo" if stereo else "tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(7):
        fr = st.next_frame()
        if fr.visual_mode is not None and int(fr.obs_mask[0, 0].sum()) == wl.sw:
            fr.visual_mode = mode
            fr.selected_slots = [0, 2] if mode == "keyframe" else [0, 1, 3]
            fr.anchor_slot[:, 0::3] = 0      # anchor is a selected clone
            fr.anchor_slot[:, 1::3] = wl.sw - 1 if mode == "keyframe" else 2   # anchor outside the selection
            fr.anchor_slot[:, 2::3] = 2 if mode == "keyframe" else 3
            fr.marg_slots = [2, 0] if mode == "keyframe" else [0]
            out = g.step(fr, noise=fp.visual_noise, want=True)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
                for fid, gam, dof, ok in f.last["gammas"]:
                    assert abs(out["visual"]["gamma"][b, fid] - gam) <= 1e-8 * max(1, abs(gam))
                assert out["visual"]["accepted"][b] == sum(1 for x in f.last["gammas"] if x[3])
            st.n_clones = g.num_clones()
        else:
            g.step(fr, noise=fp.visual_noise)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
        assert_state_close(g, orc, wl.sw, what=f"{mode} frame {i}")


@pytest.mark.parametrize("opts", [dict(), dict(gnss_chi2_test=1), dict(is_adjust_yof=1), dict(few_sats=True)])
def test_gnss_update(opts):
    wl = WORKLOADS["tiny"]
    few = opts.pop("few_sats", False)
    fp = filter_params(wl, **opts)
    fp2, st, orc, g = _warm(wl, 2, 5, fp=fp)
    fr = st.next_frame(with_visual=False)
    if few:
        fr.gnss.sys[:, :] = GPS   # 5 sats of one constellation -> 10 rows <= 14 -> joint gate runs
        fr.gnss.res_pos[1] += 500.0   # sequence 1 must be rejected by the strong gate
    else:
        fr.gnss.res_pos[:, 0] += 80.0  # an outlier for the per-row gate
    gstep(g, fr, fp)
    for b, f in enumerate(orc):
        f.step(fr.seq(b))
    assert_state_close(g, orc, wl.sw, what=f"gnss {opts} few={few}")
    fl = g.flags()
    if few:
        assert fl[1] & capi.FLAG_GNSS_REJECTED and not (fl[0] & capi.FLAG_GNSS_REJECTED)


def test_gnss_untracked_constellation_and_missing_states():
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp, with_gnss=False)
    g = make_gpu(wl, st, fp, with_gnss=False)
    fr = st.next_frame(with_visual=False)
    g.step(fr)                      # no YOF/FS in state: checkGnssStates fails -> no-op
    orc[0].step(fr.seq(0))
    assert_state_close(g, orc, wl.sw, what="gnss without states")
    for gt, val, cov in ((YOF, 0.2, 0.01), (FS, 0.0, 1.0), (GAL, 0.0, 4.0)):
        g.add_gnss_variable(gt, val, cov)
        SM.add_gnss_variable(orc[0].state, gt, val, cov)
    fr = st.next_frame(with_visual=False)   # only GAL satellites contribute
    g.step(fr)
    orc[0].step(fr.seq(0))
    assert_state_close(g, orc, wl.sw, what="gnss one constellation")


def test_delayed_init_and_replace_var_linear():
    rng = np.random.default_rng(21)
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp, with_gnss=False)
    g = make_gpu(wl, st, fp, with_gnss=False)
    for _ in range(3):
        fr = st.next_frame(with_gnss=False)
        g.step(fr)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    g.add_gnss_variable(YOF, 0.1, 0.01)
    for f in orc:
        SM.add_gnss_variable(f.state, YOF, 0.1, 0.01)
    rows = 6
    Hx = rng.uniform(-1, 1, (2, rows, 10))
    Hf = np.ones((rows, 1))
    res = 0.05 * rng.uniform(-1, 1, (2, rows))
    res[1] += np.linspace(-30, 30, rows)          # sequence 1 fails the chi^2
    order = [(0, 9), (g.gnss_idx(YOF), 1)]
    acc = g.add_variable_delayed(GPS, np.array([3.0, 4.0]), order, Hx, Hf, res, 0.4, 0.95, True,
                                 prior_cov_if_rejected=123.0)
    assert list(acc) == [True, False]
    f = orc[0]
    var = o.Scalar()
    var.set_value(3.0)
    ok = SM.add_variable_delayed(f.state, var, [f.state.extended_pose, f.state.gnss[YOF]], Hx[0].copy(), Hf.copy(),
                                 res[0].copy(), 0.4, 0.95, True)
    assert ok
    f.state.gnss[GPS] = var
    P = g.get_full_cov()
    assert np.linalg.norm(P[0] - f.cov()) <= 1e-8 * max(1, np.linalg.norm(f.cov()))
    x = g.get_state()
    assert np.max(np.abs(x[0, :39 + 36] - oracle_packed_state(f, wl.sw)[:39 + 36])) <= 1e-9
    # rejected sequence: variable present but decoupled, prior untouched
    N = P.shape[-1]
    assert P[1][N - 1, N - 1] == 123.0 and np.all(P[1][N - 1, :N - 1] == 0.0)
    # replaceVarLinear (TestMapServer.cpp:527-551) on an opaque 3-dim variable
    g.add_variable_independent(3, 2.0 * np.eye(3))
    lm = o.AnchoredLandmark()
    SM.add_variable_independent(f.state, lm, 2.0 * np.eye(3))
    cl = f.state.sw_camleft_poses[f.state.sw_times()[0]]
    H = rng.uniform(-1, 1, (3, 9))
    P0 = g.get_full_cov()[0]
    g.replace_var_linear((lm.idx(), 3), [(lm.idx(), 3), (cl.idx(), 6)], H)
    SM.replace_var_linear(f.state, lm, [lm, cl], H)
    assert np.linalg.norm(g.get_full_cov()[0] - f.cov()) <= 1e-10 * max(1, np.linalg.norm(f.cov()))
    assert not np.allclose(P0, f.cov())


def test_batch_equals_singles():
    """Sequences in one handle are independent: a batch reproduces the single-sequence runs bit for bit."""
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    _, st3, _, g3 = _warm(wl, 3, 6, fp=fp)
    P3 = g3.get_full_cov()
    X3 = g3.get_state()
    for b in range(3):
        stb = SyntheticStream(wl, 1, seq0=b)
        gb = make_gpu(wl, stb, fp)
        for _ in range(6):
            gb.step(stb.next_frame(), noise=fp.visual_noise)
        assert np.array_equal(gb.get_full_cov()[0], P3[b])
        assert np.array_equal(gb.get_state()[0], X3[b])


def test_c2_frames_against_oracle():
    """BASELINE config c2 (mono, SW=11, 150 feats, 12 sats, N=93): full frame cycles vs the oracle."""
    wl = WORKLOADS["c2"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(13):
        fr = st.next_frame()
        out = g.step(fr, noise=fp.visual_noise, want=True)
        orc[0].step(fr.seq(0))
        if "visual" in out:
            gl = orc[0].last["gammas"]
            assert out["visual"]["accepted"][0] == sum(1 for x in gl if x[3])
        assert_state_close(g, orc, wl.sw, what=f"c2 frame {i}")
    assert g.curr_cov_size() == 87 and np.all((g.flags() & ~capi.FLAG_WEAK_PIVOT) == 0)
    R, p, v = orc[0].pose()
    x = g.get_state()[0]
    assert np.max(np.abs(x[9:12] - p)) <= 1e-9 * max(1, np.max(np.abs(p)))
    assert abs(g.cov_trace()[0] - np.trace(orc[0].cov())) <= 1e-9 * np.trace(orc[0].cov())


def test_c3_stereo_frames_against_oracle():
    wl = WORKLOADS["c3"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(12):
        fr = st.next_frame()
        g.step(fr, noise=fp.visual_noise)
        orc[0].step(fr.seq(0))
    assert_state_close(g, orc, wl.sw, what="c3")


def test_full_size_properties_c2_batch():
    """Size-independent properties at BASELINE size with a batch that needs no oracle run:
    exact symmetry, positive diagonal, trace decrease under the visual update, split-QR == single QR
    (B small -> row-split path) and device-pointer mode == host-pointer mode."""
    import torch
    wl = WORKLOADS["c2"]
    fp = filter_params(wl)
    B = 4
    st = SyntheticStream(wl, B)
    g = make_gpu(wl, st, fp)
    frames = [st.next_frame() for _ in range(14)]
    for fr in frames[:13]:
        g.step(fr, noise=fp.visual_noise)
    fr = frames[13]
    g.propagate_imu(fr.gyro, fr.accel, fr.dt)
    g.augment_sliding_window_pose()
    P_prior = g.get_full_cov()
    X_prior = g.get_state()
    dof = fr.obs_total.astype(np.int32) - 1
    out = g.msckf_update(capi.VIS_ALL_OBS, fr.pf_w, fr.anchor_slot, fr.obs, fr.obs_mask, dof, fp.visual_noise,
                         fr.max_valid, want_dx=True, want_accepted=True)
    P_post = g.get_full_cov()
    X_post = g.get_state()
    assert np.array_equal(P_post, np.swapaxes(P_post, -1, -2))
    d = np.diagonal(P_post, axis1=-2, axis2=-1)
    assert np.all(d > 0)
    assert np.all(np.trace(P_post, axis1=-2, axis2=-1) < np.trace(P_prior, axis1=-2, axis2=-1))
    w = np.linalg.eigvalsh(P_prior - P_post)
    assert np.all(w.min(-1) > -1e-9), "P_prior - P_post must be PSD"
    assert np.all(out["accepted"] > 100)
    # same update with device pointers on a restored prior
    g.set_full_cov(P_prior)
    g.set_state(X_prior)
    dev = torch.device("cuda:0")
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    g.msckf_update(capi.VIS_ALL_OBS, t(fr.pf_w, torch.float64), t(fr.anchor_slot, torch.int32),
                   t(fr.obs, torch.float64), t(fr.obs_mask, torch.uint8), t(dof, torch.int32), fp.visual_noise,
                   fr.max_valid)
    g.synchronize()
    assert np.array_equal(g.get_full_cov(), P_post)
    assert np.array_equal(g.get_state(), X_post)


def test_c5_stress_frames_against_oracle():
    """BASELINE config c5 (SW=30, 400 feats, 20 sats, N=207; stack 22800 x 180): exercises the general-size
    paths (P_s read from global memory, 6 column slots in the QR kernel, EKF operands spilled to the
    global workspace) while the window fills and for two steady-state frames, FP64 bars as everywhere."""
    wl = WORKLOADS["c5"]
    fp = filter_params(wl, chi2_max_dof=200)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(32):
        fr = st.next_frame()
        out = gstep(g, fr, fp, want=(i >= 29))
        orc[0].step(fr.seq(0))
        if i in (4, 12, 20, 26) or i >= 29:
            assert_state_close(g, orc, wl.sw, what=f"c5 frame {i}")
        if i >= 29:
            gl = orc[0].last["gammas"]
            assert out["visual"]["accepted"][0] == sum(1 for x in gl if x[3])
    assert g.curr_cov_size() == 21 + 6 + 6 * 29
    assert np.all((g.flags() & 3) == 0)


@pytest.mark.parametrize("wname", ["c1", "tiny_stereo"])
def test_triangulate_matches_oracle(wname):
    """igv_triangulate (Triangulator.cpp:173-359 + the anchor-depth check) vs the oracle on the filter's own
    clone poses: same accept/reject decisions, world positions to 1e-7 m (both iterate to conv_precision)."""
    import ingvio_oracle as o
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for _ in range(wl.sw + 1):
        fr = st.next_frame()
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    fr = st.next_frame()
    g.propagate_imu(fr.gyro, fr.accel, fr.dt)
    g.augment_sliding_window_pose()
    for b, f in enumerate(orc):
        f.propagate_augment(fr.seq(b))
    rng = np.random.default_rng(3)
    mask = fr.obs_mask.copy()
    mask[:, 0, :2] = 0                       # fewer views
    mask[:, 1, :] = 0
    mask[:, 1, :2] = 1                       # <= 4 mono views -> reject (:183)
    obs = fr.obs.copy()
    obs[:, 2] += rng.normal(0, 0.3, obs[:, 2].shape)   # garbage track: may not converge / fail depth gates
    anchor = fr.anchor_slot.copy()
    prm = dict(trans_thres=0.1, conv_precision=5e-7, max_depth=60.0)
    pf, ok = g.triangulate(obs, mask, anchor, **prm)
    tri = o.Triangulator(o.TriParams(**prm))
    n_ok = 0
    for b, f in enumerate(orc):
        times = f.state.sw_times()
        poses = [(f.state.sw_camleft_poses[t].rot, f.state.sw_camleft_poses[t].vec) for t in times]
        for k in range(wl.feats):
            oko, pfo = tri.triangulate_feature(obs[b, k], mask[b, k], poses, int(anchor[b, k]), wl.stereo,
                                               (fp.T_cl2cr_R, fp.T_cl2cr_p))
            assert bool(ok[b, k]) == bool(oko), (b, k)
            if oko:
                n_ok += 1
                assert np.linalg.norm(pf[b, k] - pfo) <= 1e-7 * max(1.0, np.linalg.norm(pfo)), (b, k, pf[b, k], pfo)
    assert n_ok > wl.feats and not ok[:, 1].any()
    # device-resident chain: triangulate -> msckf_update(pf_w = pf_out, feat_ok = ok_out)
    import torch
    dev = torch.device("cuda:0")
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    pf_d = torch.zeros((2, wl.feats, 3), dtype=torch.float64, device=dev)
    ok_d = torch.zeros((2, wl.feats), dtype=torch.uint8, device=dev)
    obs_d, mask_d, anc_d = t(obs, torch.float64), t(mask, torch.uint8), t(anchor, torch.int32)
    g.triangulate(obs_d, mask_d, anc_d, pf_out=pf_d, ok_out=ok_d, **prm)
    dof = t(mask.sum(-1).astype(np.int32) - 1, torch.int32)
    g.msckf_update(capi.VIS_ALL_OBS, pf_d, anc_d, obs_d, mask_d, dof, fp.visual_noise, 0, feat_ok=ok_d)
    g.synchronize()
    assert np.array_equal(pf_d.cpu().numpy(), pf) and np.array_equal(ok_d.cpu().numpy().astype(bool), ok)
    for b, f in enumerate(orc):
        frb = fr.seq(b)
        frb.obs, frb.obs_mask, frb.pf_w = obs[b], mask[b] * ok[b][:, None], pf[b]
        frb.obs_total = mask[b].sum(-1).astype(np.int32)
        ms = f.build_map_server(frb)
        ids = [k for k in sorted(ms) if ok[b, k]]
        f.remove_lost.last_gammas = []
        f.remove_lost.max_valid_ids = 10 ** 6
        f.remove_lost.update_with_ids(f.state, ms, ids, wl.stereo, keep="cols")
    assert_state_close(g, orc, wl.sw, what="triangulate -> msckf chain")
