"""GPU parity, round-2 additions: the dispatch paths and branches the first round's cases did not reach.

* the per-track kernel's FUSED instance at its natural dispatch (B >= 296, the benchmarked path) vs the oracle;
* a 500-frame c2 run at the <= 1e-6 drift bar of SURVEY.md §8d / BASELINE.md §4;
* SwMarg with stereo observations;
* the small-angle branches of Gamma / Psi (AuxGammaFunc.cpp:53,119,172) through an IMU propagation with |w dt| below the
  cut-offs;
* propagation at the largest state dimension igv_create accepts (k_propagate's strip);
* two handles on two devices in one process (per-device kernel attributes, device guard on every entry point).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ingvio_oracle as o
from ingvio_oracle import StateManager as SM

from helpers import (TiledStream, assert_state_close, block_rel_error, filter_params, gstep, make_gpu, make_oracles,
                     oracle_blocks)
from ingvio_b200.synth import WORKLOADS, SyntheticStream


def test_fused_natural_dispatch_c2_b296():
    """c2 at B = 296 (4 distinct streams tiled): the dispatcher picks k_msckf_features<FUSE> on its own
    (igv_last_visual_path() == 2); the 4 distinct sequences and a tiled copy of each are compared with the oracle."""
    wl = WORKLOADS["c2"]
    fp = filter_params(wl)
    inner = SyntheticStream(wl, 4)
    orc = make_oracles(wl, inner, fp)
    st = TiledStream(inner, 296)
    g = make_gpu(wl, st, fp)
    saw_fused = False
    for i in range(13):
        fr = st.next_frame()
        out = g.step(fr, noise=fp.visual_noise, want=True)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
        if "visual" in out:
            saw_fused = saw_fused or g.last_visual_path() == 2
            for b, f in enumerate(orc):
                acc = sum(1 for x in f.last["gammas"] if x[3])
                assert out["visual"]["accepted"][b] == acc and out["visual"]["accepted"][292 + b] == acc
    assert saw_fused, "the fused per-track kernel was not dispatched at B = 296"
    P, X = g.get_full_cov(), g.get_state()
    for b, f in enumerate(orc):
        Po = f.cov()
        for bb in (b, 292 + b):
            err = np.linalg.norm(P[bb] - Po) / max(1.0, np.linalg.norm(Po))
            assert err <= 1e-8, (bb, err)
            berr, where = block_rel_error(P[bb], Po, oracle_blocks(f))
            assert berr <= 1e-7, (bb, where, berr)
        assert np.array_equal(P[b], P[292 + b]) and np.array_equal(X[b], X[292 + b])   # schedule independent, bitwise


def test_c2_500_frames_drift():
    """500 full frame cycles of c2 (propagate x10, augment, MSCKF, GNSS, marginalise): the CUDA path stays within
    1e-6 of the oracle (<= 1e-9 per frame, <= 1e-6 after 500 frames)."""
    wl = WORKLOADS["c2"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    worst = 0.0
    for i in range(500):
        fr = st.next_frame()
        gstep(g, fr, fp)
        orc[0].step(fr.seq(0))
        if i % 50 == 49 or i == 499:
            P, Po = g.get_full_cov()[0], orc[0].cov()
            err = np.linalg.norm(P - Po) / max(1.0, np.linalg.norm(Po))
            worst = max(worst, err)
            assert err <= 1e-6, f"frame {i}: |dP| = {err:.3e}"
    assert_state_close(g, orc, wl.sw, tol_P=1e-6, tol_x=1e-6, tol_block=1e-5, what="c2 after 500 frames")
    R, p, v = orc[0].pose()
    x = g.get_state()[0]
    assert np.max(np.abs(x[9:12] - p)) <= 1e-6
    tr, tro = g.cov_trace()[0], np.trace(orc[0].cov())
    assert abs(tr - tro) <= 1e-6 * tro
    print(f"c2 500 frames: worst |dP|_F/max(1,|P|_F) = {worst:.3e}, |dp| = {np.max(np.abs(x[9:12] - p)):.3e}")


def test_sw_marg_stereo():
    """SwMargUpdate::updateStateStereo (SwMargUpdate.cpp:216-365): selected clones with stereo rows, anchors inside and
    outside the selection."""
    wl = WORKLOADS["tiny_stereo"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    used = False
    for i in range(8):
        fr = st.next_frame()
        if fr.visual_mode is not None and int(fr.obs_mask[0, 0].sum()) == wl.sw:
            fr.visual_mode = "sw_marg"
            fr.selected_slots = [0, 1, 3]
            fr.anchor_slot[:, 0::3] = 0
            fr.anchor_slot[:, 1::3] = 2          # anchor outside the selection: its column block is appended
            fr.anchor_slot[:, 2::3] = 3
            fr.marg_slots = [0]
            out = g.step(fr, noise=fp.visual_noise, want=True)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
                for fid, gam, dof, ok in f.last["gammas"]:
                    assert abs(out["visual"]["gamma"][b, fid] - gam) <= 1e-8 * max(1, abs(gam))
                assert out["visual"]["accepted"][b] == sum(1 for x in f.last["gammas"] if x[3])
            st.n_clones = g.num_clones()
            used = True
        else:
            g.step(fr, noise=fp.visual_noise)
            for b, f in enumerate(orc):
                f.step(fr.seq(b))
        assert_state_close(g, orc, wl.sw, what=f"sw_marg stereo frame {i}")
    assert used


@pytest.mark.parametrize("scale", [1e-5, 1e-7, 1e-9, 0.0])
def test_small_angle_gamma_psi_branches(scale):
    """|w| dt below the cut-offs of GammaFunc (1e-6), Psi1 (1e-8) and Psi2 (1e-7) (AuxGammaFunc.cpp:53,119,172): the
    device takes the same series branches as the oracle."""
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    fr = st.next_frame(with_visual=False, with_gnss=False, marg_oldest=False)
    rng = np.random.default_rng(5)
    # the biases are subtracted from the raw rate: aim the UNBIASED rate at the wanted magnitude
    ini = st.initial_state()
    dirn = rng.standard_normal(fr.gyro.shape)
    dirn /= np.linalg.norm(dirn, axis=-1, keepdims=True)
    gyro = ini["bg"][:, None, :] + scale * dirn / fr.dt[..., None]
    g.propagate_imu(gyro, fr.accel, fr.dt)
    for b, f in enumerate(orc):
        f.prop.propagate_steps(f.state, gyro[b], fr.accel[b], fr.dt[b])
    assert_state_close(g, orc, wl.sw, tol_P=1e-10, what=f"small angle {scale}")


def test_propagate_at_max_dim():
    """k_propagate's strip needs 20 * N * 8 bytes of shared memory: N = 411 (64 clones + 6 GNSS scalars) must work,
    and the result must equal the oracle's propagateStateCov on the same (random, dense) covariance."""
    from ingvio_b200.filter import BatchFilter
    rng = np.random.default_rng(3)
    B, SW = 2, 64
    g = BatchFilter(B, SW, 1, 1)
    eye = np.tile(np.eye(3).reshape(1, 9), (B, 1))
    z3 = np.zeros((B, 3))
    g.init_state_and_cov(eye, z3, z3, z3, z3, eye, z3, np.full(21, 1e-2))
    for gt in range(6):
        g.add_gnss_variable(gt, 0.0, 1.0)
    for _ in range(SW):
        g.augment_sliding_window_pose()
    N = g.curr_cov_size()
    assert N == 21 + 6 + 6 * SW
    A = rng.standard_normal((B, N, N)) * 0.1
    P0 = A @ A.transpose(0, 2, 1) + np.eye(N)
    g.set_full_cov(P0)
    Phi = rng.uniform(-1, 1, (B, 15, 15))
    G = rng.uniform(-1, 1, (B, 15, 12))
    dt = np.array([0.7, 0.02])
    g.propagate_state_cov(Phi, G, dt)
    P = g.get_full_cov()
    # oracle: a State with the same variable order
    fp = o.FilterParams(max_sw_clones=SW, enable_gnss=1)
    for b in range(B):
        f = o.OracleFilter(fp, stereo=False, max_valid_ids=1)
        f.init(0.0, np.eye(3), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3))
        for gt in range(6):
            SM.add_gnss_variable(f.state, gt, 0.0, 1.0)
        for k in range(SW):
            f.state.timestamp = 0.05 * (k + 1)
            SM.augment_sliding_window_pose(f.state)
        f.state.cov[:, :] = P0[b]
        SM.propagate_state_cov(f.state, Phi[b], G[b], dt[b])
        err = np.linalg.norm(P[b] - f.cov()) / max(1.0, np.linalg.norm(f.cov()))
        assert err <= 1e-10, err


def test_two_devices_one_process():
    """One handle per device in ONE process: kernel attributes are per device and every entry point runs on its handle's
    device whatever the caller's current device is."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from ingvio_b200.filter import BatchFilter
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    gs = [make_gpu(wl, st, fp, device=d) for d in (0, 1)]
    for i in range(6):
        fr = st.next_frame()
        for d, g in enumerate(gs):
            torch.cuda.set_device(1 - d)       # the caller's current device is the OTHER one
            gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    for g in gs:
        assert_state_close(g, orc, wl.sw, what="two devices")


@pytest.mark.parametrize("adjust_yof", [0, 1])
def test_gnss_add_new_tracked_sys(adjust_yof):
    """GnssUpdate::addNewTrackedSys (GnssUpdate.cpp:317-476): a clock bias of a constellation seen for the first time and
    the clock drift FS are added through the device-built rows + delayed initialisation; ragged satellite counts per
    sequence, an outlier epoch that the 0.95 gate rejects, and the reference's R_ecef2enu quirk in the yaw-offset column."""
    from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF
    from ingvio_oracle.gnss_update import GnssEpoch
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl, is_adjust_yof=adjust_yof)
    st = SyntheticStream(wl, 3)
    orc = make_oracles(wl, st, fp, with_gnss=False)
    g = make_gpu(wl, st, fp, with_gnss=False)
    for gt, val, cov in ((YOF, 0.3, 0.015 ** 2), (GPS, 1.0, 4.0)):
        g.add_gnss_variable(gt, val, cov)
        for f in orc:
            SM.add_gnss_variable(f.state, gt, val, cov)
    for _ in range(5):                       # realistic cross-covariances first
        fr = st.next_frame(with_gnss=False)
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    fr = st.next_frame(with_visual=False)
    G = fr.gnss
    S = G.unit.shape[1]
    G.sys[:, :] = np.array([GPS, GAL, GAL, GLO, GAL] + [GLO] * (S - 5))[None, :S]
    G.sys[1, 1] = GLO                        # sequence 1 sees one GAL satellite fewer (ragged rows)
    G.res_pos[2] += 3000.0                   # sequence 2: inconsistent epoch, the 0.95 gate must reject GAL
    R_e2n = np.transpose(G.R_enu2ecef, (0, 2, 1)).copy()
    R_e2n[:, 0, 1] += 1e-3                   # make the reference's getRecef2enu() quirk observable
    spp = {GAL: 0.7, FS: 0.05}
    for gt in (GAL, FS):
        acc = g.gnss_add_new_tracked_sys(gt, spp[gt], G.unit, G.res_pos, G.res_vel, G.sigma_psr(fp.psr_noise_amp),
                                         G.sigma_dopp(fp.dopp_noise_amp), G.sys, G.R_enu2ecef.reshape(-1, 9),
                                         R_ecef2enu=R_e2n.reshape(-1, 9), is_adjust_yof=adjust_yof,
                                         prior_cov_if_rejected=1.0)
        for b, f in enumerate(orc):
            ep = GnssEpoch(unit=G.unit[b], res_pos=G.res_pos[b], res_vel=G.res_vel[b], sys=G.sys[b], ura=G.ura[b],
                           psr_std=G.psr_std[b], dopp_std_mps=G.dopp_std_mps[b], el=G.el[b])
            out = f.gnss.add_new_tracked_sys(f.state, ep, G.R_enu2ecef[b], [gt], spp, R_ecef2enu=R_e2n[b])
            assert bool(acc[b]) == bool(out[gt]), (gt, b, acc[b], out)
            if not out[gt]:
                # batch semantics: a rejected sequence keeps a decoupled scalar (documented deviation); mirror it
                SM.add_gnss_variable(f.state, gt, spp[gt], 1.0)
        assert_state_close(g, orc, wl.sw, what=f"add new sys {gt}")
    assert not acc is None


@pytest.mark.parametrize("wname", ["c2", "c3", "c5"])
def test_triangulate_group_kernel_vs_thread_kernel(wname, monkeypatch):
    """k_triangulate_grp (a group of lanes per track, the default) against k_triangulate (one thread per track,
    IGV_TRI_CFG=1) at the BASELINE window sizes -- 16-lane groups at c2 (11 views), 32-lane groups at c3 (22 stereo
    views) and at c5 (30 views): identical accept / reject decisions and the same positions to 1e-9 m, including ragged
    tracks, tracks with too few views and garbage tracks that fail the depth gates; the thread kernel itself is held to
    the oracle by test_gpu_parity.py::test_triangulate_matches_oracle."""
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    B = 2
    st = SyntheticStream(wl, B)
    frames = []
    for _ in range(wl.sw + 1):
        frames.append(st.next_frame())
    fr = frames[-1]
    rng = np.random.default_rng(9)
    mask = fr.obs_mask.copy()
    mask[:, 0::7, :3] = 0                    # ragged
    mask[:, 1, :] = 0
    mask[:, 1, :2] = 1                       # too few views
    obs = fr.obs.copy()
    obs[:, 2::11] += rng.normal(0, 0.3, obs[:, 2::11].shape)   # garbage tracks
    res = {}
    for cfg in ("0", "1"):
        monkeypatch.setenv("IGV_TRI_CFG", cfg)
        g = make_gpu(wl, SyntheticStream(wl, B), fp)
        s2 = SyntheticStream(wl, B)
        for _ in range(wl.sw + 1):
            gstep(g, s2.next_frame(), fp)
        g.augment_sliding_window_pose() if g.num_clones() < wl.sw else None
        res[cfg] = g.triangulate(obs, mask, fr.anchor_slot, trans_thres=0.1, conv_precision=5e-7, max_depth=60.0)
        g.close()
    (pf0, ok0), (pf1, ok1) = res["0"], res["1"]
    assert np.array_equal(ok0, ok1)
    assert ok0.sum() > 0.5 * ok0.size and not ok0[:, 1].any()
    assert np.abs(pf0 - pf1)[ok0].max() <= 1e-9


@pytest.mark.parametrize("graph", ["1", "0"])
def test_frame_step_graph_replay(graph, monkeypatch):
    """igv_frame_step: a whole frame cycle behind one call. Device-resident argument buffers at fixed addresses are
    refilled every frame; from the third frame of a steady window the kernel sequence is a CUDA-graph replay
    (two graphs alternate with the covariance ping-pong). Every frame is compared with the oracle, and the graph run
    must equal the call-by-call run bit for bit."""
    import torch
    from ingvio_b200 import capi
    monkeypatch.setenv("IGV_GRAPH", graph)
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    B = 2
    st = SyntheticStream(wl, B)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    gref = make_gpu(wl, SyntheticStream(wl, B), fp)
    dev = torch.device("cuda:0")
    buf = None
    replays_seen = 0
    for i in range(wl.sw + 12):
        fr = st.next_frame()
        gstep(gref, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
        G = fr.gnss
        host = dict(gyro=fr.gyro, accel=fr.accel, dt=fr.dt, pf_w=fr.pf_w, anchor_slot=fr.anchor_slot.astype(np.int32),
                    obs=fr.obs, obs_mask=fr.obs_mask.astype(np.uint8), chi2_dof=(fr.obs_total.astype(np.int32) - 1),
                    unit=G.unit, res_pos=G.res_pos, res_vel=G.res_vel, sigma_psr=G.sigma_psr(fp.psr_noise_amp),
                    sigma_dopp=G.sigma_dopp(fp.dopp_noise_amp), sys=G.sys.astype(np.int32), R_enu2ecef=G.R_enu2ecef.reshape(B, 9))
        if buf is None or any(buf[k].shape != torch.Size(v.shape) for k, v in host.items()):
            buf = {k: torch.as_tensor(np.ascontiguousarray(v)).to(dev) for k, v in host.items()}     # (re)allocated: new addresses
        else:
            for k, v in host.items():
                buf[k].copy_(torch.as_tensor(np.ascontiguousarray(v)))                               # same addresses, new contents
        vis = None
        if fr.visual_mode is not None:
            vis = dict(mode=capi.VIS_ALL_OBS, pf_w=buf["pf_w"], anchor_slot=buf["anchor_slot"], obs=buf["obs"], obs_mask=buf["obs_mask"],
                       chi2_dof=buf["chi2_dof"], noise=fp.visual_noise, max_valid=fr.max_valid)
        gn = dict(unit=buf["unit"], res_pos=buf["res_pos"], res_vel=buf["res_vel"], sigma_psr=buf["sigma_psr"], sigma_dopp=buf["sigma_dopp"],
                  sys=buf["sys"], R_enu2ecef=buf["R_enu2ecef"], is_adjust_yof=fp.is_adjust_yof, chi2_test=fp.gnss_chi2_test,
                  strong_reject=fp.gnss_strong_reject)
        torch.cuda.synchronize()
        g.frame_step(buf["gyro"], buf["accel"], buf["dt"], visual=vis, marg_slots=sorted(fr.marg_slots, reverse=True), gnss=gn)
        g.synchronize()
        assert_state_close(g, orc, wl.sw, what=f"frame_step frame {i} (graph={graph})")
        assert np.array_equal(g.get_full_cov(), gref.get_full_cov()) and np.array_equal(g.get_state(), gref.get_state())
        assert g.curr_cov_size() == gref.curr_cov_size() and g.num_clones() == gref.num_clones()
        replays_seen = g.graph_replays
    if graph == "1":
        assert replays_seen >= 6, replays_seen        # steady state was served by graph launches
    else:
        assert replays_seen == 0


@pytest.mark.parametrize("stereo,sw", [(True, 9), (False, 18)])
def test_large_gate_blocked_cholesky(stereo, sw):
    """Gates with more than 32 rows (stereo windows from 9 clones, mono from 18: the 41-row stereo gate of c3, the 57-row
    gate of c5) run the warp-level blocked Cholesky (register diagonal blocks, DMMA trailing updates): gate statistic of
    every track and the state against the oracle, small enough to also run on the CPU model."""
    from ingvio_b200.synth import Workload
    wl = Workload("big_gate", 31 + int(stereo), sw, 6, 0, stereo=stereo)
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp, with_gnss=False)
    g = make_gpu(wl, st, fp, with_gnss=False)
    for i in range(sw + 2):
        fr = st.next_frame(with_gnss=False)
        out = g.step(fr, noise=fp.visual_noise, want=True)
        orc[0].step(fr.seq(0))
        if "visual" in out:
            for fid, gam, dof, ok in orc[0].last["gammas"]:
                assert abs(out["visual"]["gamma"][0, fid] - gam) <= 1e-8 * max(1, abs(gam)), (i, fid, out["visual"]["gamma"][0, fid], gam)
        assert_state_close(g, orc, wl.sw, what=f"frame {i}")
