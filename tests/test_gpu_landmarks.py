"""SLAM landmarks kept in the state (mono) on the CUDA path vs the oracle's restatement of
LandmarkUpdate.cpp (which tests/test_ref_pin.py pins to the reference's own LandmarkUpdate.cpp): delayed initialisation of
3-dim anchored variables (igv_landmark_init), the per-frame landmark update with its chi^2 gate (igv_landmark_update),
AnchoredLandmark's retraction inside OTHER updates (visual / GNSS), anchor change through replaceVarLinear
(igv_landmark_change_anchor), marginalisation of landmarks and of clones around them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ingvio_oracle as o
from ingvio_oracle import StateManager as SM
from ingvio_oracle import landmark_update as olm
from ingvio_oracle.visual_update import SLAM, FeatureInfo

from helpers import assert_state_close, filter_params, gstep, make_gpu, make_oracles
from ingvio_b200.synth import WORKLOADS, SyntheticStream


def _project(f, pf):
    """Normalised image coordinates of world point pf in the CURRENT left camera of oracle filter f (+ its clones)."""
    st = f.state
    out = {}
    for t in st.sw_times():
        c = st.sw_camleft_poses[t]
        pc = c.value_linear().T @ (pf - c.value_trans())
        out[t] = pc[:2] / pc[2]
    return out


def test_landmark_lifecycle():
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl, max_lm_feats=4)
    B, L = 2, 4
    st = SyntheticStream(wl, B)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp, max_landmarks=L)
    rng = np.random.default_rng(3)
    for _ in range(wl.sw + 2):                   # full window, realistic cross-covariances
        fr = st.next_frame()
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    lms = [o.LandmarkUpdate(fp) for _ in range(B)]
    maps = [dict() for _ in range(B)]
    for f in orc:
        f.state.state_params.max_landmarks = L
    ncl = g.num_clones()
    # ---- delayed initialisation of three landmarks, anchors at different clones; the third is inconsistent in sequence 1 ----
    for k, anchor_slot in enumerate((0, ncl - 1, 2)):
        pf = np.zeros((B, 3))
        obs = np.zeros((B, ncl, 2))
        mask = np.ones((B, ncl), dtype=np.uint8)
        if k == 1:
            mask[:, 1] = 0                       # ragged: not seen at clone 1
        for b, f in enumerate(orc):
            times = f.state.sw_times()
            cam = f.state.sw_camleft_poses[times[-1]]
            pf_true = cam.value_trans() + cam.value_linear() @ np.array([rng.uniform(-2, 2), rng.uniform(-1, 1), rng.uniform(6, 15)])
            pr = _project(f, pf_true)
            for s, t in enumerate(times):
                obs[b, s] = pr[t] + rng.normal(0, 0.01, 2)
            pf[b] = pf_true + rng.normal(0, 0.05, 3)
            if k == 2 and b == 1:
                obs[b, :, 0] += np.linspace(-0.8, 0.8, ncl)      # gross outliers: the 0.95 gate must reject
        acc = g.landmark_init(pf, anchor_slot, obs, mask, fp.visual_noise, prior_cov_if_rejected=1.0)
        for b, f in enumerate(orc):
            times = f.state.sw_times()
            feat = FeatureInfo(100 + k, pf[b], f.state.sw_camleft_poses[times[anchor_slot]])
            for s, t in enumerate(times):
                if mask[b, s]:
                    feat.mono_obs[t] = obs[b, s].copy()
            lm = o.AnchoredLandmark()
            lm.reset_anchored_pose(feat.anchor)
            lm.set_value_pos_xyz(pf[b])
            feat.landmark = lm
            res, Hx, Hf = lms[b].calc_res_jacobian_single_feat_all_mono_obs(feat, f.state)
            sw_vars = [f.state.sw_camleft_poses[t] for t in times]
            ok = SM.add_variable_delayed(f.state, lm, sw_vars, Hx, Hf, res, fp.visual_noise, 0.95, True)
            assert bool(acc[b]) == bool(ok), (k, b, acc[b], ok)
            if not ok:       # batch semantics: a rejected sequence keeps a decoupled landmark with the given prior
                SM.add_variable_independent(f.state, lm, np.eye(3))
            f.state.anchored_landmarks[100 + k] = lm
            feat.ftype = SLAM
            maps[b][100 + k] = feat
        assert g.num_landmarks() == k + 1 and g.landmark_anchor(k) == anchor_slot
        assert g.landmark_idx(k) == orc[0].state.anchored_landmarks[100 + k].idx()
        _check_lm(g, orc, wl, f"landmark init {k}", tol_P=1e-7 if k == 0 else 1e-7)
    assert not acc[1] and acc[0]
    # ---- per-frame landmark update: current observations, one landmark invalid in sequence 0, an outlier in sequence 1 ----
    uv = np.zeros((B, 3, 2))
    valid = np.ones((B, 3), dtype=np.uint8)
    valid[0, 1] = 0
    for b, f in enumerate(orc):
        e, x = f.state.extended_pose, f.state.camleft_imu_extrinsics
        f.state.timestamp = float(f.state.timestamp)
        for l, lid in enumerate((100, 101, 102)):
            pfw = f.state.anchored_landmarks[lid].value_pos_xyz()
            pcl = x.rot.T @ (e.rot.T @ (pfw - e.vec1) - x.vec)
            uv[b, l] = pcl[:2] / pcl[2] + rng.normal(0, 0.02, 2)
        uv[1, 0] += 0.9
    out = g.landmark_update(uv, valid, fp.visual_noise)
    for b, f in enumerate(orc):
        ms = {}
        for l, lid in enumerate((100, 101, 102)):
            if valid[b, l]:
                maps[b][lid].mono_obs = {f.state.timestamp: uv[b, l].copy()}
                ms[lid] = maps[b][lid]
        # the oracle walks state.anchored_landmarks: present only the valid ones (the device skips `valid == 0`)
        saved = dict(f.state.anchored_landmarks)
        f.state.anchored_landmarks = {k_: v for k_, v in saved.items() if k_ in ms}
        lms[b].update_landmark_mono(f.state, ms)
        f.state.anchored_landmarks = saved
        gam = {fid: (gm, ok) for fid, gm, dof, ok in lms[b].last_gammas}
        for l, lid in enumerate((100, 101, 102)):
            if valid[b, l]:
                assert abs(out["gamma"][b, l] - gam[lid][0]) <= 1e-7 * max(1.0, abs(gam[lid][0])), (b, l)
        assert out["accepted"][b] == sum(1 for v in gam.values() if v[1])
    assert out["accepted"][1] < 3                 # the outlier was gated out
    _check_lm(g, orc, wl, "landmark update")
    # ---- anchor change of landmark 0 (anchored at the oldest clone, which is about to leave) to the newest clone ----
    newest = g.num_clones() - 1
    g.landmark_change_anchor(0, newest)
    for b, f in enumerate(orc):
        times = f.state.sw_times()
        olm.change_anchored_pose(maps[b][100], f.state, times[-1])
    assert g.landmark_anchor(0) == newest
    _check_lm(g, orc, wl, "anchor change")
    with pytest.raises(Exception):
        g.marg_sliding_window_pose(2)             # landmark 2 is still anchored there
    # ---- landmarks ride along in the other updates (their retraction uses the anchor's rotation correction); the
    # oldest clone leaves underneath them ----
    fr = st.next_frame()
    gstep(g, fr, fp)
    for b, f in enumerate(orc):
        f.step(fr.seq(b))
    assert g.landmark_anchor(0) == newest - 1 and g.landmark_anchor(1) == newest - 1 and g.landmark_anchor(2) == 1
    _check_lm(g, orc, wl, "visual + GNSS update and clone marginalisation with landmarks in the state")
    # ---- a landmark leaves the state; the remaining ones keep their values and order ----
    g.landmark_marginalize(1)
    for f in orc:
        olm.marg_anchored_landmark_in_state(f.state, 101)
    assert g.num_landmarks() == 2
    _check_lm(g, orc, wl, "landmark marginalised")
    fr = st.next_frame()
    gstep(g, fr, fp)
    for b, f in enumerate(orc):
        f.step(fr.seq(b))
    _check_lm(g, orc, wl, "frames after the landmark life cycle")


def _check_lm(g, orc, wl, what, tol_P=1e-7):
    # landmark covariances are O(1..100) right after initialisation: the global bar is relative to |P|_F
    assert_state_close(g, orc, wl.sw, tol_P=1e-8, tol_block=tol_P, what=what)
    vals = g.landmark_values()
    for b, f in enumerate(orc):
        ref = [lm.value_pos_xyz() for lid, lm in sorted(f.state.anchored_landmarks.items())]
        assert vals.shape[1] == len(ref), (what, vals.shape, len(ref))
        for l, r in enumerate(ref):
            assert np.abs(vals[b, l] - r).max() <= 1e-8 * max(1.0, np.abs(r).max()), (what, b, l, vals[b, l], r)
