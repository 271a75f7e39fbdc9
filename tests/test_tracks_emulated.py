"""Track-table kernels (ingvio_b200/csrc/k_tracks.cu) executed on the CPU through tests/emul (threads + barriers model
of a CTA) against the oracle MapServer, bit for bit, in the reference's per-frame call order. The same scenario runs on
the GPU through the C-ABI in tests/test_gpu_tracks.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))

from trk_emul import EmulatedTrackTable  # noqa: E402
from track_scenario import run_scenario  # noqa: E402


@pytest.mark.parametrize("mode,stereo,SW,seed", [("sw_marg", False, 5, 11), ("keyframe", False, 6, 12),
                                                 ("sw_marg", True, 4, 13), ("keyframe", True, 5, 14)])
def test_emulated_kernels_follow_the_oracle_map_server(mode, stereo, SW, seed):
    B, F, T = 4, 24, 40
    cap = SW + 1 if mode == "sw_marg" else SW
    tab = EmulatedTrackTable(B, cap, F, T, stereo)
    try:
        cov = run_scenario(tab, tab.augment, tab.marg, tab.clone_poses, mode, B, SW, stereo, frames=22, seed=seed, F=F)
    finally:
        tab.close()
    # the scenario must actually exercise the branches
    assert cov["lost"] > 20 and cov["lost_ok"] > 0 and cov["lost_ok"] < cov["lost"]
    assert cov["seen"] > 20 and cov["seen_ok"] > 0
    assert cov["reanchored"] > 0 and cov["dup_frames"] > 0
    assert cov["max_tracks"] <= T


def test_table_full_and_gather_cut_are_flagged():
    B, F, T, SW = 2, 4, 6, 3
    tab = EmulatedTrackTable(B, SW, F, T, False)
    try:
        R = np.tile(np.eye(3), (B, 1, 1))
        tab.augment(R, np.zeros((B, 3)))
        ids = np.tile(np.arange(100, 109, dtype=np.uint64), (B, 1))
        uv = np.random.default_rng(0).standard_normal((B, 9, 2))
        n = np.array([9, 5], np.int32)
        tab.collect_meas(n, ids, uv)
        fl = tab.flags()
        assert fl[0] == 8 and fl[1] == 0                       # IGV_FLAG_TRACKS_FULL only where 9 > 6
        d = tab.get_map_server()
        assert d["n_tracks"].tolist() == [6, 5]
        assert sorted(d["id"][0][d["used"][0] == 1].tolist()) == [100, 101, 102, 103, 104, 105]   # message order wins
        tab.augment(R, np.ones((B, 3)))
        tab.collect_meas(np.zeros(B, np.int32), ids, uv)       # nobody is seen at the new clone
        tab.mark_marg_features()
        g = tab.gather_tracks(0, n_feats=F, obs_slots=SW)
        assert g["n_sel"].tolist() == [4, 4]
        assert tab.flags().tolist() == [16, 16]                # IGV_FLAG_GATHER_CUT: 6 resp. 5 lost tracks > F = 4
        assert g["track_id"][0].tolist() == [100, 101, 102, 103]
    finally:
        tab.close()


def depth_branches(tab, augment):
    """changeMSCKFAnchor (erase when not triangulated / too shallow in the new anchor, SwMargUpdate.cpp:236-255,
    KeyframeUpdate.cpp:304-322) and eraseInvalidFeatures (MapServerManager.cpp:462-479) on hand-made landmarks.
    Shared with tests/test_gpu_tracks.py."""
    B = tab.B
    R = np.tile(np.eye(3), (B, 1, 1))
    augment(R, np.zeros((B, 3)))                         # clone 0 at the origin, looking along +z
    ids = np.tile(np.array([10, 11, 12, 13, 14], np.uint64), (B, 1))
    uv = np.zeros((B, 5, tab.rho))
    n = np.full(B, 5, np.int32)
    tab.collect_meas(n, ids, uv)
    augment(R, np.tile(np.array([0.0, 0.0, 1.0]), (B, 1)))   # clone 1 one metre further along +z
    tab.collect_meas(n, ids, uv)
    g = tab.gather_tracks(1, selected_slots=[0, 1], n_feats=8, obs_slots=tab.max_clones)
    assert g["n_sel"].tolist() == [5] * B and g["track_id"][0, :5].tolist() == [10, 11, 12, 13, 14]
    pf = np.zeros((B, 8, 3))
    ok = np.zeros((B, 8), np.uint8)
    pf[:, 0] = [0.0, 0.0, 5.0]      # id 10: fine in both cameras
    pf[:, 1] = [0.0, 0.0, 1.2]      # id 11: 0.2 m in front of clone 1 -> erased by the keyframe threshold 0.3 only
    pf[:, 2] = [0.0, 0.0, 0.5]      # id 12: behind clone 1 -> erased on re-anchoring
    pf[:, 3] = [0.0, 0.0, 0.1]      # id 13: 0.1 m in front of its anchor (clone 0) -> eraseInvalidFeatures
    ok[:, :4] = 1                   # id 14 is never triangulated
    fo = g["feat_ok"].copy()
    tab.commit_triangulation(g["track_entry"], pf, ok, fo)
    assert fo[0].tolist() == [1, 1, 1, 1, 0, 0, 0, 0]
    d = tab.get_map_server(obs_slots=tab.max_clones)
    by = {int(i): k for k, i in enumerate(d["id"][0]) if d["used"][0][k]}
    assert [int(d["is_tri"][0][by[i]]) for i in (10, 11, 12, 13, 14)] == [1, 1, 1, 1, 0]
    assert np.array_equal(d["pf"][0][by[11]], [0.0, 0.0, 1.2]) and np.array_equal(d["pf_fej"][0][by[11]], [0.0, 0.0, 1.2])
    tab.erase_invalid_features(0.2)
    d = tab.get_map_server(obs_slots=tab.max_clones)
    assert sorted(d["id"][0][d["used"][0] == 1].tolist()) == [10, 11, 12, 14]
    tab.clean_obs_at([0])
    tab.change_msckf_anchor([0], 0.3)
    d = tab.get_map_server(obs_slots=tab.max_clones)
    u = d["used"][0] == 1
    assert sorted(d["id"][0][u].tolist()) == [10] and d["anchor_slot"][0][u].tolist() == [1]
    assert d["n_tracks"].tolist() == [1] * B
    # a second successful triangulation moves the value but keeps the first-estimate (FEJ) position
    g = tab.gather_tracks(1, selected_slots=[1], n_feats=8, obs_slots=tab.max_clones)
    pf[:, 0] = [0.1, 0.0, 5.5]
    tab.commit_triangulation(g["track_entry"], pf, ok, None)
    d = tab.get_map_server(obs_slots=tab.max_clones)
    k = int(np.nonzero(d["used"][0])[0][0])
    assert np.array_equal(d["pf"][0][k], [0.1, 0.0, 5.5]) and np.array_equal(d["pf_fej"][0][k], [0.0, 0.0, 5.0])


def test_depth_branches_emulated():
    tab = EmulatedTrackTable(2, 3, 8, 8, False)
    try:
        depth_branches(tab, tab.augment)
    finally:
        tab.close()


@pytest.mark.parametrize("wname", ["c1", "tiny_stereo"])
def test_emulated_triangulation_kernel_matches_oracle(wname):
    """k_triangulate (ingvio_b200/csrc/k_tri.cu, the kernel igv_triangulate launches) executed on the CPU against the oracle
    Triangulator on the oracle filter's clone poses: same accept / reject decisions, positions to 1e-7 relative (the GPU
    twin is tests/test_gpu_parity.py::test_triangulate_matches_oracle)."""
    import ingvio_oracle as o
    from helpers import filter_params, make_oracles, oracle_packed_state
    from ingvio_b200.synth import WORKLOADS, SyntheticStream
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    for _ in range(wl.sw + 1):
        fr = st.next_frame()
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    fr = st.next_frame()
    for b, f in enumerate(orc):
        f.propagate_augment(fr.seq(b))
    tab = EmulatedTrackTable(2, wl.sw, wl.feats, 8, wl.stereo)
    try:
        tab.X[:] = np.stack([oracle_packed_state(f, wl.sw) for f in orc])
        tab.n_clones = len(orc[0].state.sw_camleft_poses)
        tab.T_cl2cr = (fp.T_cl2cr_R, fp.T_cl2cr_p)
        rng = np.random.default_rng(3)
        mask = fr.obs_mask.copy()
        mask[:, 0, :2] = 0                       # fewer views
        mask[:, 1, :] = 0
        mask[:, 1, :2] = 1                       # <= 4 mono views -> reject (Triangulator.cpp:183)
        obs = fr.obs.copy()
        obs[:, 2] += rng.normal(0, 0.3, obs[:, 2].shape)   # garbage track
        prm = dict(trans_thres=0.1, conv_precision=5e-7, max_depth=60.0)
        pf, ok = tab.triangulate(obs, mask, fr.anchor_slot, **prm)
        tri = o.Triangulator(o.TriParams(**prm))
        n_ok = 0
        for b, f in enumerate(orc):
            times = f.state.sw_times()
            poses = [(f.state.sw_camleft_poses[t].rot, f.state.sw_camleft_poses[t].vec) for t in times]
            for k in range(wl.feats):
                oko, pfo = tri.triangulate_feature(obs[b, k], mask[b, k], poses, int(fr.anchor_slot[b, k]), wl.stereo,
                                                   (fp.T_cl2cr_R, fp.T_cl2cr_p))
                assert bool(ok[b, k]) == bool(oko), (b, k)
                if oko:
                    n_ok += 1
                    assert np.linalg.norm(pf[b, k] - pfo) <= 1e-7 * max(1.0, np.linalg.norm(pfo)), (b, k, pf[b, k], pfo)
        assert n_ok > wl.feats and not ok[:, 1].any()
    finally:
        tab.close()


def test_maximum_sizes_emulated():
    """T = M = 4096 (the limits of igv_tracks_create / igv_tracks_collect): every id of a full message gets its own entry in
    message order, the next message finds all of them again (in another order), a third one with 4096 other ids overflows."""
    T = M = 4096
    tab = EmulatedTrackTable(1, 2, 8, T, False)
    try:
        R = np.eye(3)[None]
        tab.augment(R, np.zeros((1, 3)))
        rng = np.random.default_rng(1)
        ids = (rng.permutation(M).astype(np.uint64) * np.uint64(7) + np.uint64(1 << 33))[None]      # distinct after narrowing
        uv = rng.standard_normal((1, M, 2))
        tab.collect_meas(np.array([M], np.int32), ids, uv)
        d = tab.get_map_server(obs_slots=2)
        assert d["n_tracks"][0] == T and tab.flags()[0] == 0
        narrowed = (ids[0] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        narrowed = np.where(narrowed >= 2 ** 31, narrowed - 2 ** 32, narrowed).astype(np.int32)
        assert np.array_equal(d["id"][0], narrowed)                        # entry k <- k-th measurement of the message
        assert np.array_equal(d["obs"][0][:, 0, :], uv[0])
        tab.augment(R, np.ones((1, 3)))
        perm = rng.permutation(M)
        uv2 = rng.standard_normal((1, M, 2))
        tab.collect_meas(np.array([M], np.int32), ids[:, perm], uv2)
        d = tab.get_map_server(obs_slots=2)
        assert d["n_tracks"][0] == T and np.all(d["slot_mask"][0] == 3) and tab.flags()[0] == 0
        inv = np.empty(M, np.int64)
        inv[perm] = np.arange(M)
        assert np.array_equal(d["obs"][0][:, 1, :], uv2[0][inv])
        other = (np.arange(M, dtype=np.uint64) + np.uint64(10 ** 9))[None]
        tab.collect_meas(np.array([M], np.int32), other, uv)
        assert tab.flags()[0] == 8 and tab.get_map_server(obs_slots=2)["n_tracks"][0] == T      # IGV_FLAG_TRACKS_FULL
    finally:
        tab.close()
