"""GPU: the device track table (igv_tracks_*, ingvio_b200/csrc/k_tracks.cu) through the C-ABI.

  * the random-message scenario of tests/track_scenario.py against the oracle MapServer, bit for bit (table contents
    after every call, every gathered selection), mono / stereo, SwMarg / keyframe window policies;
  * the reference's own MapServer test (TestMapServer.cpp:184-308) and the hand-made depth branches;
  * whole frames: tracker messages -> DeviceMapServer (DEVICE pointer mode chain: gather -> igv_triangulate ->
    igv_msckf_update, nothing returns to the host) against the oracle filter driven by the oracle MapServer, at the
    parity bar of the other GPU tests (|dP|_F <= 1e-8 max(1,|P|_F), state <= 1e-9 relative).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ingvio_oracle as o

from helpers import assert_state_close, cov_diag21, filter_params, make_gpu, make_oracles
from ingvio_b200 import capi
from ingvio_b200.filter import BatchFilter
from ingvio_b200.synth import SyntheticStream, TrackerStream, Workload
from test_oracle_map_server import run_reference_test_on_table
from test_tracks_emulated import depth_branches
from track_frames import OracleFrontEnd
from track_scenario import check_tables, run_scenario


def _bare_filter(B, max_clones, max_feats, max_tracks, stereo):
    g = BatchFilter(B, max_clones, max_feats, 1, stereo=stereo)
    eye = np.tile(np.eye(3).reshape(1, 9), (B, 1))
    z = np.zeros((B, 3))
    g.init_state_and_cov(eye, z, z, z, z, eye, z, np.full(21, 1e-2))
    g.create_map_server(max_tracks)
    return g


def _window_ops(g):
    B = g.B

    def augment(R, p):
        g.augment_sliding_window_pose_cov(np.tile(np.eye(3).reshape(1, 9), (B, 1)), np.asarray(R).reshape(B, 9),
                                          np.asarray(p).reshape(B, 3))

    def marg(slot):
        g.marg_sliding_window_pose(slot)

    def clone_poses(b):
        X = g.get_state()
        return [(X[b, 39 + 12 * s:39 + 12 * s + 9].reshape(3, 3).copy(), X[b, 39 + 12 * s + 9:39 + 12 * s + 12].copy())
                for s in range(g.num_clones())]

    return augment, marg, clone_poses


@pytest.mark.parametrize("mode,stereo,SW,seed", [("sw_marg", False, 5, 11), ("keyframe", False, 6, 12),
                                                 ("sw_marg", True, 4, 13), ("keyframe", True, 5, 14)])
def test_track_table_follows_the_oracle_map_server(mode, stereo, SW, seed):
    B, F, T = 4, 24, 40
    cap = SW + 1 if mode == "sw_marg" else SW
    g = _bare_filter(B, cap, F, T, stereo)
    augment, marg, clone_poses = _window_ops(g)
    cov = run_scenario(g, augment, marg, clone_poses, mode, B, SW, stereo, frames=16, seed=seed, F=F)
    assert cov["lost"] > 20 and 0 < cov["lost_ok"] < cov["lost"] and cov["seen"] > 20 and cov["seen_ok"] > 0
    assert cov["reanchored"] > 0 and cov["dup_frames"] > 0
    g.close()


class _Tiled:
    """Drives a filter of B = reps * b0 sequences with the inputs of b0 sequences tiled `reps` times, checks that every
    copy returns the same bits and hands the first b0 back: the scenario's oracle only has to follow b0 sequences."""

    def __init__(self, g, b0):
        self.g, self.b0, self.reps = g, b0, g.B // b0
        self.B, self.rho, self.max_clones, self.max_feats = b0, g.rho, g.max_clones, g.max_feats

    def _tile(self, x):
        x = np.asarray(x)
        return np.ascontiguousarray(np.tile(x, (self.reps,) + (1,) * (x.ndim - 1)))

    def _first(self, d, what):
        out = {}
        for k, v in d.items():
            v = v.reshape((self.reps, self.b0) + v.shape[1:])
            assert all(np.array_equal(v[0], v[r]) for r in range(1, self.reps)), (what, k)
            out[k] = v[0].copy()
        return out

    def collect_meas(self, n, ids, uv):
        self.g.collect_meas(self._tile(n), self._tile(ids), self._tile(uv))

    def mark_marg_features(self):
        self.g.mark_marg_features()

    def gather_tracks(self, rule, **kw):
        out = self._first(self.g.gather_tracks(rule, **kw), "gather")
        # entry numbers are per-sequence table positions: identical across copies by determinism
        return out

    def commit_triangulation(self, entry, pf, ok, feat_ok=None):
        fo = self._tile(feat_ok) if feat_ok is not None else None
        self.g.commit_triangulation(self._tile(entry), self._tile(pf), self._tile(ok), fo)
        if feat_ok is not None:
            feat_ok[...] = self._first(dict(f=fo), "commit")["f"]
        return feat_ok

    def erase_tracks(self, entry):
        self.g.erase_tracks(self._tile(entry))

    def clean_obs_at(self, slots):
        self.g.clean_obs_at(slots)

    def change_msckf_anchor(self, slots, thr):
        self.g.change_msckf_anchor(slots, thr)

    def erase_invalid_features(self, thr=0.2):
        self.g.erase_invalid_features(thr)

    def get_map_server(self, obs_slots=None, with_obs=True):
        return self._first(self.g.get_map_server(obs_slots, with_obs), "dump")

    def flags(self, clear=True):
        return self._first(dict(f=self.g.flags(clear)), "flags")["f"]


def test_track_table_large_batch_and_capacity():
    """More sequences than SMs (160 = 20 copies of 8), a table and a message stride that need the opt-in shared-memory
    size of k_trk_collect (55 KB)."""
    b0, reps, F, T, SW = 8, 20, 16, 2000, 4
    g = _bare_filter(b0 * reps, SW + 1, F, T, False)
    aug, mrg, poses = _window_ops(g)
    tb = _Tiled(g, b0)
    cov = run_scenario(tb, lambda R, p: aug(tb._tile(R), tb._tile(p)), mrg, poses, "sw_marg", b0, SW, False, frames=5,
                       seed=21, F=F, meas_target=6, meas_stride=4096)
    assert cov["lost"] > 0 and cov["seen"] > 0
    g.close()


@pytest.mark.parametrize("stereo", [False, True])
def test_reference_map_server_test_on_device(stereo):
    g = _bare_filter(1, 4, 8, 16, stereo)
    augment, _, _ = _window_ops(g)
    run_reference_test_on_table(g, augment, stereo)
    g.close()


def test_depth_branches_on_device():
    g = _bare_filter(2, 3, 8, 8, False)
    augment, _, _ = _window_ops(g)
    depth_branches(g, augment)
    g.close()


def test_errors():
    g = BatchFilter(1, 3, 4, 1)
    with pytest.raises(capi.IgvError) as e:
        g.mark_marg_features()
    assert e.value.status == capi.IGV_ERR_STATE                       # table not created
    eye = np.eye(3).reshape(1, 9)
    z = np.zeros((1, 3))
    g.init_state_and_cov(eye, z, z, z, z, eye, z, np.full(21, 1e-2))
    g.create_map_server(8)
    with pytest.raises(capi.IgvError) as e:
        g.collect_meas(np.array([1], np.int32), np.array([[5]], np.uint64), np.zeros((1, 1, 2)))
    assert e.value.status == capi.IGV_ERR_STATE                       # "Meas timestamp not in sw!"
    g.augment_sliding_window_pose()
    with pytest.raises(capi.IgvError) as e:
        g.clean_obs_at([2])
    assert e.value.status == capi.IGV_ERR_STATE                       # slot not in the window
    with pytest.raises(capi.IgvError) as e:
        g.create_map_server(8)
    assert e.value.status == capi.IGV_ERR_STATE
    g.close()


@pytest.mark.parametrize("keyframe,stereo", [(False, False), (True, False), (False, True)])
def test_frames_from_tracker_messages(keyframe, stereo):
    """Tracker messages in, filter state out: DeviceMapServer (device-pointer chain) vs the oracle front end."""
    frames_from_tracker_messages(keyframe, stereo, "cuda")


def frames_from_tracker_messages(keyframe, stereo, device, n_frames=14, B=3):
    """device "cpu": the same sequence with host scratch tensors, i.e. HOST pointer mode on every call
    (tests/test_capi_on_cpu_model.py runs that on the CPU model of the library)."""
    import torch
    from ingvio_b200.map_server import DeviceMapServer
    SW, F = 5, 32
    wl = Workload("trk", 11 + int(stereo), SW + (0 if keyframe else 1), F, 0, stereo=stereo)
    fp = filter_params(wl, max_sw_clones=SW, frame_select_interval=2)
    st = SyntheticStream(wl, B)
    trk = TrackerStream(st, 18, 32, id_base=(1 << 33))
    fes = [OracleFrontEnd(f, keyframe) for f in make_oracles(wl, st, fp, with_gnss=False)]
    g = make_gpu(wl, st, fp, with_gnss=False)
    dms = DeviceMapServer(g, 96, device=device)
    used = 0
    for k in range(n_frames):
        st.n_clones = 0
        fr = st.next_frame(with_visual=False, with_gnss=False, marg_oldest=False)
        n, ids, uv = trk.message(fr.t)
        infos = [fe.frame(fr.seq(b), int(n[b]), ids[b], uv[b]) for b, fe in enumerate(fes)]
        info = infos[0]
        g.propagate_imu(fr.gyro, fr.accel, fr.dt)
        g.augment_sliding_window_pose()
        dms.collect(n, ids, uv)
        dms.remove_lost_update(fp.visual_noise, max_valid=20)
        used += int(dms.selected_counts().sum())
        if info["marg_slots"]:
            dms.selected_update(info["sel_slots"], fp.visual_noise, dof_fixed=info["dof_fixed"])
            used += int(dms.selected_counts().sum())
            dms.slide(info["marg_slots"], info["thr"])
        dms.erase_invalid(0.2)
        assert_state_close(g, [fe.f for fe in fes], wl.sw, what=f"frame {k}")

        class _Orc:   # adaptor for check_tables
            pass
        orc = _Orc()
        orc.B, orc.stereo = B, stereo
        orc.maps = [fe.ms for fe in fes]
        orc.states = [fe.f.state for fe in fes]
        check_tables(g, orc, wl.sw, f"frame {k} tables", pf_tol=1e-7)
    assert used > 0 and not (g.flags() & ~32).any()   # 32 = IGV_FLAG_WEAK_PIVOT, informational
    if device == "cuda":
        torch.cuda.synchronize()
    g.close()
