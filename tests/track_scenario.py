"""Shared scenario for the track-table parity tests: random frame messages (births, deaths, repeated ids,
shuffled order, ids beyond 32 bits) driven through the oracle MapServer (oracle/ingvio_oracle/map_server.py) and through
a backend with BatchFilter's track-table methods -- the CPU-emulated kernels (tests/emul) or the CUDA library -- in the
call order of IngvioFilter::callbackMonoFrame (IngvioFilter.cpp:143-205), comparing the whole table after every
step and every gathered selection bit for bit."""
import numpy as np

import ingvio_oracle as o
from ingvio_oracle import map_server as oms
from ingvio_oracle.types import SE3

from helpers import rand_rot

TRK_LOST, TRK_SEEN_AT = 0, 1
R_CL2CR = np.eye(3)
P_CL2CR = np.array([0.001, -0.12, 0.003])     # TestTriangulator.cpp:37-38


def _rot_small(rng, s):
    w = rng.standard_normal(3) * s
    return o.gamma_func(w, 0)


class World:
    """Per-sequence camera path, landmarks and the tracker's frame messages."""

    def __init__(self, seed, B, stereo, meas_target=14, meas_stride=24, id_bases=None):
        self.rng = np.random.default_rng(seed)
        self.B, self.stereo, self.rho = B, stereo, 4 if stereo else 2
        self.M = meas_stride
        self.target = meas_target
        self.id_bases = id_bases or [0, (1 << 32) + 5, 0x7FFFFFF0, (1 << 40) + 0xFFFFFFF0]
        self.R = [rand_rot(self.rng) for _ in range(B)]
        self.p = [self.rng.standard_normal(3) for _ in range(B)]
        self.next_id = [1] * B
        self.active = [dict() for _ in range(B)]   # local id -> (landmark, noise sigma)
        self.retired = [[] for _ in range(B)]
        self.k = 0

    def advance_pose(self):
        for b in range(self.B):
            step = np.array([0.45 + 0.1 * np.sin(0.7 * self.k + b), 0.15 * np.cos(0.5 * self.k), 0.12])
            self.p[b] = self.p[b] + self.R[b] @ step
            self.R[b] = self.R[b] @ _rot_small(self.rng, 0.02)
        self.k += 1
        return np.stack(self.R), np.stack(self.p)

    def _project(self, b, lm, sigma):
        pc = self.R[b].T @ (lm - self.p[b])
        z = [pc[0] / pc[2], pc[1] / pc[2]]
        if self.stereo:
            pr = R_CL2CR @ pc + P_CL2CR
            z += [pr[0] / pr[2], pr[1] / pr[2]]
        return np.array(z) + self.rng.standard_normal(self.rho) * sigma

    def message(self):
        """(n_meas (B,), ids (B,M) uint64, uv (B,M,rho)) for the current poses."""
        rng = self.rng
        n_meas = np.zeros(self.B, np.int32)
        ids = np.zeros((self.B, self.M), np.uint64)
        uv = np.zeros((self.B, self.M, self.rho))
        for b in range(self.B):
            act = self.active[b]
            for lid in list(act.keys()):           # deaths
                if rng.random() < 0.22:
                    self.retired[b].append(lid)
                    del act[lid]
            while len(act) < self.target:          # births (sometimes an id that died earlier comes back)
                if self.retired[b] and rng.random() < 0.15:
                    lid = self.retired[b].pop(rng.integers(len(self.retired[b])))
                    if lid in act:
                        continue
                else:
                    lid = self.next_id[b]
                    self.next_id[b] += int(rng.integers(1, 4))
                depth = rng.uniform(4.0, 25.0)
                lm = self.p[b] + self.R[b] @ np.array([rng.uniform(-0.4, 0.4) * depth, rng.uniform(-0.3, 0.3) * depth, depth])
                sigma = 0.002 if rng.random() < 0.8 else 0.2   # some tracks do not triangulate
                act[lid] = (lm, sigma)
            entries = []
            for lid, (lm, sigma) in act.items():
                entries.append((lid, self._project(b, lm, sigma)))
            order = rng.permutation(len(entries))
            entries = [entries[i] for i in order]
            for _ in range(2):                     # repeated ids inside one message: the first one must win
                if len(entries) < self.M and rng.random() < 0.5:
                    lid = entries[rng.integers(len(entries))][0]
                    entries.insert(int(rng.integers(len(entries) + 1)), (lid, rng.standard_normal(self.rho)))
            n = min(len(entries), self.M)
            n_meas[b] = n
            base = self.id_bases[b % len(self.id_bases)]
            for i in range(n):
                ids[b, i] = np.uint64(base + entries[i][0])
                uv[b, i] = entries[i][1]
            # garbage beyond n_meas must be ignored
            ids[b, n:] = np.uint64(base + 1)
            uv[b, n:] = 7.0
        return n_meas, ids, uv


class OracleSide:
    def __init__(self, B, SW, stereo):
        fp = o.FilterParams(max_sw_clones=SW, enable_gnss=0, cam_nums=2 if stereo else 1)
        fp.T_cl2cr_R, fp.T_cl2cr_p = R_CL2CR.copy(), P_CL2CR.copy()
        self.fp, self.B, self.stereo = fp, B, stereo
        self.states = [o.State(fp) for _ in range(B)]
        self.maps = [oms.MapServer() for _ in range(B)]
        self.tri = o.Triangulator(o.TriParams())
        self.kf = [o.KeyframeUpdate(fp) for _ in range(B)]
        self.swm = o.SwMargUpdate(fp)
        self.t = 0.0

    def augment(self, R, p):
        self.t += 0.05
        for b, st in enumerate(self.states):
            c = SE3()
            c.set_value(R[b], p[b])
            st.sw_camleft_poses[self.t] = c
            st.timestamp = self.t

    def marg_times(self, b, times):
        for t in times:
            del self.states[b].sw_camleft_poses[t]


def _expected_gather(ms, st, keys, stereo, SW, rho, rule, sel_ts, dof_fixed, min_obs):
    times = st.sw_times()
    n = len(keys)
    e = dict(ids=np.array(keys, np.int32), obs=np.zeros((n, SW, rho)), mask_all=np.zeros((n, SW), np.uint8),
             mask_upd=np.zeros((n, SW), np.uint8), anchor=np.zeros(n, np.int32), dof=np.zeros(n, np.int32),
             ok=np.zeros(n, np.uint8))
    for i, k in enumerate(keys):
        f = ms[k]
        ob = f.stereo_obs if stereo else f.mono_obs
        for s, t in enumerate(times):
            if t in ob:
                e["mask_all"][i, s] = 1
                e["obs"][i, s] = ob[t]
                if rule == TRK_LOST or t in sel_ts:
                    e["mask_upd"][i, s] = 1
            if f.anchor is st.sw_camleft_poses[t]:
                e["anchor"][i] = s
        if rule == TRK_LOST:
            e["dof"][i] = max(len(ob) - 1, 1)
            e["ok"][i] = 1 if len(ob) >= min_obs else 0
        else:
            e["dof"][i] = dof_fixed if dof_fixed > 0 else max(len(sel_ts) - 1, 1)
            e["ok"][i] = 1
    return e


def _check_gather(g, b, e, what):
    n = len(e["ids"])
    assert g["n_sel"][b] == n, (what, b, g["n_sel"][b], n)
    assert np.array_equal(g["track_id"][b, :n], e["ids"]), (what, b, g["track_id"][b, :n], e["ids"])
    assert np.all(g["track_entry"][b, :n] >= 0) and np.all(g["track_entry"][b, n:] == -1), what
    assert np.array_equal(g["mask_all"][b, :n], e["mask_all"]), what
    assert np.array_equal(g["mask_upd"][b, :n], e["mask_upd"]), what
    assert np.array_equal(g["obs"][b, :n], e["obs"]), what           # bit-exact copies
    assert np.array_equal(g["anchor_slot"][b, :n], e["anchor"]), what
    assert np.array_equal(g["chi2_dof"][b, :n], e["dof"]), what
    assert np.array_equal(g["feat_ok"][b, :n], e["ok"]), what
    assert not g["mask_all"][b, n:].any() and not g["mask_upd"][b, n:].any() and not g["feat_ok"][b, n:].any(), what
    assert not g["obs"][b, n:].any(), what


def compare_tables(d, snaps, SW, what, pf_tol=0.0):
    """Backend dump `d` (igv_tracks_get layout) against per-sequence canonical snapshots (oracle or golden)."""
    for b, snap in enumerate(snaps):
        used = d["used"][b].astype(bool)
        assert d["n_tracks"][b] == used.sum() == len(snap["id"]), (what, b, d["n_tracks"][b], used.sum(), len(snap["id"]))
        order = np.argsort(d["id"][b][used], kind="stable")
        sel = np.nonzero(used)[0][order]
        assert np.array_equal(d["id"][b][sel], snap["id"]), (what, b)
        assert np.array_equal(d["to_marg"][b][sel], snap["to_marg"]), (what, b, "to_marg")
        assert np.array_equal(d["is_tri"][b][sel], snap["is_tri"]), (what, b, "is_tri")
        bits = (d["slot_mask"][b][sel][:, None] >> np.arange(SW, dtype=np.uint64)[None, :]) & np.uint64(1)
        assert np.array_equal(bits.astype(np.uint8), snap["mask"]), (what, b, "mask")
        assert np.array_equal(d["anchor_slot"][b][sel], snap["anchor_slot"]), (what, b, "anchor")
        assert np.array_equal(d["obs"][b][sel], snap["obs"]), (what, b, "obs")
        if pf_tol == 0.0:
            assert np.array_equal(d["pf"][b][sel], snap["pf"]), (what, b, "pf")
        else:
            assert np.allclose(d["pf"][b][sel], snap["pf"], rtol=0, atol=pf_tol), (what, b, "pf")
        # unused entries read back as empty
        assert not d["slot_mask"][b][~used].any() and np.all(d["anchor_slot"][b][~used] == -1), (what, b)


def check_tables(backend, orc, SW, what, pf_tol=0.0, rec=None):
    snaps = [oms.table_snapshot(orc.maps[b], orc.states[b], orc.stereo, SW) for b in range(orc.B)]
    if rec is not None:
        rec.append(("tables", dict(SW=SW, snaps=snaps)))
    compare_tables(backend.get_map_server(obs_slots=SW), snaps, SW, what, pf_tol)


def _triangulate_gathered(orc, backend_poses, g, b, n, stereo):
    """Oracle triangulation of the gathered arrays of sequence b (what igv_triangulate computes on the device)."""
    F = g["obs"].shape[1]
    pf = np.zeros((F, 3))
    ok = np.zeros(F, np.uint8)
    for f in range(n):
        good, p = orc.tri.triangulate_feature(g["obs"][b, f], g["mask_all"][b, f], backend_poses, int(g["anchor_slot"][b, f]),
                                              stereo, (R_CL2CR, P_CL2CR))
        if good and not np.isnan(p).any():
            ok[f] = 1
            pf[f] = p
    return pf, ok


def run_scenario(backend, augment, marg, clone_poses, mode, B, SW, stereo, frames, seed, F, pf_tol=0.0,
                 meas_target=14, meas_stride=24, rec=None):
    """backend: object with BatchFilter's track-table methods. augment(R,p) / marg(slot) / clone_poses(b) are supplied
    by the caller (emulated window or real filter). mode: "sw_marg" | "keyframe". Returns simple coverage counters."""
    # rec: optional list that receives every input and every oracle expectation, in order
    # (tests/golden/make_golden_tracks.py stores it; replay() feeds it to a backend without the oracle)
    R_ = (lambda kind, **kw: rec.append((kind, kw))) if rec is not None else (lambda kind, **kw: None)
    rho = 4 if stereo else 2
    world = World(seed, B, stereo, meas_target, meas_stride)
    orc = OracleSide(B, SW, stereo)
    cap = SW + 1 if mode == "sw_marg" else SW
    cov = dict(lost=0, lost_ok=0, seen=0, seen_ok=0, reanchored=0, erased_invalid=0, dup_frames=0, max_tracks=0)
    min_obs = 3 if stereo else 4
    for k in range(frames):
        R, p = world.advance_pose()
        augment(R, p)
        orc.augment(R, p)
        n_meas, ids, uv = world.message()
        R_("augment", R=R, p=p)
        R_("collect", n_meas=n_meas, ids=ids, uv=uv)
        backend.collect_meas(n_meas, ids, uv)
        for b in range(B):
            oms.collect_meas(orc.maps[b], orc.states[b], ids[b, :n_meas[b]], uv[b, :n_meas[b]], stereo)
            cov["dup_frames"] += int(len(set(ids[b, :n_meas[b]].tolist())) < n_meas[b])
        check_tables(backend, orc, cap, f"frame {k} collect", pf_tol, rec)

        # ---- RemoveLostUpdate::updateState* : track selection, triangulation, erase -----------------------------
        backend.mark_marg_features()
        g = backend.gather_tracks(TRK_LOST, n_feats=F, obs_slots=cap)
        pf = np.zeros((B, F, 3))
        ok = np.zeros((B, F), np.uint8)
        exp_update, exp_g = [], []
        for b in range(B):
            ms, st = orc.maps[b], orc.states[b]
            oms.mark_marg_features(ms, st, stereo)
            keys = [key for key in ms.ids() if ms[key].is_to_marg]
            e = _expected_gather(ms, st, keys, stereo, cap, rho, TRK_LOST, (), 0, min_obs)
            exp_g.append(e)
            _check_gather(g, b, e, f"frame {k} gather lost")
            pf[b], ok[b] = _triangulate_gathered(orc, clone_poses(b), g, b, len(keys), stereo)
            upd = oms.select_lost(ms, orc.tri, st, stereo)
            exp_update.append(upd)
            cov["lost"] += len(keys)
            cov["lost_ok"] += len(upd)
        R_("mark_gather", rule=TRK_LOST, sel_slots=[], dof_fixed=0, F=F, SW=cap, expect=exp_g)
        R_("commit", pf=pf, ok=ok, expect_update=exp_update)
        feat_ok = g["feat_ok"].copy()
        backend.commit_triangulation(g["track_entry"], pf, ok, feat_ok)
        for b in range(B):
            got = [int(g["track_id"][b, f]) for f in range(int(g["n_sel"][b])) if feat_ok[b, f]]
            assert got == exp_update[b], (k, b, got, exp_update[b])
        R_("erase")
        backend.erase_tracks(g["track_entry"])
        for b in range(B):
            for key in exp_update[b]:
                del orc.maps[b][key]
        check_tables(backend, orc, cap, f"frame {k} remove-lost", pf_tol, rec)

        # ---- SwMargUpdate / KeyframeUpdate : selected clones ------------------------------------------------------
        st0 = orc.states[0]
        if mode == "sw_marg":
            mt = st0.next_marg_time()
            marg_ts = [] if mt == float("inf") else [mt]
            sel_ts = orc.swm.select_sw_timestamps(st0.sw_camleft_poses, mt) if marg_ts else []
            dof_fixed, depth_thr = 0, 0.0
        else:
            marg_ts = orc.kf[0].get_marg_kfs(st0)
            for b in range(1, B):
                assert orc.kf[b].get_marg_kfs(orc.states[b]) == marg_ts
            sel_ts = list(marg_ts)
            dof_fixed, depth_thr = 2, 0.3
        if marg_ts:
            times = st0.sw_times()
            sel_slots = [times.index(t) for t in sel_ts]
            marg_slots = [times.index(t) for t in marg_ts]
            g = backend.gather_tracks(TRK_SEEN_AT, selected_slots=sel_slots, dof_fixed=dof_fixed, n_feats=F, obs_slots=cap)
            pf = np.zeros((B, F, 3))
            ok = np.zeros((B, F), np.uint8)
            exp_update, exp_g = [], []
            for b in range(B):
                ms, st = orc.maps[b], orc.states[b]
                obs_of = (lambda f: f.stereo_obs) if stereo else (lambda f: f.mono_obs)
                keys = [key for key in ms.ids() if all(t in obs_of(ms[key]) for t in sel_ts)]
                e = _expected_gather(ms, st, keys, stereo, cap, rho, TRK_SEEN_AT, sel_ts, dof_fixed, min_obs)
                exp_g.append(e)
                _check_gather(g, b, e, f"frame {k} gather seen-at")
                pf[b], ok[b] = _triangulate_gathered(orc, clone_poses(b), g, b, len(keys), stereo)
                upd = oms.select_seen_at(ms, orc.tri, st, sel_ts, stereo)
                exp_update.append(upd)
                cov["seen"] += len(keys)
                cov["seen_ok"] += len(upd)
            R_("gather", rule=TRK_SEEN_AT, sel_slots=sel_slots, dof_fixed=dof_fixed, F=F, SW=cap, expect=exp_g)
            R_("commit", pf=pf, ok=ok, expect_update=exp_update)
            feat_ok = g["feat_ok"].copy()
            backend.commit_triangulation(g["track_entry"], pf, ok, feat_ok)
            for b in range(B):
                got = [int(g["track_id"][b, f]) for f in range(int(g["n_sel"][b])) if feat_ok[b, f]]
                assert got == exp_update[b], (k, b, got, exp_update[b])
            check_tables(backend, orc, cap, f"frame {k} selected update", pf_tol, rec)

            # clean*ObsAtMargTime -> changeMSCKFAnchor -> margSwPose
            R_("clean", slots=marg_slots)
            backend.clean_obs_at(marg_slots)
            for b in range(B):
                oms.clean_obs_at(orc.maps[b], marg_ts, stereo)
            check_tables(backend, orc, cap, f"frame {k} clean", pf_tol, rec)
            R_("anchor", slots=marg_slots, thr=depth_thr)
            backend.change_msckf_anchor(marg_slots, depth_thr)
            for b in range(B):
                before = {key: f.anchor for key, f in orc.maps[b].items()}
                oms.change_msckf_anchor(orc.maps[b], orc.states[b], marg_ts, depth_thr)
                cov["reanchored"] += sum(1 for key, f in orc.maps[b].items() if f.anchor is not before[key])
            check_tables(backend, orc, cap, f"frame {k} change anchor", pf_tol, rec)
            R_("marg", slots=sorted(marg_slots, reverse=True))
            for s in sorted(marg_slots, reverse=True):
                marg(s)
            for b in range(B):
                orc.marg_times(b, marg_ts)
            check_tables(backend, orc, cap, f"frame {k} marg", pf_tol, rec)

        # ---- eraseInvalidFeatures -----------------------------------------------------------------------------------
        R_("erase_invalid", thr=0.2)
        backend.erase_invalid_features(0.2)
        for b in range(B):
            n0 = len(orc.maps[b])
            oms.erase_invalid_features(orc.maps[b], 0.2)
            cov["erased_invalid"] += n0 - len(orc.maps[b])
            cov["max_tracks"] = max(cov["max_tracks"], len(orc.maps[b]))
        check_tables(backend, orc, cap, f"frame {k} erase invalid", pf_tol, rec)
        assert not backend.flags(clear=True).any()
    return cov


def replay(backend, augment, marg, events, pf_tol=0.0):
    """Feeds a recorded scenario (inputs + oracle expectations, tests/golden/tracks_*.pkl.gz) to a backend; no oracle."""
    g = None
    n_checks = 0
    for i, (kind, kw) in enumerate(events):
        what = f"event {i} {kind}"
        if kind == "augment":
            augment(kw["R"], kw["p"])
        elif kind == "collect":
            backend.collect_meas(kw["n_meas"], kw["ids"], kw["uv"])
        elif kind in ("mark_gather", "gather"):
            if kind == "mark_gather":
                backend.mark_marg_features()
            g = backend.gather_tracks(kw["rule"], selected_slots=kw["sel_slots"], dof_fixed=kw["dof_fixed"], n_feats=kw["F"],
                                      obs_slots=kw["SW"])
            for b, e in enumerate(kw["expect"]):
                _check_gather(g, b, e, what)
            n_checks += 1
        elif kind == "commit":
            feat_ok = g["feat_ok"].copy()
            backend.commit_triangulation(g["track_entry"], kw["pf"], kw["ok"], feat_ok)
            for b, exp in enumerate(kw["expect_update"]):
                got = [int(g["track_id"][b, f]) for f in range(int(g["n_sel"][b])) if feat_ok[b, f]]
                assert got == exp, (what, b, got, exp)
        elif kind == "erase":
            backend.erase_tracks(g["track_entry"])
        elif kind == "clean":
            backend.clean_obs_at(kw["slots"])
        elif kind == "anchor":
            backend.change_msckf_anchor(kw["slots"], kw["thr"])
        elif kind == "marg":
            for s in kw["slots"]:
                marg(s)
        elif kind == "erase_invalid":
            backend.erase_invalid_features(kw["thr"])
        elif kind == "tables":
            compare_tables(backend.get_map_server(obs_slots=kw["SW"]), kw["snaps"], kw["SW"], what, pf_tol)
            n_checks += 1
        else:
            raise ValueError(kind)
    assert not backend.flags(clear=True).any()
    return n_checks
