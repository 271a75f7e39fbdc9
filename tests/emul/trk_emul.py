"""TEST INFRASTRUCTURE: builds tests/emul/trk_emul.cpp (the track-table kernels of ingvio_b200/csrc/k_tracks.cu compiled
for the CPU through cuda_emul.h) and wraps it with the method names of BatchFilter's track-table calls, so that
tests/track_scenario.py drives the emulated kernels (CPU) and the real C-ABI (GPU) through one code path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "trk_emul.cpp")
OUT = os.path.join(HERE, "_build", "libtrk_emul.so")
DEPS = [SRC, os.path.join(HERE, "cuda_emul.h"), os.path.join(ROOT, "ingvio_b200", "csrc", "k_tracks.cu"),
        os.path.join(ROOT, "ingvio_b200", "csrc", "igv_internal.h"), os.path.join(ROOT, "include", "ingvio_b200.h")]


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I" + cuda_inc, "-I" + HERE, SRC, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("trk_emul build failed:\n" + r.stdout + r.stderr)
    return OUT


_vp, _i, _d = C.c_void_p, C.c_int, C.c_double


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class EmulatedTrackTable:
    """Same track-table method names as ingvio_b200.filter.BatchFilter, backed by the CPU-emulated kernels."""

    def __init__(self, batch, max_clones, max_feats, max_tracks, stereo=False):
        self.lib = C.CDLL(build())
        L = self.lib
        L.emu_create.restype = _vp
        L.emu_create.argtypes = [_i] * 5
        L.emu_flags.restype = C.POINTER(C.c_int)
        L.emu_flags.argtypes = [_vp]
        L.emu_destroy.argtypes = [_vp]
        L.emu_set_X.argtypes = [_vp, _vp]
        L.emu_on_augment.argtypes = [_vp]
        L.emu_on_marg.argtypes = [_vp, _i]
        L.emu_collect.argtypes = [_vp, _vp, _i, _vp, _vp]
        L.emu_mark_lost.argtypes = [_vp]
        L.emu_gather.argtypes = [_vp, _i, _i, _vp, _i, _i, _i, _i] + [_vp] * 9
        L.emu_commit_tri.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
        L.emu_erase.argtypes = [_vp, _i, _vp]
        L.emu_clean.argtypes = [_vp, _i, _vp]
        L.emu_change_anchor.argtypes = [_vp, _i, _vp, _d]
        L.emu_erase_invalid.argtypes = [_vp, _d]
        L.emu_dump.argtypes = [_vp, _i] + [_vp] * 10
        L.emu_triangulate.argtypes = [_vp, _i, _i, _i] + [_vp] * 8
        self.B, self.max_clones, self.max_feats, self.max_tracks = batch, max_clones, max_feats, max_tracks
        self.stereo = bool(stereo)
        self.rho = 4 if stereo else 2
        self.xsize = 39 + 12 * max_clones
        self.h = L.emu_create(batch, max_tracks, max_clones, self.rho, self.xsize)
        self.X = np.zeros((batch, self.xsize))
        self.n_clones = 0
        self.T_cl2cr = (np.eye(3), np.zeros(3))

    def close(self):
        if self.h:
            self.lib.emu_destroy(self.h)
            self.h = None

    # window bookkeeping (the real handle does this inside igv_augment_clone* / igv_marginalize_clone)
    def augment(self, clone_R, clone_p):
        assert self.n_clones < self.max_clones
        s = self.n_clones
        self.X[:, 39 + 12 * s:39 + 12 * s + 9] = np.asarray(clone_R).reshape(self.B, 9)
        self.X[:, 39 + 12 * s + 9:39 + 12 * s + 12] = np.asarray(clone_p).reshape(self.B, 3)
        self.n_clones += 1
        assert self.lib.emu_on_augment(self.h) >= 0

    def marg(self, slot):
        self.lib.emu_on_marg(self.h, int(slot))
        blk = self.X[:, 39:].reshape(self.B, self.max_clones, 12)
        blk[:, slot:self.n_clones - 1] = blk[:, slot + 1:self.n_clones].copy()
        blk[:, self.n_clones - 1] = 0.0
        self.n_clones -= 1

    def num_clones(self):
        return self.n_clones

    def clone_poses(self, b):
        return [(self.X[b, 39 + 12 * s:39 + 12 * s + 9].reshape(3, 3).copy(),
                 self.X[b, 39 + 12 * s + 9:39 + 12 * s + 12].copy()) for s in range(self.n_clones)]

    def _sync_X(self):
        self.lib.emu_set_X(self.h, _p(np.ascontiguousarray(self.X)))

    def collect_meas(self, n_meas, ids, uv):
        n_meas = np.ascontiguousarray(n_meas, np.int32)
        ids = np.ascontiguousarray(ids, np.uint64)
        uv = np.ascontiguousarray(uv, np.float64)
        if self.n_clones == 0:
            raise RuntimeError("[FeatureInfoManager]: Meas timestamp not in sw!")
        self.lib.emu_collect(self.h, _p(n_meas), int(ids.shape[1]), _p(ids), _p(uv))

    def mark_marg_features(self):
        self.lib.emu_mark_lost(self.h)

    def gather_tracks(self, rule, selected_slots=(), min_obs=None, dof_fixed=0, n_feats=None, obs_slots=None, out=None):
        F = int(n_feats or self.max_feats)
        SW = int(obs_slots or self.max_clones)
        B = self.B
        out = dict(track_entry=np.zeros((B, F), np.int32), n_sel=np.zeros(B, np.int32),
                   track_id=np.zeros((B, F), np.int32), obs=np.zeros((B, F, SW, self.rho)),
                   mask_all=np.zeros((B, F, SW), np.uint8), mask_upd=np.zeros((B, F, SW), np.uint8),
                   anchor_slot=np.zeros((B, F), np.int32), chi2_dof=np.zeros((B, F), np.int32),
                   feat_ok=np.zeros((B, F), np.uint8))
        sel = np.ascontiguousarray(list(selected_slots), dtype=np.int32)
        rc = self.lib.emu_gather(self.h, int(rule), len(sel), _p(sel) if len(sel) else None,
                                 int(min_obs if min_obs is not None else (3 if self.stereo else 4)), int(dof_fixed), F, SW,
                                 _p(out["track_entry"]), _p(out["n_sel"]), _p(out["track_id"]), _p(out["obs"]),
                                 _p(out["mask_all"]), _p(out["mask_upd"]), _p(out["anchor_slot"]), _p(out["chi2_dof"]),
                                 _p(out["feat_ok"]))
        assert rc == 0
        return out

    def commit_triangulation(self, track_entry, pf, ok, feat_ok=None):
        e = np.ascontiguousarray(track_entry, np.int32)
        pf = np.ascontiguousarray(pf, np.float64)
        ok = np.ascontiguousarray(ok, np.uint8)
        self.lib.emu_commit_tri(self.h, int(e.shape[1]), _p(e), _p(pf), _p(ok), _p(feat_ok))
        return feat_ok

    def erase_tracks(self, track_entry):
        e = np.ascontiguousarray(track_entry, np.int32)
        self.lib.emu_erase(self.h, int(e.shape[1]), _p(e))

    def clean_obs_at(self, slots):
        s = np.ascontiguousarray(list(slots), np.int32)
        assert self.lib.emu_clean(self.h, len(s), _p(s)) == 0

    def change_msckf_anchor(self, old_slots, min_depth):
        self._sync_X()
        s = np.ascontiguousarray(list(old_slots), np.int32)
        assert self.lib.emu_change_anchor(self.h, len(s), _p(s), float(min_depth)) == 0

    def erase_invalid_features(self, min_depth=0.2):
        self._sync_X()
        if self.n_clones:
            self.lib.emu_erase_invalid(self.h, float(min_depth))

    def get_map_server(self, obs_slots=None, with_obs=True):
        B, T = self.B, self.max_tracks
        SW = int(obs_slots or self.max_clones)
        out = dict(id=np.zeros((B, T), np.int32), used=np.zeros((B, T), np.uint8), to_marg=np.zeros((B, T), np.uint8),
                   is_tri=np.zeros((B, T), np.uint8), slot_mask=np.zeros((B, T), np.uint64),
                   anchor_slot=np.zeros((B, T), np.int32), pf=np.zeros((B, T, 3)), pf_fej=np.zeros((B, T, 3)),
                   n_tracks=np.zeros(B, np.int32), obs=np.zeros((B, T, SW, self.rho)))
        self.lib.emu_dump(self.h, SW, _p(out["id"]), _p(out["used"]), _p(out["to_marg"]), _p(out["is_tri"]),
                          _p(out["slot_mask"]), _p(out["anchor_slot"]), _p(out["pf"]), _p(out["pf_fej"]), _p(out["obs"]),
                          _p(out["n_tracks"]))
        return out

    def triangulate(self, obs, obs_mask, anchor_slot=None, **prm):
        """k_triangulate (ingvio_b200/csrc/k_tri.cu) on the CPU, same arguments as BatchFilter.triangulate (host arrays)."""
        class P(C.Structure):
            _fields_ = [("trans_thres", _d), ("huber_epsilon", _d), ("conv_precision", _d), ("init_damping", _d),
                        ("outer_loop_max_iter", _i), ("inner_loop_max_iter", _i), ("max_depth", _d), ("min_depth", _d)]
        d = dict(trans_thres=0.1, huber_epsilon=0.01, conv_precision=5e-7, init_damping=1e-3, outer_loop_max_iter=10,
                 inner_loop_max_iter=10, max_depth=60.0, min_depth=0.2)
        d.update(prm)
        p = P(**d)
        obs = np.ascontiguousarray(obs, np.float64)
        mask = np.ascontiguousarray(obs_mask, np.uint8)
        anc = np.ascontiguousarray(anchor_slot, np.int32) if anchor_slot is not None else None
        F, SW = mask.shape[1], mask.shape[2]
        pf = np.zeros((self.B, F, 3))
        ok = np.zeros((self.B, F), np.uint8)
        self._sync_X()
        Rc = np.ascontiguousarray(self.T_cl2cr[0], np.float64).reshape(9)
        pc = np.ascontiguousarray(self.T_cl2cr[1], np.float64).reshape(3)
        self.lib.emu_triangulate(self.h, self.n_clones, F, SW, _p(obs), _p(mask), _p(anc), C.addressof(p), _p(Rc), _p(pc),
                                 _p(pf), _p(ok))
        return pf, ok.astype(bool)

    def flags(self, clear=True):
        f = np.array([self.lib.emu_flags(self.h)[b] for b in range(self.B)], dtype=np.int32)
        if clear:
            for b in range(self.B):
                self.lib.emu_flags(self.h)[b] = 0
        return f
