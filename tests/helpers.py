"""Shared helpers for the parity tests: build an oracle filter and a CUDA BatchFilter in the same state."""
import numpy as np

import ingvio_oracle as o
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF, StateManager as SM

from ingvio_b200 import synth
from ingvio_b200.synth import WORKLOADS, SyntheticStream


def rand_rot(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def filter_params(wl, **kw):
    fp = o.FilterParams(max_sw_clones=wl.sw, enable_gnss=1 if wl.sats > 0 else 0, cam_nums=2 if wl.stereo else 1)
    fp.T_cl2i_R, fp.T_cl2i_p = synth.R_C2I.copy(), synth.P_C2I.copy()
    fp.T_cl2cr_R, fp.T_cl2cr_p = synth.R_CL2CR.copy(), synth.P_CL2CR.copy()
    for k, v in kw.items():
        setattr(fp, k, v)
    return fp


GNSS_INIT = ((GPS, 1.0, 4.0), (GLO, -2.0, 4.0), (GAL, 0.5, 4.0), (BDS, 3.0, 4.0), (FS, 0.1, 1.0), (YOF, 0.3, 0.015 ** 2))


def make_oracles(wl, stream, fp, with_gnss=True):
    """One OracleFilter per sequence, initialised like State::initStateAndCov + addGNSSVariable."""
    ini = stream.initial_state()
    fs = []
    for b in range(stream.B):
        f = o.OracleFilter(fp, stereo=wl.stereo, max_valid_ids=wl.feats)
        f.init(0.0, ini["R"][b], ini["p"][b], ini["v"][b], ini["bg"][b], ini["ba"][b])
        if with_gnss and wl.sats > 0:
            for g, val, cov in GNSS_INIT:
                SM.add_gnss_variable(f.state, g, val, cov)
        fs.append(f)
    return fs


def cov_diag21(fp):
    sp = o.StateParams(fp)
    return np.array([sp.init_cov_rot] * 3 + [sp.init_cov_pos] * 3 + [sp.init_cov_vel] * 3 + [sp.init_cov_bg] * 3 +
                    [sp.init_cov_ba] * 3 + [sp.init_cov_ext_rot] * 3 + [sp.init_cov_ext_pos] * 3) ** 2.0


def make_gpu(wl, stream, fp, with_gnss=True, max_feats=None, max_clones=None, max_sats=None, device=0, max_landmarks=0):
    from ingvio_b200.filter import BatchFilter
    sp = o.StateParams(fp)
    B = stream.B
    g = BatchFilter(B, max_clones or wl.sw, max_feats or max(wl.feats, 1), max_sats or max(wl.sats, 1), stereo=wl.stereo, device=device,
                    noise=dict(noise_g=sp.noise_g, noise_a=sp.noise_a, noise_bg=sp.noise_bg, noise_ba=sp.noise_ba,
                               noise_clockbias=sp.noise_clockbias, noise_cb_rw=sp.noise_cb_rw),
                    gravity=(0.0, 0.0, -fp.gravity_norm), T_cl2cr=(fp.T_cl2cr_R, fp.T_cl2cr_p),
                    chi2_max_dof=max(fp.chi2_max_dof, 160), chi2_thres=fp.chi2_thres, max_landmarks=max_landmarks)
    ini = stream.initial_state()
    g.init_state_and_cov(ini["R"].reshape(B, 9), ini["p"], ini["v"], ini["bg"], ini["ba"],
                         np.tile(fp.T_cl2i_R.reshape(1, 9), (B, 1)), np.tile(fp.T_cl2i_p, (B, 1)), cov_diag21(fp))
    if with_gnss and wl.sats > 0:
        for gt, val, cov in GNSS_INIT:
            g.add_gnss_variable(gt, val, cov)
    return g


def oracle_packed_state(f, max_clones):
    """The packed mean of include/ingvio_b200.h for one oracle filter."""
    st = f.state
    x = np.zeros(39 + 12 * max_clones)
    e = st.extended_pose
    x[0:9] = e.rot.reshape(9)
    x[9:12], x[12:15] = e.vec1, e.vec2
    x[15:18], x[18:21] = st.bg.value(), st.ba.value()
    x[21:30] = st.camleft_imu_extrinsics.rot.reshape(9)
    x[30:33] = st.camleft_imu_extrinsics.vec
    for gt in range(6):
        if gt in st.gnss:
            x[33 + gt] = st.gnss[gt].value()
    for s, t in enumerate(st.sw_times()):
        c = st.sw_camleft_poses[t]
        x[39 + 12 * s:39 + 12 * s + 9] = c.rot.reshape(9)
        x[39 + 12 * s + 9:39 + 12 * s + 12] = c.vec
    return x


def block_rel_error(P, Po, blocks):
    """Largest error of any variable block pair (i, j), relative to the natural scale of that block:
    |dP_ij|_F / sqrt(|Po_ii|_F |Po_jj|_F).  The global Frobenius test is dominated by the O(1..10) clock states; this one
    holds the 1e-4..1e-6 rotation / extrinsic blocks to the same number of digits."""
    worst, where = 0.0, None
    dn = [np.linalg.norm(Po[i:i + s, i:i + s]) for i, s in blocks]
    for a, (i, si) in enumerate(blocks):
        for c, (j, sj) in enumerate(blocks):
            scale = np.sqrt(dn[a] * dn[c])
            if scale <= 0.0:
                continue
            e = np.linalg.norm(P[i:i + si, j:j + sj] - Po[i:i + si, j:j + sj]) / scale
            if e > worst:
                worst, where = e, (i, j)
    return worst, where


def oracle_blocks(f):
    """(idx, size) of every variable of an oracle state, in covariance order (State::_err_variables)."""
    return [(v.idx(), v.size()) for v in f.state.err_variables]


def assert_state_close(g, oracles, max_clones, tol_P=1e-8, tol_x=1e-9, what="", tol_block=None):
    P = g.get_full_cov()
    X = g.get_state()
    tol_block = 10.0 * tol_P if tol_block is None else tol_block
    for b, f in enumerate(oracles):
        Po = f.cov()
        assert P[b].shape == Po.shape, (what, P[b].shape, Po.shape)
        err = np.linalg.norm(P[b] - Po) / max(1.0, np.linalg.norm(Po))
        assert err <= tol_P, f"{what}: seq {b}: |dP|_F/max(1,|P|_F) = {err:.3e}"
        berr, where = block_rel_error(P[b], Po, oracle_blocks(f))
        assert berr <= tol_block, f"{what}: seq {b}: block {where}: |dP_ij|_F/sqrt(|P_ii||P_jj|) = {berr:.3e}"
        xo = oracle_packed_state(f, max_clones)
        n_used = 39 + 12 * len(f.state.sw_camleft_poses)
        ex = np.max(np.abs(X[b, :n_used] - xo[:n_used]) / np.maximum(1.0, np.abs(xo[:n_used])))
        assert ex <= tol_x, f"{what}: seq {b}: state mismatch {ex:.3e}"


def gstep(g, fr, fp, want=False):
    """BatchFilter.step with the updater options of `fp` (IngvioParams)."""
    return g.step(fr, noise=fp.visual_noise, psr_amp=fp.psr_noise_amp, dopp_amp=fp.dopp_noise_amp,
                  is_adjust_yof=fp.is_adjust_yof, gnss_chi2_test=fp.gnss_chi2_test,
                  gnss_strong_reject=fp.gnss_strong_reject, want=want)


class TiledStream:
    """A SyntheticStream of a few distinct sequences presented as a batch of B (sequence b = b % distinct): lets a
    chip-filling batch (the dispatcher's B >= 296 paths) be checked against a handful of oracle filters."""

    def __init__(self, stream, B):
        self.inner, self.B = stream, B

    def initial_state(self):
        ini = self.inner.initial_state()
        reps = -(-self.B // self.inner.B)
        return {k: (np.concatenate([v] * reps, axis=0)[:self.B] if isinstance(v, np.ndarray) and v.ndim else v)
                for k, v in ini.items()}

    def next_frame(self, **kw):
        return self.inner.next_frame(**kw).tiled(self.B)
