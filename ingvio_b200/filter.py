"""Host-side mirror of the reference's filter interface for a batch of sequences, on top of the C-ABI.

Method names follow the reference (`StateManager::propagateStateCov`, `augmentSlidingWindowPose`,
`marginalize`, `ekfUpdate`, `getFullCov`, `getMarginalCov`, `boxPlus`, `addGNSSVariable`, ...:
/root/reference/ingvio_estimator/src/StateManager.h:38-127) with snake_case spelling; a variable is
addressed by its (idx, size) exactly as `Type::idx()/size()` (VecState.h:32-54).

Array arguments may be numpy arrays (HOST pointer mode: copied inside the call) or torch CUDA tensors
(DEVICE pointer mode: consumed in place). torch is used only for device memory and streams.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import IgvError
from .frames import FramePacket


def chi2_table(max_dof=150, thres=0.95):
    """UpdateBase::setChiSquaredTable (Update.cpp:27-34): boost quantile -> scipy.stats.chi2.ppf."""
    from scipy.stats import chi2
    return np.array([chi2.ppf(thres, d) for d in range(1, max_dof + 1)], dtype=np.float64)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Arg:
    """Keeps the backing array alive and yields a void* for ctypes."""

    def __init__(self, x, dtype):
        self.is_dev = False
        if x is None:
            self.keep, self.ptr = None, None
            return
        if _is_torch(x):
            import torch
            tdt = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8,
                   np.uint64: torch.int64}[dtype]   # message ids: same 64 bits
            if x.dtype != tdt or not x.is_contiguous():
                x = x.to(tdt).contiguous()
            self.keep = x
            self.ptr = C.c_void_p(x.data_ptr())
            self.is_dev = x.is_cuda
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            self.keep = a
            self.ptr = C.c_void_p(a.ctypes.data)


class BatchFilter:
    """B independent invariant-EKF filters on one GPU (B = 1: drop-in for the reference's single State)."""

    def __init__(self, batch, max_clones, max_feats, max_sats, stereo=False, device=0, stream=None,
                 max_dim=None, noise=None, gravity=(0.0, 0.0, -9.8), T_cl2cr=None, chi2_max_dof=150,
                 chi2_thres=0.95, max_landmarks=0):
        self.lib = capi.load()
        self.B = batch
        self.stereo = bool(stereo)
        self.rho = 4 if stereo else 2
        max_dim = max_dim or (21 + 6 + 6 * max_clones + 3 * max_landmarks)
        cfg = capi.igv_config(batch, max_dim, max_clones, max_feats, max_sats, int(self.stereo), device,
                              C.c_void_p(stream) if stream else None, int(max_landmarks))
        self.max_landmarks = int(max_landmarks)
        self.h = C.c_void_p()
        st = self.lib.igv_create(C.byref(cfg), C.byref(self.h))
        if st != capi.IGV_OK:
            raise IgvError(st, "igv_create failed (is a CUDA device visible?)")
        self.max_clones, self.max_feats, self.max_sats = max_clones, max_feats, max_sats
        p = capi.igv_params()
        nz = dict(noise_g=0.004, noise_a=0.08, noise_bg=0.0002, noise_ba=0.008, noise_clockbias=0.2,
                  noise_cb_rw=0.2)
        nz.update(noise or {})
        for k, v in nz.items():
            setattr(p, k, v)
        p.gravity = (C.c_double * 3)(*gravity)
        Rc, pc = (np.eye(3), np.zeros(3)) if T_cl2cr is None else T_cl2cr
        p.T_cl2cr_R = (C.c_double * 9)(*np.asarray(Rc, float).reshape(9))
        p.T_cl2cr_p = (C.c_double * 3)(*np.asarray(pc, float).reshape(3))
        self._ck(self.lib.igv_set_params(self.h, C.byref(p)))
        tab = chi2_table(max(chi2_max_dof, 16), chi2_thres)
        self._ck(self.lib.igv_set_chi2_table(self.h, tab.ctypes.data_as(capi.c_dp), len(tab)))
        self._mode = capi.IGV_PTR_HOST

    # ---- plumbing --------------------------------------------------------------------------------
    def _ck(self, st):
        if st != capi.IGV_OK:
            raise IgvError(st, self.lib.igv_last_error(self.h).decode())

    def _set_mode(self, args):
        dev = [a.is_dev for a in args if a.ptr is not None]
        mode = capi.IGV_PTR_DEVICE if (dev and all(dev)) else capi.IGV_PTR_HOST
        if dev and any(dev) and not all(dev):
            raise ValueError("mixing host and device arrays in one call")
        if mode != self._mode:
            self._ck(self.lib.igv_set_pointer_mode(self.h, mode))
            self._mode = mode
        return mode

    def close(self):
        if self.h:
            self.lib.igv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_compression(self, kind):
        """capi.COMPRESS_AUTO | COMPRESS_HOUSEHOLDER | COMPRESS_GRAM: how msckf_update forms [R | Q^T r]."""
        self._ck(self.lib.igv_set_compression(self.h, int(kind)))

    def set_precision(self, mode):
        """capi.PREC_FP64 (default) | capi.PREC_FP32_STACK: projected per-track blocks stored in single precision |
        capi.PREC_TF32_GRAM: the same, and their Gram matrix formed on the tcgen05 tensor cores (3 x TF32 split)."""
        self._ck(self.lib.igv_set_precision(self.h, int(mode)))

    def last_gram_tensor(self):
        """1 if the last visual update's Gram matrix came from the tcgen05 kernel (k_gram_tc.cuh)."""
        return int(self.lib.igv_last_gram_tensor(self.h))

    def last_visual_path(self):
        """0 Householder QR, 1 Gram of the materialised stack, 2 Gram fused into the per-track kernel."""
        return int(self.lib.igv_last_visual_path(self.h))

    def synchronize(self):
        self._ck(self.lib.igv_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.igv_launch_count(self.h))

    # ---- layout (Type::idx / size) ---------------------------------------------------------------
    def curr_cov_size(self):
        return self.lib.igv_dim(self.h)

    def curr_err_variable_size(self):
        return self.lib.igv_num_variables(self.h)

    def num_clones(self):
        return self.lib.igv_num_clones(self.h)

    def clone_idx(self, slot):
        return self.lib.igv_clone_idx(self.h, slot)

    def gnss_idx(self, gtype):
        return self.lib.igv_gnss_idx(self.h, gtype)

    # ---- State::initStateAndCov ------------------------------------------------------------------
    def init_state_and_cov(self, R_i2w, p, v, bg, ba, R_ext, p_ext, cov_diag21):
        a = [_Arg(x, np.float64) for x in (R_i2w, p, v, bg, ba, R_ext, p_ext)]
        self._set_mode(a)
        d = np.ascontiguousarray(cov_diag21, dtype=np.float64)
        assert d.shape == (21,)
        self._ck(self.lib.igv_state_init(self.h, *[x.ptr for x in a], d.ctypes.data_as(capi.c_dp)))

    def get_state(self):
        n = self.lib.igv_state_size(self.h)
        out = np.empty((self.B, n))
        self._set_mode([_Arg(out, np.float64)])
        self._ck(self.lib.igv_state_get(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def get_state_async(self, out):
        """Enqueue the read-back of the packed mean into `out` (B x state_size; page-locked host tensor/array or a
        device tensor) and return; pair with fence_record / fence_wait."""
        a = _Arg(out, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_state_get_async(self.h, a.ptr))

    def cov_trace_async(self, out):
        a = _Arg(out, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_cov_trace_async(self.h, a.ptr))

    def fence_record(self, fence):
        self._ck(self.lib.igv_fence_record(self.h, int(fence)))

    def fence_wait(self, fence):
        self._ck(self.lib.igv_fence_wait(self.h, int(fence)))

    def state_size(self):
        return int(self.lib.igv_state_size(self.h))

    def set_state(self, x):
        a = _Arg(x, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_state_set(self.h, a.ptr))

    # ---- StateManager statics --------------------------------------------------------------------
    def get_full_cov(self):
        n = self.curr_cov_size()
        out = np.empty((self.B, n, n))
        self._set_mode([_Arg(out, np.float64)])
        self._ck(self.lib.igv_cov_get(self.h, C.c_void_p(out.ctypes.data), n))
        return out  # symmetric: column-major == row-major

    def set_full_cov(self, P):
        P = np.ascontiguousarray(np.swapaxes(np.asarray(P, float), -1, -2))  # to column-major
        n = self.curr_cov_size()
        assert P.shape == (self.B, n, n)
        a = _Arg(P, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_cov_set(self.h, a.ptr, n))

    @staticmethod
    def _blocks(var_order):
        idx = (C.c_int * len(var_order))(*[int(v[0]) for v in var_order])
        size = (C.c_int * len(var_order))(*[int(v[1]) for v in var_order])
        return idx, size, sum(int(v[1]) for v in var_order)

    def get_marginal_cov(self, var_order):
        idx, size, n = self._blocks(var_order)
        out = np.empty((self.B, n, n))
        self._set_mode([_Arg(out, np.float64)])
        self._ck(self.lib.igv_cov_get_blocks(self.h, len(var_order), idx, size, C.c_void_p(out.ctypes.data)))
        return np.swapaxes(out, -1, -2).copy()

    def add_gnss_variable(self, gtype, value, cov):
        v = np.broadcast_to(np.asarray(value, float), (self.B,)).copy()
        a = _Arg(v, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_add_gnss_variable(self.h, gtype, a.ptr, float(cov)))

    def marg_gnss_variable(self, gtype):
        self._ck(self.lib.igv_marg_gnss_variable(self.h, gtype))

    def add_variable_independent(self, size, cov_block):
        c = np.asfortranarray(np.asarray(cov_block, float).reshape(size, size))
        self._ck(self.lib.igv_add_variable_independent(self.h, size, c.ctypes.data_as(capi.c_dp)))

    def marginalize(self, idx):
        self._ck(self.lib.igv_marginalize(self.h, idx))

    def marg_sliding_window_pose(self, slot=0):
        self._ck(self.lib.igv_marginalize_clone(self.h, slot))

    def propagate_state_cov(self, Phi, G, dt):
        """Phi: (B,15,15), G: (B,15,12) in the usual row/col meaning; dt: (B,)."""
        if not _is_torch(Phi):
            Phi = np.ascontiguousarray(np.swapaxes(np.asarray(Phi, float), -1, -2))   # column-major
            G = np.ascontiguousarray(np.swapaxes(np.asarray(G, float), -1, -2))
        a = [_Arg(Phi, np.float64), _Arg(G, np.float64), _Arg(dt, np.float64)]
        self._set_mode(a)
        self._ck(self.lib.igv_propagate_cov(self.h, a[0].ptr, a[1].ptr, a[2].ptr))

    def propagate_imu(self, gyro, accel, dt):
        a = [_Arg(gyro, np.float64), _Arg(accel, np.float64), _Arg(dt, np.float64)]
        self._set_mode(a)
        n_steps = int(a[2].keep.shape[-1])
        self._ck(self.lib.igv_propagate_imu(self.h, n_steps, a[0].ptr, a[1].ptr, a[2].ptr))

    def augment_sliding_window_pose(self):
        self._ck(self.lib.igv_augment_clone(self.h))

    def augment_sliding_window_pose_cov(self, R_i2w, clone_R=None, clone_p=None):
        a = [_Arg(R_i2w, np.float64), _Arg(clone_R, np.float64), _Arg(clone_p, np.float64)]
        self._set_mode(a)
        self._ck(self.lib.igv_augment_clone_cov(self.h, a[0].ptr, a[1].ptr, a[2].ptr))

    def _r_arg(self, R, rows):
        R = np.asarray(R, float)
        if R.ndim == 0 or R.shape == (self.B,):
            return np.broadcast_to(R, (self.B,)).copy(), capi.R_ISO
        if R.shape[-1] == rows and (R.ndim == 1 or R.shape == (self.B, rows)):
            return np.broadcast_to(R, (self.B, rows)).copy(), capi.R_DIAG
        R = np.broadcast_to(R, (self.B, rows, rows))
        return np.ascontiguousarray(np.swapaxes(R, -1, -2)), capi.R_FULL

    def ekf_update(self, var_order, H, res, R, return_dx=True):
        """StateManager::ekfUpdate. H: (B,rows,n) row/col meaning; res: (B,rows); R: scalar sigma^2,
        (rows,) diagonal or (rows,rows) dense (optionally with a leading batch dim)."""
        H = np.asarray(H, float)
        rows, n = H.shape[-2], H.shape[-1]
        Hc = np.ascontiguousarray(np.swapaxes(np.broadcast_to(H, (self.B, rows, n)), -1, -2))
        idx, size, nn = self._blocks(var_order)
        assert nn == n
        Rv, kind = self._r_arg(R, rows)
        dx = np.zeros((self.B, self.curr_cov_size())) if return_dx else None
        a = [_Arg(Hc, np.float64), _Arg(np.broadcast_to(np.asarray(res, float), (self.B, rows)).copy(), np.float64),
             _Arg(Rv, np.float64)]
        self._set_mode(a)
        self._ck(self.lib.igv_ekf_update(self.h, len(var_order), idx, size, rows, a[0].ptr, rows, a[1].ptr, a[2].ptr,
                                         kind, C.c_void_p(dx.ctypes.data) if return_dx else None))
        return dx

    def whiten_residual(self, var_order, H, res, R):
        """UpdateBase::whitenResidual -> gamma (B,)."""
        H = np.asarray(H, float)
        rows, n = H.shape[-2], H.shape[-1]
        Hc = np.ascontiguousarray(np.swapaxes(np.broadcast_to(H, (self.B, rows, n)), -1, -2))
        idx, size, _ = self._blocks(var_order)
        Rv, kind = self._r_arg(R, rows)
        g = np.zeros(self.B)
        a = [_Arg(Hc, np.float64), _Arg(np.broadcast_to(np.asarray(res, float), (self.B, rows)).copy(), np.float64),
             _Arg(Rv, np.float64)]
        self._set_mode(a)
        self._ck(self.lib.igv_chi2_whiten(self.h, len(var_order), idx, size, rows, a[0].ptr, rows, a[1].ptr, a[2].ptr,
                                          kind, C.c_void_p(g.ctypes.data)))
        return g

    def box_plus(self, dx):
        a = _Arg(dx, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_box_plus(self.h, a.ptr))

    def add_variable_delayed(self, gtype, value, var_old_order, H_old, H_new, res, noise_iso, chi2_mult=0.95,
                             do_chi2=True, prior_cov_if_rejected=1.0):
        H_old = np.asarray(H_old, float)
        rows, n = H_old.shape[-2], H_old.shape[-1]
        Ho = np.ascontiguousarray(np.swapaxes(np.broadcast_to(H_old, (self.B, rows, n)), -1, -2))
        Hn = np.broadcast_to(np.asarray(H_new, float).reshape(-1, rows), (self.B, rows)).copy()
        rr = np.broadcast_to(np.asarray(res, float), (self.B, rows)).copy()
        val = np.broadcast_to(np.asarray(value, float), (self.B,)).copy()
        idx, size, nn = self._blocks(var_old_order)
        assert nn == n
        acc = np.zeros(self.B, dtype=np.int32)
        a = [_Arg(Ho, np.float64), _Arg(Hn, np.float64), _Arg(rr, np.float64), _Arg(val, np.float64)]
        self._set_mode(a)
        self._ck(self.lib.igv_add_variable_delayed(self.h, gtype, a[3].ptr, len(var_old_order), idx, size, rows,
                                                   a[0].ptr, a[1].ptr, a[2].ptr, float(noise_iso), float(chi2_mult),
                                                   int(do_chi2), float(prior_cov_if_rejected),
                                                   C.c_void_p(acc.ctypes.data), None))
        return acc.astype(bool)

    def replace_var_linear(self, target, dependence_order, H):
        H = np.asarray(H, float)
        ts, n = H.shape[-2], H.shape[-1]
        Hc = np.ascontiguousarray(np.swapaxes(np.broadcast_to(H, (self.B, ts, n)), -1, -2))
        idx, size, nn = self._blocks(dependence_order)
        assert nn == n and ts == target[1]
        a = _Arg(Hc, np.float64)
        self._set_mode([a])
        self._ck(self.lib.igv_replace_var_linear(self.h, target[0], target[1], len(dependence_order), idx, size, a.ptr))

    # ---- fused updaters ------------------------------------------------------------------------------
    def msckf_update(self, mode, pf_w, anchor_slot, obs, obs_mask, chi2_dof, noise, max_valid=0, want_dx=False,
                     want_gamma=False, want_accepted=False, feat_ok=None):
        """RemoveLostUpdate / KeyframeUpdate / SwMargUpdate ::updateState* after track selection."""
        a = [_Arg(pf_w, np.float64), _Arg(anchor_slot, np.int32), _Arg(obs, np.float64), _Arg(obs_mask, np.uint8),
             _Arg(chi2_dof, np.int32), _Arg(feat_ok, np.uint8)]
        mode_ptr = self._set_mode(a)
        F = int(a[0].keep.shape[1])
        SW = int(a[3].keep.shape[2])
        args = capi.igv_msckf_args()
        args.mode = mode
        args.n_feats = F
        args.pf_w, args.anchor_slot, args.obs, args.obs_mask, args.chi2_dof, args.feat_ok = [x.ptr for x in a]
        args.obs_slots = SW
        args.noise = float(noise)
        args.max_valid = int(max_valid)
        outs = {}
        host = mode_ptr == capi.IGV_PTR_HOST
        if want_dx and host:
            outs["dx"] = np.zeros((self.B, self.curr_cov_size()))
            args.dx_out = C.c_void_p(outs["dx"].ctypes.data)
        if want_gamma and host:
            outs["gamma"] = np.zeros((self.B, F))
            args.gamma_out = C.c_void_p(outs["gamma"].ctypes.data)
        if want_accepted and host:
            outs["accepted"] = np.zeros(self.B, dtype=np.int32)
            args.n_accepted_out = C.c_void_p(outs["accepted"].ctypes.data)
        self._ck(self.lib.igv_msckf_update(self.h, C.byref(args)))
        return outs

    def triangulate(self, obs, obs_mask, anchor_slot=None, pf_out=None, ok_out=None, **prm):
        """Triangulator::triangulate{Mono,Stereo}Obs for every track (+ the anchor-depth check).
        Host arrays in -> (pf (B,F,3), ok (B,F)) host arrays out; with torch CUDA tensors pass pf_out/ok_out."""
        a = [_Arg(obs, np.float64), _Arg(obs_mask, np.uint8), _Arg(anchor_slot, np.int32)]
        dev_out = pf_out is not None
        if dev_out:
            a += [_Arg(pf_out, np.float64), _Arg(ok_out, np.uint8)]
        mode_ptr = self._set_mode(a)
        F, SW = int(a[1].keep.shape[1]), int(a[1].keep.shape[2])
        args = capi.igv_tri_args()
        args.n_feats, args.obs, args.obs_mask, args.obs_slots, args.anchor_slot = F, a[0].ptr, a[1].ptr, SW, a[2].ptr
        d = dict(trans_thres=0.1, huber_epsilon=0.01, conv_precision=5e-7, init_damping=1e-3, outer_loop_max_iter=10,
                 inner_loop_max_iter=10, max_depth=60.0, min_depth=0.2)
        d.update(prm)
        for k, v in d.items():
            setattr(args.prm, k, v)
        if dev_out:
            args.pf_out, args.ok_out = a[3].ptr, a[4].ptr
            self._ck(self.lib.igv_triangulate(self.h, C.byref(args)))
            return pf_out, ok_out
        assert mode_ptr == capi.IGV_PTR_HOST
        pf = np.zeros((self.B, F, 3))
        ok = np.zeros((self.B, F), dtype=np.uint8)
        args.pf_out, args.ok_out = C.c_void_p(pf.ctypes.data), C.c_void_p(ok.ctypes.data)
        self._ck(self.lib.igv_triangulate(self.h, C.byref(args)))
        return pf, ok.astype(bool)

    def gnss_update(self, unit, res_pos, res_vel, sigma_psr, sigma_dopp, sys, R_enu2ecef, is_adjust_yof=0,
                    chi2_test=0, strong_reject=1, want_dx=False):
        """GnssUpdate::updateTrackedSys from the psr_res/dopp_res boundary."""
        a = [_Arg(unit, np.float64), _Arg(res_pos, np.float64), _Arg(res_vel, np.float64),
             _Arg(sigma_psr, np.float64), _Arg(sigma_dopp, np.float64), _Arg(sys, np.int32),
             _Arg(R_enu2ecef, np.float64)]
        mode_ptr = self._set_mode(a)
        args = capi.igv_gnss_args()
        args.n_sats = int(a[1].keep.shape[1])
        (args.unit, args.res_pos, args.res_vel, args.sigma_psr, args.sigma_dopp, args.sys,
         args.R_enu2ecef) = [x.ptr for x in a]
        args.is_adjust_yof, args.chi2_test, args.strong_reject = int(is_adjust_yof), int(chi2_test), int(strong_reject)
        dx = None
        if want_dx and mode_ptr == capi.IGV_PTR_HOST:
            dx = np.zeros((self.B, self.curr_cov_size()))
            args.dx_out = C.c_void_p(dx.ctypes.data)
        self._ck(self.lib.igv_gnss_update(self.h, C.byref(args)))
        return dx

    def gnss_add_new_tracked_sys(self, gtype, value, unit, res_pos, res_vel, sigma_psr, sigma_dopp, sys, R_enu2ecef,
                                 R_ecef2enu=None, is_adjust_yof=0, chi2_mult=0.95, prior_cov_if_rejected=1.0):
        """GnssUpdate::addNewTrackedSys for one system (GnssUpdate.cpp:317-476): rows built on the device, then
        StateManager::addVariableDelayed. Returns accepted (B,) bool (None in device-pointer mode)."""
        if np.isscalar(value):
            value = np.full(self.B, float(value))
        a = [_Arg(value, np.float64), _Arg(unit, np.float64), _Arg(res_pos, np.float64), _Arg(res_vel, np.float64),
             _Arg(sigma_psr, np.float64), _Arg(sigma_dopp, np.float64), _Arg(sys, np.int32), _Arg(R_enu2ecef, np.float64)]
        if R_ecef2enu is not None:
            a.append(_Arg(R_ecef2enu, np.float64))
        mode = self._set_mode(a)
        args = capi.igv_gnss_new_sys_args()
        args.n_sats, args.gtype = int(a[2].keep.shape[1]), int(gtype)
        (args.value, args.unit, args.res_pos, args.res_vel, args.sigma_psr, args.sigma_dopp, args.sys,
         args.R_enu2ecef) = [x.ptr for x in a[:8]]
        args.R_ecef2enu = a[8].ptr if R_ecef2enu is not None else None
        args.is_adjust_yof, args.chi2_mult, args.prior_cov_if_rejected = int(is_adjust_yof), float(chi2_mult), float(prior_cov_if_rejected)
        acc = None
        if mode == capi.IGV_PTR_HOST:
            acc = np.zeros(self.B, dtype=np.int32)
            args.accepted_out = C.c_void_p(acc.ctypes.data)
        self._ck(self.lib.igv_gnss_add_new_tracked_sys(self.h, C.byref(args)))
        return None if acc is None else acc.astype(bool)

    # ---- SLAM landmarks in the state (LandmarkUpdate, mono) -------------------------------------------
    def num_landmarks(self):
        return int(self.lib.igv_num_landmarks(self.h))

    def landmark_idx(self, lm_slot):
        return int(self.lib.igv_landmark_idx(self.h, int(lm_slot)))

    def landmark_anchor(self, lm_slot):
        return int(self.lib.igv_landmark_anchor(self.h, int(lm_slot)))

    def landmark_values(self):
        """World positions (B, L, 3) of the landmarks in the state, in slot order."""
        x = self.get_state()
        off = 39 + 12 * self.max_clones
        L = self.num_landmarks()
        return x[:, off:off + 3 * L].reshape(self.B, L, 3)

    def landmark_init(self, pf_w, anchor_slot, obs, obs_mask, noise, chi2_mult=0.95, prior_cov_if_rejected=1.0):
        """LandmarkUpdate::initNewLandmarkMono for one track (delayed initialisation of a 3-dim anchored variable)."""
        a = [_Arg(pf_w, np.float64), _Arg(obs, np.float64), _Arg(obs_mask, np.uint8)]
        mode = self._set_mode(a)
        args = capi.igv_lm_init_args()
        args.obs_slots, args.anchor_slot = int(a[2].keep.shape[1]), int(anchor_slot)
        args.pf_w, args.obs, args.obs_mask = [x.ptr for x in a]
        args.noise, args.chi2_mult, args.prior_cov_if_rejected = float(noise), float(chi2_mult), float(prior_cov_if_rejected)
        acc = None
        if mode == capi.IGV_PTR_HOST:
            acc = np.zeros(self.B, dtype=np.int32)
            args.accepted_out = C.c_void_p(acc.ctypes.data)
        self._ck(self.lib.igv_landmark_init(self.h, C.byref(args)))
        return None if acc is None else acc.astype(bool)

    def landmark_update(self, uv, valid, noise):
        """LandmarkUpdate::updateLandmarkMono; returns dict(accepted (B,), gamma (B, L)) in host mode."""
        a = [_Arg(uv, np.float64), _Arg(valid, np.uint8)]
        mode = self._set_mode(a)
        args = capi.igv_lm_update_args()
        args.uv, args.valid, args.noise = a[0].ptr, a[1].ptr, float(noise)
        out = None
        if mode == capi.IGV_PTR_HOST:
            out = dict(accepted=np.zeros(self.B, dtype=np.int32), gamma=np.zeros((self.B, max(1, self.num_landmarks()))))
            args.n_accepted_out = C.c_void_p(out["accepted"].ctypes.data)
            args.gamma_out = C.c_void_p(out["gamma"].ctypes.data)
        self._ck(self.lib.igv_landmark_update(self.h, C.byref(args)))
        return out

    def landmark_change_anchor(self, lm_slot, new_clone_slot):
        self._ck(self.lib.igv_landmark_change_anchor(self.h, int(lm_slot), int(new_clone_slot)))

    def landmark_marginalize(self, lm_slot):
        self._ck(self.lib.igv_landmark_marginalize(self.h, int(lm_slot)))

    def sat_states(self, eph, t_obs_rel, psr, sys, out=None):
        """gnss_comm::sat_states: ephemeris records (B,S,24) + observation time relative to toe + L1 pseudo-range ->
        dict(sat_pos, sat_vel, sat_clk, ttx_rel). Host arrays in, numpy out; device tensors need `out`."""
        a = [_Arg(eph, np.float64), _Arg(t_obs_rel, np.float64), _Arg(psr, np.float64), _Arg(sys, np.int32)]
        mode = self._set_mode(a)
        S = int(a[3].keep.shape[1])
        shapes = dict(sat_pos=(self.B, S, 3), sat_vel=(self.B, S, 3), sat_clk=(self.B, S, 3), ttx_rel=(self.B, S))
        if mode == capi.IGV_PTR_HOST:
            out = {k: np.zeros(v) for k, v in shapes.items()}
        else:
            assert out is not None and all(k in out for k in shapes), "device mode needs preallocated outputs"
        o = {k: _Arg(out[k], np.float64) for k in shapes}
        args = capi.igv_sat_state_args()
        args.n_sats = S
        args.eph, args.t_obs_rel, args.psr, args.sys = [x.ptr for x in a]
        for k in shapes:
            setattr(args, k, o[k].ptr)
        self._ck(self.lib.igv_sat_states(self.h, C.byref(args)))
        return out

    def gnss_residuals(self, sat_pos, sat_vel, sat_clk, obs, obs_std, ttx, sys, T_enu2ecef, iono=None, psr_amp=1.0,
                       dopp_amp=1.0, out=None, clock_init=None):
        """gnss_comm::psr_res / dopp_res (+ az/el, delays, sigmas) at the filter's receiver state; returns a dict
        with unit, res_pos, res_vel, sigma_psr, sigma_dopp, azel, atmos -- the inputs of gnss_update. Host arrays in,
        numpy arrays out; device tensors in, `out` must hold preallocated device tensors under the same keys."""
        a = [_Arg(sat_pos, np.float64), _Arg(sat_vel, np.float64), _Arg(sat_clk, np.float64), _Arg(obs, np.float64),
             _Arg(obs_std, np.float64), _Arg(ttx, np.float64), _Arg(sys, np.int32), _Arg(T_enu2ecef, np.float64),
             _Arg(iono, np.float64)]
        ci = _Arg(clock_init, np.float64) if clock_init is not None else None
        mode = self._set_mode(a + ([ci] if ci is not None else []))
        S = int(a[6].keep.shape[1])
        shapes = dict(unit=(self.B, S, 3), res_pos=(self.B, S), res_vel=(self.B, S), sigma_psr=(self.B, S),
                      sigma_dopp=(self.B, S), azel=(self.B, S, 2), atmos=(self.B, S, 2))
        if mode == capi.IGV_PTR_HOST:
            out = {k: np.zeros(v) for k, v in shapes.items()}
        else:
            assert out is not None and all(k in out for k in shapes), "device mode needs preallocated outputs"
        o = {k: _Arg(out[k], np.float64) for k in shapes}
        args = capi.igv_gnss_res_args()
        args.n_sats = S
        (args.sat_pos, args.sat_vel, args.sat_clk, args.obs, args.obs_std, args.ttx, args.sys, args.T_enu2ecef,
         args.iono) = [x.ptr for x in a]
        args.psr_noise_amp, args.dopp_noise_amp = float(psr_amp), float(dopp_amp)
        args.clock_init = ci.ptr if ci is not None else None
        for k in shapes:
            setattr(args, k, o[k].ptr)
        self._ck(self.lib.igv_gnss_residuals(self.h, C.byref(args)))
        return out

    def frame_step(self, gyro, accel, dt, visual=None, marg_slots=(), gnss=None):
        """One whole frame cycle behind ONE C-ABI call (igv_frame_step): propagate + augment, the visual update
        (`visual`: dict with the arguments of msckf_update: mode, pf_w, anchor_slot, obs, obs_mask, chi2_dof, noise,
        max_valid, optional feat_ok / n_accepted_out / gamma_out), marginalisation of `marg_slots`, the GNSS update
        (`gnss`: dict with the arguments of gnss_update). With device tensors at fixed addresses the frame is replayed as a
        CUDA graph from its third occurrence on (graph_replays counts them)."""
        a = [_Arg(gyro, np.float64), _Arg(accel, np.float64), _Arg(dt, np.float64)]
        keep = list(a)
        fa = capi.igv_frame_args()
        fa.n_imu = int(a[2].keep.shape[1]) if a[2].keep is not None else 0
        fa.gyro, fa.accel, fa.dt = a[0].ptr, a[1].ptr, a[2].ptr
        va = ga = None
        if visual is not None:
            v = [_Arg(visual["pf_w"], np.float64), _Arg(visual["anchor_slot"], np.int32), _Arg(visual["obs"], np.float64),
                 _Arg(visual["obs_mask"], np.uint8), _Arg(visual["chi2_dof"], np.int32), _Arg(visual.get("feat_ok"), np.uint8),
                 _Arg(visual.get("n_accepted_out"), np.int32), _Arg(visual.get("gamma_out"), np.float64)]
            keep += v
            va = capi.igv_msckf_args()
            va.mode, va.n_feats, va.obs_slots = int(visual["mode"]), int(v[0].keep.shape[1]), int(v[3].keep.shape[2])
            va.noise, va.max_valid = float(visual["noise"]), int(visual.get("max_valid", 0))
            (va.pf_w, va.anchor_slot, va.obs, va.obs_mask, va.chi2_dof, va.feat_ok, va.n_accepted_out,
             va.gamma_out) = [x.ptr for x in v]
            fa.visual = C.cast(C.pointer(va), C.c_void_p)
        if gnss is not None:
            gk = [_Arg(gnss[k], np.float64) for k in ("unit", "res_pos", "res_vel", "sigma_psr", "sigma_dopp")]
            gk += [_Arg(gnss["sys"], np.int32), _Arg(gnss["R_enu2ecef"], np.float64)]
            keep += gk
            ga = capi.igv_gnss_args()
            ga.n_sats = int(gk[1].keep.shape[1])
            (ga.unit, ga.res_pos, ga.res_vel, ga.sigma_psr, ga.sigma_dopp, ga.sys, ga.R_enu2ecef) = [x.ptr for x in gk]
            ga.is_adjust_yof, ga.chi2_test = int(gnss.get("is_adjust_yof", 0)), int(gnss.get("chi2_test", 0))
            ga.strong_reject = int(gnss.get("strong_reject", 1))
            fa.gnss = C.cast(C.pointer(ga), C.c_void_p)
        self._set_mode([x for x in keep if x.ptr is not None])
        ms = (C.c_int * max(1, len(marg_slots)))(*[int(s) for s in marg_slots])
        fa.n_marg, fa.marg_slots = len(marg_slots), ms
        self._ck(self.lib.igv_frame_step(self.h, C.byref(fa)))

    @property
    def graph_replays(self):
        return int(self.lib.igv_graph_replays(self.h))

    # ---- one frame cycle (IngvioFilter.cpp:143-231 order) ----------------------------------------------
    def step(self, fr: FramePacket, noise=0.12, psr_amp=1.0, dopp_amp=1.0, is_adjust_yof=0, gnss_chi2_test=0,
             gnss_strong_reject=1, want=False):
        out = {}
        self.propagate_imu(fr.gyro, fr.accel, fr.dt)
        self.augment_sliding_window_pose()
        if fr.visual_mode is not None and fr.pf_w.shape[1] > 0:
            ncl = self.num_clones()
            mask = fr.obs_mask
            if fr.visual_mode == "all_obs":
                mode = capi.VIS_ALL_OBS
                dof = fr.obs_total.astype(np.int32) - 1                      # RemoveLostUpdate.cpp:95-96
                max_valid = fr.max_valid
            else:
                mode = capi.VIS_SELECTED
                sel = np.zeros(mask.shape[-1], dtype=np.uint8)
                sel[list(fr.selected_slots)] = 1
                mask = mask * sel
                d = 2 if fr.visual_mode == "keyframe" else len(fr.selected_slots) - 1
                dof = np.full(mask.shape[:2], d, dtype=np.int32)
                max_valid = 0
            out["visual"] = self.msckf_update(mode, fr.pf_w, fr.anchor_slot, fr.obs, mask, dof, noise, max_valid,
                                              want_dx=want, want_gamma=want, want_accepted=want)
            assert ncl <= mask.shape[-1]
        for s in sorted(fr.marg_slots, reverse=True):
            self.marg_sliding_window_pose(s)
        if fr.gnss is not None:
            g = fr.gnss
            out["gnss_dx"] = self.gnss_update(g.unit, g.res_pos, g.res_vel, g.sigma_psr(psr_amp), g.sigma_dopp(dopp_amp),
                                              g.sys, g.R_enu2ecef, is_adjust_yof, gnss_chi2_test, gnss_strong_reject,
                                              want_dx=want)
        return out

    # ---- track table: MapServer / MapServerManager on the device -------------------------
    def create_map_server(self, max_tracks):
        """map_server = std::make_shared<MapServer>() with room for `max_tracks` features per sequence."""
        self._ck(self.lib.igv_tracks_create(self.h, int(max_tracks)))
        self.max_tracks = int(max_tracks)

    def reset_map_server(self):
        self._ck(self.lib.igv_tracks_reset(self.h))

    def collect_meas(self, n_meas, ids, uv):
        """MapServerManager::collectMonoMeas / collectStereoMeas: one frame message per sequence at the newest clone.
        n_meas (B,), ids (B,M) uint64 as in feature_tracker/msg/*Meas.msg, uv (B,M,rho)."""
        a = [_Arg(n_meas, np.int32), _Arg(ids, np.uint64), _Arg(uv, np.float64)]
        self._set_mode(a)
        M = int(a[1].keep.shape[1])
        self._ck(self.lib.igv_tracks_collect(self.h, a[0].ptr, M, a[1].ptr, a[2].ptr))

    def mark_marg_features(self):
        """MapServerManager::markMargMonoFeatures / markMargStereoFeatures."""
        self._ck(self.lib.igv_tracks_mark_lost(self.h))

    def gather_tracks(self, rule, selected_slots=(), min_obs=None, dof_fixed=0, n_feats=None, obs_slots=None, out=None):
        """Track selection of RemoveLostUpdate (rule TRK_LOST) or SwMargUpdate / KeyframeUpdate (TRK_SEEN_AT), emitted
        in ascending id order in the array layout of triangulate() / msckf_update(). `out`: dict of CUDA tensors
        (track_entry, n_sel, track_id, obs, mask_all, mask_upd, anchor_slot, chi2_dof, feat_ok) to stay on the device;
        otherwise host arrays are returned."""
        F = int(n_feats or self.max_feats)
        SW = int(obs_slots or self.max_clones)
        B = self.B
        if out is None:
            out = dict(track_entry=np.zeros((B, F), np.int32), n_sel=np.zeros(B, np.int32),
                       track_id=np.zeros((B, F), np.int32), obs=np.zeros((B, F, SW, self.rho)),
                       mask_all=np.zeros((B, F, SW), np.uint8), mask_upd=np.zeros((B, F, SW), np.uint8),
                       anchor_slot=np.zeros((B, F), np.int32), chi2_dof=np.zeros((B, F), np.int32),
                       feat_ok=np.zeros((B, F), np.uint8))
        dt = dict(track_entry=np.int32, n_sel=np.int32, track_id=np.int32, obs=np.float64, mask_all=np.uint8,
                  mask_upd=np.uint8, anchor_slot=np.int32, chi2_dof=np.int32, feat_ok=np.uint8)
        a = {k: _Arg(out[k], dt[k]) for k in dt}
        for k in dt:   # outputs must be written in place
            if a[k].keep is not out[k]:
                raise ValueError(f"gather_tracks: `{k}` must be a contiguous {dt[k].__name__} array")
        self._set_mode(list(a.values()))
        g = capi.igv_track_gather_args()
        g.rule = int(rule)
        sel = np.ascontiguousarray(list(selected_slots), dtype=np.int32)
        g.n_selected = len(sel)
        g.selected_slots = sel.ctypes.data_as(capi.c_ip) if len(sel) else None
        g.min_obs = int(min_obs if min_obs is not None else (3 if self.stereo else 4))
        g.dof_fixed = int(dof_fixed)
        g.n_feats, g.obs_slots = F, SW
        for k in dt:
            setattr(g, k, a[k].ptr)
        self._ck(self.lib.igv_tracks_gather(self.h, C.byref(g)))
        return out

    def commit_triangulation(self, track_entry, pf, ok, feat_ok=None):
        """Second half of FeatureInfoManager::triangulateFeatureInfo*: store landmark values, set _isTri; feat_ok
        (in/out, optional) &= ok."""
        a = [_Arg(track_entry, np.int32), _Arg(pf, np.float64), _Arg(ok, np.uint8), _Arg(feat_ok, np.uint8)]
        if feat_ok is not None and a[3].keep is not feat_ok:
            raise ValueError("commit_triangulation: feat_ok must be a contiguous uint8 array (updated in place)")
        self._set_mode(a)
        F = int(a[0].keep.shape[1])
        self._ck(self.lib.igv_tracks_commit_tri(self.h, F, a[0].ptr, a[1].ptr, a[2].ptr, a[3].ptr))
        return feat_ok

    def erase_tracks(self, track_entry):
        a = [_Arg(track_entry, np.int32)]
        self._set_mode(a)
        self._ck(self.lib.igv_tracks_erase(self.h, int(a[0].keep.shape[1]), a[0].ptr))

    def clean_obs_at(self, slots):
        """SwMargUpdate / KeyframeUpdate ::clean{Mono,Stereo}ObsAtMargTime for the clones at these window slots."""
        s = np.ascontiguousarray(list(slots), dtype=np.int32)
        self._ck(self.lib.igv_tracks_clean_obs(self.h, len(s), s.ctypes.data_as(capi.c_ip)))

    def change_msckf_anchor(self, old_slots, min_depth):
        """changeMSCKFAnchor: SwMargUpdate uses min_depth 0, KeyframeUpdate 0.3."""
        s = np.ascontiguousarray(list(old_slots), dtype=np.int32)
        self._ck(self.lib.igv_tracks_change_anchor(self.h, len(s), s.ctypes.data_as(capi.c_ip), float(min_depth)))

    def erase_invalid_features(self, min_depth=0.2):
        """MapServerManager::eraseInvalidFeatures."""
        self._ck(self.lib.igv_tracks_erase_invalid(self.h, float(min_depth)))

    def get_map_server(self, obs_slots=None, with_obs=True):
        """Host dump of the table in entry order (see igv_track_dump)."""
        B, T = self.B, self.max_tracks
        SW = int(obs_slots or self.max_clones)
        out = dict(id=np.zeros((B, T), np.int32), used=np.zeros((B, T), np.uint8), to_marg=np.zeros((B, T), np.uint8),
                   is_tri=np.zeros((B, T), np.uint8), slot_mask=np.zeros((B, T), np.uint64),
                   anchor_slot=np.zeros((B, T), np.int32), pf=np.zeros((B, T, 3)), pf_fej=np.zeros((B, T, 3)),
                   n_tracks=np.zeros(B, np.int32))
        if with_obs:
            out["obs"] = np.zeros((B, T, SW, self.rho))
        self._set_mode([_Arg(out["id"], np.int32)])
        d = capi.igv_track_dump()
        d.obs_slots = SW
        for k, v in out.items():
            setattr(d, k, C.c_void_p(v.ctypes.data))
        self._ck(self.lib.igv_tracks_get(self.h, C.byref(d)))
        return out

    def flags(self, clear=True):
        f = np.zeros(self.B, dtype=np.int32)
        self._set_mode([_Arg(f, np.int32)])
        self._ck(self.lib.igv_get_flags(self.h, C.c_void_p(f.ctypes.data), int(clear)))
        return f

    def cov_trace(self):
        t = np.zeros(self.B)
        self._set_mode([_Arg(t, np.float64)])
        self._ck(self.lib.igv_cov_trace(self.h, C.c_void_p(t.ctypes.data)))
        return t
