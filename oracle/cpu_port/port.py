"""ctypes wrapper of the oracle's C++ CPU port (oracle/cpu_port/ingvio_cpu_port.cpp).

TEST / MEASUREMENT INFRASTRUCTURE. Used by tests/ (cross-check against the numpy oracle) and by
bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(_HERE), "_build", "libingvio_cpu_port.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ingvio_cpu_port.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_init.argtypes = [C.c_void_p] + [_dp] * 8
        _lib.orc_add_gnss.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        for n in ("orc_dim", "orc_num_clones", "orc_n_accepted"):
            getattr(_lib, n).restype = C.c_int
            getattr(_lib, n).argtypes = [C.c_void_p]
        _lib.orc_get_cov.argtypes = [C.c_void_p, _dp]
        _lib.orc_get_state.argtypes = [C.c_void_p, _dp]
        _lib.orc_get_gammas.argtypes = [C.c_void_p, _dp, C.c_int]
        _lib.orc_step.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int, _dp, _ip, _dp, _up, _ip, C.c_int,
                                  C.c_double, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _dp,
                                  C.c_int, C.c_int]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class CpuPortFilter:
    """One sequence of the C++ port; `step(frame)` takes the same per-sequence frame view as the numpy oracle."""

    def __init__(self, noise6, gravity, T_cl2cr, stereo, chi2_table):
        self.lib = load()
        k = [_d(noise6), _d(gravity), _d(np.asarray(T_cl2cr[0]).reshape(9)), _d(T_cl2cr[1]), _d(chi2_table)]
        self.h = C.c_void_p(self.lib.orc_create(k[0][1], k[1][1], k[2][1], k[3][1], int(stereo), k[4][1], len(chi2_table)))
        self.stereo = bool(stereo)

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def init(self, R, p, v, bg, ba, Rext, pext, diag21):
        k = [_d(np.asarray(R).reshape(9)), _d(p), _d(v), _d(bg), _d(ba), _d(np.asarray(Rext).reshape(9)), _d(pext), _d(diag21)]
        self.lib.orc_init(self.h, *[x[1] for x in k])

    def add_gnss(self, gtype, value, cov):
        self.lib.orc_add_gnss(self.h, int(gtype), float(value), float(cov))

    def step(self, fr, noise, psr_amp=1.0, dopp_amp=1.0, adjust_yof=0, strong_reject=1):
        K = fr.dt.shape[0]
        g, a, dt = _d(fr.gyro), _d(fr.accel), _d(fr.dt)
        F = fr.pf_w.shape[0] if fr.visual_mode == "all_obs" else 0
        pf, anc, obs = _d(fr.pf_w), _i(fr.anchor_slot), _d(fr.obs)
        m = np.ascontiguousarray(fr.obs_mask, dtype=np.uint8)
        dof = _i(np.asarray(fr.obs_total, dtype=np.int32) - 1)
        marg = _i(sorted(fr.marg_slots, reverse=True) or [0])
        if fr.gnss is not None:
            gn = fr.gnss
            s = np.sin(gn["el"])
            s = np.where(np.abs(s) < 1e-6, 1e-6, s)
            sp = psr_amp * np.sqrt(gn["ura"] * gn["psr_std"] / (s * s))
            sd = dopp_amp * np.sqrt(gn["ura"] * gn["dopp_std_mps"] / (s * s))
            S = gn["unit"].shape[0]
            gu, rp, rv, gsp, gsd, gs, Re = (_d(gn["unit"]), _d(gn["res_pos"]), _d(gn["res_vel"]), _d(sp), _d(sd),
                                            _i(gn["sys"]), _d(np.asarray(fr.R_enu2ecef).reshape(9)))
        else:
            S = 0
            z = _d(np.zeros(9))
            gu = rp = rv = gsp = gsd = Re = z
            gs = _i(np.zeros(1))
        self.lib.orc_step(self.h, K, g[1], a[1], dt[1], F, pf[1], anc[1], obs[1], m.ctypes.data_as(_up), dof[1],
                          fr.obs_mask.shape[1], float(noise), int(fr.max_valid), len(fr.marg_slots), marg[1], S,
                          gu[1], rp[1], rv[1], gsp[1], gsd[1], gs[1], Re[1], int(adjust_yof), int(strong_reject))

    def cov(self):
        n = self.lib.orc_dim(self.h)
        out = np.empty((n, n))
        self.lib.orc_get_cov(self.h, out.ctypes.data_as(_dp))
        return out

    def state(self):
        out = np.zeros(39 + 12 * self.lib.orc_num_clones(self.h))
        self.lib.orc_get_state(self.h, out.ctypes.data_as(_dp))
        return out

    def n_accepted(self):
        return self.lib.orc_n_accepted(self.h)
