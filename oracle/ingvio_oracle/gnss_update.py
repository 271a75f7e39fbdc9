"""GNSS pseudo-range / Doppler update and delayed initialisation of clock states.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/GnssUpdate.cpp:84-293 (updateTrackedSys),
:317-521 (addNewTrackedSys, getResJacobianOfSys) and GnssManager.cpp:60-134.

Inputs enter at the output boundary of gnss_comm::psr_res / dopp_res
(/root/reference/gnss_comm/src/gnss_spp.cpp:99-146, :256-282): per satellite the
receiver->satellite unit vector (= -J[:, 0:3]), the residuals `res_pos`, `res_vel`, the
constellation, and the quantities the noise model reads (ura, psr_std, dopp_std*c/f, elevation).
The ephemeris / atmosphere code before that boundary is out of scope.
"""
import math
from dataclasses import dataclass

import numpy as np

from .lie import skew
from .state import FS, YOF, State
from .state_manager import StateManager
from .types import Scalar
from .visual_update import UpdateBase


@dataclass
class GnssEpoch:
    unit: np.ndarray        # (S,3) unit receiver->satellite, ECEF
    res_pos: np.ndarray     # (S,)  psr_estimated - psr_measured   (gnss_spp.cpp:141)
    res_vel: np.ndarray     # (S,)  dopp_estimated + dopp*lambda   (gnss_spp.cpp:279)
    sys: np.ndarray         # (S,)  GNSSType GPS..BDS of each satellite
    ura: np.ndarray         # (S,)
    psr_std: np.ndarray     # (S,)
    dopp_std_mps: np.ndarray  # (S,) dopp_std * c / f
    el: np.ndarray          # (S,) elevation [rad]

    def psr_noise(self, amp):
        s = np.sin(self.el)
        s = np.where(np.abs(s) < 1e-6, 1e-6, s)
        return amp * np.sqrt(self.ura * self.psr_std / (s * s))   # GnssUpdate.cpp:180-187

    def dopp_noise(self, amp):
        s = np.sin(self.el)
        s = np.where(np.abs(s) < 1e-6, 1e-6, s)
        return amp * np.sqrt(self.ura * self.dopp_std_mps / (s * s))  # GnssUpdate.cpp:249-256


def calc_R_w2enu(yaw):
    """GnssManager.cpp:60-63 (AngleAxis about +z)."""
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def dot_R_w2enu(yaw):
    """GnssManager.cpp:100-114."""
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[-s, -c, 0.0], [c, -s, 0.0], [0.0, 0.0, 0.0]])


def check_gnss_states(state: State):
    """GnssManager.cpp:86-98."""
    if YOF not in state.gnss or FS not in state.gnss:
        return False
    return any(i in state.gnss for i in range(4))


class GnssUpdate(UpdateBase):
    """GnssUpdate.h:36-110."""

    def __init__(self, fp):
        super().__init__(fp.chi2_max_dof, fp.chi2_thres)
        self.psr_noise_amp = fp.psr_noise_amp
        self.dopp_noise_amp = fp.dopp_noise_amp
        self.is_adjust_yof = bool(fp.is_adjust_yof)
        self.is_gnss_chi2_test = bool(fp.gnss_chi2_test)
        self.is_gnss_strong_reject = bool(fp.gnss_strong_reject)

    def build_rows(self, state: State, ep: GnssEpoch, R_enu2ecef):
        """The row-building part of GnssUpdate.cpp:124-284. Returns (var_order, H, res, R)."""
        ext = state.extended_pose
        yof = state.gnss[YOF]
        var_order = [ext, yof]
        local = {id(ext): 0, id(yof): 9}
        S = ep.unit.shape[0]
        max_rows, max_cols = 2 * S, ext.size() + 6
        Rw2ecef = R_enu2ecef @ calc_R_w2enu(yof.value())
        dRw = R_enu2ecef @ dot_R_w2enu(yof.value())
        res = np.zeros(max_rows)
        H = np.zeros((max_rows, max_cols))
        Rm = np.zeros((max_rows, max_rows))
        row_cnt, col_cnt = 0, 10
        psr_sig = ep.psr_noise(self.psr_noise_amp)
        dop_sig = ep.dopp_noise(self.dopp_noise_amp)
        for i in range(S):
            u = ep.unit[i]
            g = int(ep.sys[i])
            if g not in state.gnss:
                continue
            cb = state.gnss[g]
            H_i = np.zeros((1, 11))
            H_i[0, 0:3] = u @ Rw2ecef @ skew(ext.value_trans1())
            H_i[0, 3:6] = -u @ Rw2ecef
            if self.is_adjust_yof:
                H_i[0, 9] = -u @ dRw @ ext.value_trans1()
            H_i[0, 10] = 1.0
            res_i = np.array([-ep.res_pos[i]])
            if self.is_gnss_chi2_test and not self.test_chi_squared(state, res_i, H_i, [ext, yof, cb],
                                                                    float(psr_sig[i])):
                continue
            res[row_cnt] = res_i[0]
            Rm[row_cnt, row_cnt] = psr_sig[i] ** 2
            H[row_cnt, 0:9] = H_i[0, 0:9]
            H[row_cnt, 9] = H_i[0, 9]
            if id(cb) not in local:
                local[id(cb)] = col_cnt
                col_cnt += 1
                var_order.append(cb)
            H[row_cnt, local[id(cb)]] = 1.0
            row_cnt += 1
        fs = state.gnss[FS]
        local[id(fs)] = col_cnt
        col_cnt += 1
        var_order.append(fs)
        for i in range(S):
            u = ep.unit[i]
            g = int(ep.sys[i])
            if g not in state.gnss:
                continue
            H_i = np.zeros((1, 11))
            H_i[0, 0:3] = u @ Rw2ecef @ skew(ext.value_trans2())
            H_i[0, 6:9] = -u @ Rw2ecef
            if self.is_adjust_yof:
                H_i[0, 9] = -u @ dRw @ ext.value_trans2()
            H_i[0, 10] = 1.0
            res_i = np.array([-ep.res_vel[i]])
            if self.is_gnss_chi2_test and not self.test_chi_squared(state, res_i, H_i, [ext, yof, fs],
                                                                    float(dop_sig[i])):
                continue
            res[row_cnt] = res_i[0]
            Rm[row_cnt, row_cnt] = dop_sig[i] ** 2
            H[row_cnt, 0:9] = H_i[0, 0:9]
            H[row_cnt, 9] = H_i[0, 9]
            H[row_cnt, local[id(fs)]] = H_i[0, 10]
            row_cnt += 1
        return var_order, H[:row_cnt, :col_cnt], res[:row_cnt], Rm[:row_cnt, :row_cnt]

    def update_tracked_sys(self, state: State, ep: GnssEpoch, R_enu2ecef, is_aligned=True):
        """GnssUpdate.cpp:84-293."""
        self.last_gammas = []
        if not state.state_params.enable_gnss or not is_aligned:
            return None
        if ep.unit.shape[0] <= 0:
            return None
        if not check_gnss_states(state):
            return None
        var_order, H, res, Rm = self.build_rows(state, ep, R_enu2ecef)
        if res.shape[0] <= 14 and self.is_gnss_strong_reject and \
                not self.test_chi_squared(state, res, H, var_order, Rm, res.shape[0]):
            return None
        return StateManager.ekf_update(state, var_order, H, res, Rm, return_dx=True)

    def add_new_tracked_sys(self, state: State, ep: GnssEpoch, R_enu2ecef, sys_to_add, spp_values,
                            R_ecef2enu=None):
        """GnssUpdate.cpp:317-476 with `sys_to_add` (set of GNSSType) and the SPP initial values
        (dict GNSSType -> value) supplied by the caller (getSysInSppMeas/calcSysToAdd are bookkeeping).
        Residuals in `ep` must have been evaluated with those initial values, as the reference does
        (:351-370). Returns {gtype: accepted}."""
        out = {}
        if YOF not in state.gnss:
            return out
        ext = state.extended_pose
        yofv = state.gnss[YOF].value()
        Rw2ecef = R_enu2ecef @ calc_R_w2enu(yofv)
        if R_ecef2enu is None:
            R_ecef2enu = R_enu2ecef.T
        dR_quirk = R_ecef2enu @ dot_R_w2enu(yofv)  # getRecef2enu() here vs getRenu2ecef() above (:403,:464)
        for g in sys_to_add:
            x_order = [ext, state.gnss[YOF]]
            if g == FS:
                res = -ep.res_vel.copy()
                U = ep.unit
                Hx = np.zeros((res.shape[0], 10))
                Hx[:, 0:3] = U @ Rw2ecef @ skew(ext.value_trans2())
                Hx[:, 6:9] = -U @ Rw2ecef
                if self.is_adjust_yof:
                    Hx[:, 9] = -U @ dR_quirk @ ext.value_trans2()
                s = np.sin(ep.el)
                s = np.where(np.abs(s) < 1e-6, 1e-6, s)
                avg = self.dopp_noise_amp * math.sqrt(float(np.mean(ep.ura * ep.dopp_std_mps / (s * s))))
            else:
                rows = np.nonzero(ep.sys == g)[0]
                res = -ep.res_pos[rows]
                U = ep.unit[rows]
                Hx = np.zeros((res.shape[0], 10))
                Hx[:, 0:3] = U @ Rw2ecef @ skew(ext.value_trans1())
                Hx[:, 3:6] = -U @ Rw2ecef
                if self.is_adjust_yof:
                    Hx[:, 9] = -U @ dR_quirk @ ext.value_trans1()
                s = np.sin(ep.el[rows])
                s = np.where(np.abs(s) < 1e-6, 1e-6, s)
                avg = self.psr_noise_amp * math.sqrt(float(np.mean(ep.ura[rows] * ep.psr_std[rows] / (s * s))))
            Hf = np.ones((res.shape[0], 1))
            var = Scalar()
            var.set_value(spp_values[g])
            ok = StateManager.add_variable_delayed(state, var, x_order, Hx, Hf, res, avg, 0.95, True)
            if ok:
                state.gnss[g] = var
            out[g] = ok
        return out
