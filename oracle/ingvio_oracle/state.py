"""State container (covariance + ordered error variables).

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/State.h:36-136 and State.cpp:25-167.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from .types import SE3, SE23, Vec3

# State.h:75  enum GNSSType {GPS = 0, GLO, GAL, BDS, FS, YOF}
GPS, GLO, GAL, BDS, FS, YOF = 0, 1, 2, 3, 4, 5


@dataclass
class FilterParams:
    """The subset of IngvioParams the hot path reads (values: config/sportsfield/ingvio_mono.yaml:10-76)."""
    cam_nums: int = 1
    max_sw_clones: int = 11
    max_lm_feats: int = 0
    enable_gnss: int = 1
    noise_g: float = 0.004
    noise_a: float = 0.08
    noise_bg: float = 0.0002
    noise_ba: float = 0.008
    noise_clockbias: float = 2.0
    noise_cb_rw: float = 0.2
    init_cov_rot: float = 0.0
    init_cov_pos: float = 0.0
    init_cov_vel: float = 0.25
    init_cov_bg: float = 0.01
    init_cov_ba: float = 0.01
    init_cov_ext_rot: float = 1.8e-2
    init_cov_ext_pos: float = 2e-3
    init_cov_rcv_clockbias: float = 2.0
    init_cov_rcv_clockbias_randomwalk: float = 1.0
    init_cov_yof: float = 0.015
    gravity_norm: float = 9.8
    chi2_max_dof: int = 150
    chi2_thres: float = 0.95
    visual_noise: float = 0.12
    frame_select_interval: int = 28
    psr_noise_amp: float = 1.0
    dopp_noise_amp: float = 1.0
    is_adjust_yof: int = 0
    gnss_chi2_test: int = 0
    gnss_strong_reject: int = 1
    is_key_frame: int = 1
    T_cl2i_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    T_cl2i_p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    T_cl2cr_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    T_cl2cr_p: np.ndarray = field(default_factory=lambda: np.zeros(3))


class StateParams:
    """State.cpp:25-58, including the clock-noise mix-up at :51-52:
    `_noise_clockbias` is assigned twice (second time from `noise_cb_rw`) and
    `_noise_cb_rw` is never assigned, so it keeps its default 0.2 (State.h:52-53)."""

    def __init__(self, fp: FilterParams = None):
        self.noise_g = 0.005
        self.noise_a = 0.05
        self.noise_bg = 0.001
        self.noise_ba = 0.01
        self.noise_clockbias = 2.0
        self.noise_cb_rw = 0.2
        self.cam_nums = 2
        self.max_sw_poses = 20
        self.max_landmarks = 25
        self.enable_gnss = True
        self.T_cl2cr_R, self.T_cl2cr_p = np.eye(3), np.zeros(3)
        self.T_cl2i_R, self.T_cl2i_p = np.eye(3), np.zeros(3)
        self.init_cov_rot = self.init_cov_pos = self.init_cov_vel = 0.0
        self.init_cov_bg = self.init_cov_ba = 0.0
        self.init_cov_ext_rot = self.init_cov_ext_pos = 0.0
        self.init_cov_rcv_clockbias = self.init_cov_rcv_clockbias_randomwalk = 0.0
        self.init_cov_yof = 0.0
        if fp is None:
            return
        self.cam_nums = fp.cam_nums
        self.max_sw_poses = fp.max_sw_clones
        self.max_landmarks = fp.max_lm_feats          # State.cpp:29
        self.T_cl2cr_R, self.T_cl2cr_p = np.array(fp.T_cl2cr_R, float), np.array(fp.T_cl2cr_p, float)
        self.T_cl2i_R, self.T_cl2i_p = np.array(fp.T_cl2i_R, float), np.array(fp.T_cl2i_p, float)
        self.enable_gnss = bool(fp.enable_gnss)
        self.noise_a, self.noise_g = fp.noise_a, fp.noise_g
        self.noise_ba, self.noise_bg = fp.noise_ba, fp.noise_bg
        self.init_cov_rot, self.init_cov_pos, self.init_cov_vel = fp.init_cov_rot, fp.init_cov_pos, fp.init_cov_vel
        self.init_cov_bg, self.init_cov_ba = fp.init_cov_bg, fp.init_cov_ba
        self.init_cov_ext_rot, self.init_cov_ext_pos = fp.init_cov_ext_rot, fp.init_cov_ext_pos
        if self.enable_gnss:
            self.noise_clockbias = fp.noise_clockbias
            self.noise_clockbias = fp.noise_cb_rw  # State.cpp:52 (quirk kept on purpose)
            self.init_cov_rcv_clockbias = fp.init_cov_rcv_clockbias
            self.init_cov_rcv_clockbias_randomwalk = fp.init_cov_rcv_clockbias_randomwalk
            self.init_cov_yof = fp.init_cov_yof


class State:
    """State.cpp:60-91 (ctor: SE23@0, bg@9, ba@12, cam-IMU extrinsics@15 -> 21; cov = 1e-6 I)."""

    def __init__(self, fp: FilterParams = None):
        self.state_params = StateParams(fp)
        self.timestamp = -1.0
        self.err_variables = []
        idx = 0
        self.extended_pose = SE23()
        self.bg = Vec3()
        self.ba = Vec3()
        self.camleft_imu_extrinsics = SE3()
        for var in (self.extended_pose, self.bg, self.ba, self.camleft_imu_extrinsics):
            var.set_cov_idx(idx)
            self.err_variables.append(var)
            idx += var.size()
        self.gnss = {}
        self.sw_camleft_poses = {}  # timestamp -> SE3 (std::map ordered by key; iterate sorted())
        self.anchored_landmarks = {}
        self.cov = (1e-3 ** 2) * np.eye(idx)
        self.camleft_imu_extrinsics.set_value(self.state_params.T_cl2i_R, self.state_params.T_cl2i_p)

    def curr_cov_size(self):
        return self.cov.shape[0]

    def curr_err_variable_size(self):
        return len(self.err_variables)

    def sw_times(self):
        return sorted(self.sw_camleft_poses.keys())

    def next_marg_time(self):
        """State.h:85-95."""
        t = math.inf
        if len(self.sw_camleft_poses) > self.state_params.max_sw_poses:
            t = min(self.sw_camleft_poses.keys())
        return t

    def init_state_and_cov(self, t0, R_i2w, pos=None, vel=None, bg=None, ba=None):
        """State.cpp:126-167 (every diagonal entry of the 21x21 block is overwritten)."""
        sp = self.state_params
        self.timestamp = float(t0)
        d = np.array([sp.init_cov_rot] * 3 + [sp.init_cov_pos] * 3 + [sp.init_cov_vel] * 3
                     + [sp.init_cov_bg] * 3 + [sp.init_cov_ba] * 3
                     + [sp.init_cov_ext_rot] * 3 + [sp.init_cov_ext_pos] * 3) ** 2.0
        for i in range(21):
            self.cov[i, i] = d[i]
        self.extended_pose.rot = np.array(R_i2w, dtype=np.float64).reshape(3, 3)
        self.extended_pose.vec1 = np.zeros(3) if pos is None else np.array(pos, dtype=np.float64)
        self.extended_pose.vec2 = np.zeros(3) if vel is None else np.array(vel, dtype=np.float64)
        self.bg.set_value(np.zeros(3) if bg is None else bg)
        self.ba.set_value(np.zeros(3) if ba is None else ba)
        self.camleft_imu_extrinsics.set_value(sp.T_cl2i_R, sp.T_cl2i_p)
