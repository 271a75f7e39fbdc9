"""Frame packets: the per-frame inputs of the hot path in the array layout the C-ABI consumes.

A packet holds, for B independent sequences, everything one cycle of
IngvioFilter::callbackMonoFrame (/root/reference/ingvio_estimator/src/IngvioFilter.cpp:124-234)
feeds into the filter core *after* the ROS / MapServer / gnss_comm front-ends:

  * IMU samples of the frame interval            (ImuPropagator.cpp:246-272 loop inputs)
  * triangulated MSCKF tracks over the window    (MapServer.h:69-132 contents; Triangulator output)
  * one GNSS epoch at the psr_res/dopp_res output boundary (gnss_spp.cpp:99-146, :256-282)

Pure numpy; no CUDA dependency (the oracle-side driver consumes the same object).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class GnssArrays:
    """Batched GNSS epoch. `sys` uses State::GNSSType numbering (GPS=0, GLO, GAL, BDS)."""
    unit: np.ndarray          # (B,S,3) receiver->satellite unit vectors, ECEF
    res_pos: np.ndarray       # (B,S)
    res_vel: np.ndarray       # (B,S)
    sys: np.ndarray           # (B,S) int32
    ura: np.ndarray           # (B,S)
    psr_std: np.ndarray       # (B,S)
    dopp_std_mps: np.ndarray  # (B,S)
    el: np.ndarray            # (B,S)
    R_enu2ecef: np.ndarray    # (B,3,3)

    def sigma_psr(self, amp=1.0):
        s = np.sin(self.el)
        s = np.where(np.abs(s) < 1e-6, 1e-6, s)
        return amp * np.sqrt(self.ura * self.psr_std / (s * s))    # GnssUpdate.cpp:180-187

    def sigma_dopp(self, amp=1.0):
        s = np.sin(self.el)
        s = np.where(np.abs(s) < 1e-6, 1e-6, s)
        return amp * np.sqrt(self.ura * self.dopp_std_mps / (s * s))  # GnssUpdate.cpp:249-256


@dataclass
class FramePacket:
    t: float
    gyro: np.ndarray          # (B,K,3)
    accel: np.ndarray         # (B,K,3)
    dt: np.ndarray            # (B,K)
    pf_w: np.ndarray          # (B,F,3)   triangulated landmark, world frame
    anchor_slot: np.ndarray   # (B,F) int32  window slot (0 = oldest clone after augmentation)
    obs: np.ndarray           # (B,F,SW,rho) normalised image coords, rho = 2 mono / 4 stereo
    obs_mask: np.ndarray      # (B,F,SW) uint8
    obs_total: np.ndarray     # (B,F) int32 size of the track's observation map (chi^2 dof rule)
    visual_mode: Optional[str] = "all_obs"     # "all_obs" | "keyframe" | "sw_marg" | None
    selected_slots: List[int] = field(default_factory=list)
    marg_slots: List[int] = field(default_factory=list)
    max_valid: int = 20
    gnss: Optional[GnssArrays] = None

    @property
    def batch(self):
        return self.gyro.shape[0]

    def tiled(self, B):
        """The same packet repeated to B sequences (sequence b is a copy of b % batch): a chip-filling batch built
        from a few distinct streams."""
        def t(a):
            reps = -(-B // a.shape[0])
            return np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:B])
        g = None
        if self.gnss is not None:
            g = GnssArrays(**{k: t(getattr(self.gnss, k)) for k in self.gnss.__dataclass_fields__})
        return FramePacket(t=self.t, gyro=t(self.gyro), accel=t(self.accel), dt=t(self.dt), pf_w=t(self.pf_w),
                           anchor_slot=t(self.anchor_slot), obs=t(self.obs), obs_mask=t(self.obs_mask),
                           obs_total=t(self.obs_total), visual_mode=self.visual_mode,
                           selected_slots=list(self.selected_slots), marg_slots=list(self.marg_slots),
                           max_valid=self.max_valid, gnss=g)

    def seq(self, b):
        """Single-sequence view with the attribute names oracle/ingvio_oracle/frame.py reads."""
        g = None
        R = None
        if self.gnss is not None:
            G = self.gnss
            g = dict(unit=G.unit[b], res_pos=G.res_pos[b], res_vel=G.res_vel[b], sys=G.sys[b],
                     ura=G.ura[b], psr_std=G.psr_std[b], dopp_std_mps=G.dopp_std_mps[b], el=G.el[b])
            R = G.R_enu2ecef[b]
        return _SeqFrame(t=self.t, gyro=self.gyro[b], accel=self.accel[b], dt=self.dt[b],
                         pf_w=self.pf_w[b], anchor_slot=self.anchor_slot[b], obs=self.obs[b],
                         obs_mask=self.obs_mask[b], obs_total=self.obs_total[b],
                         visual_mode=self.visual_mode, selected_slots=list(self.selected_slots),
                         marg_slots=list(self.marg_slots), max_valid=self.max_valid, gnss=g,
                         R_enu2ecef=R)


@dataclass
class _SeqFrame:
    t: float
    gyro: np.ndarray
    accel: np.ndarray
    dt: np.ndarray
    pf_w: np.ndarray
    anchor_slot: np.ndarray
    obs: np.ndarray
    obs_mask: np.ndarray
    obs_total: np.ndarray
    visual_mode: Optional[str]
    selected_slots: list
    marg_slots: list
    max_valid: int
    gnss: Optional[dict]
    R_enu2ecef: Optional[np.ndarray]
