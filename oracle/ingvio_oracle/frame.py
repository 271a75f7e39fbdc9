"""Per-frame driver: the call order of IngvioFilter::callbackMonoFrame / callbackStereoFrame.

TEST INFRASTRUCTURE (oracle). Follows /root/reference/ingvio_estimator/src/IngvioFilter.cpp:124-234
(mono) / :252-361 (stereo): propagate+augment -> visual update -> marginalise clones -> GNSS update,
with the ROS / MapServer bookkeeping replaced by plain per-frame arrays (the "frame packet" that the
C-ABI consumes, include/ingvio_b200.h).
"""
import numpy as np

from .gnss_update import GnssEpoch, GnssUpdate
from .imu_propagator import ImuPropagator
from .state import FilterParams, State
from .state_manager import StateManager
from .visual_update import FeatureInfo, KeyframeUpdate, RemoveLostUpdate, SwMargUpdate


class OracleFilter:
    """One sequence. `frame` objects are duck-typed (see ingvio_b200/frames.py: FramePacket)."""

    def __init__(self, fp: FilterParams, stereo=False, max_valid_ids=20):
        self.fp = fp
        self.stereo = stereo
        self.state = State(fp)
        self.prop = ImuPropagator(fp.gravity_norm)
        self.remove_lost = RemoveLostUpdate(fp, max_valid_ids)
        self.keyframe = KeyframeUpdate(fp)
        self.sw_marg = SwMargUpdate(fp)
        self.gnss = GnssUpdate(fp)
        self.last = {}

    def init(self, t0, R_i2w, pos, vel, bg, ba):
        self.state.init_state_and_cov(t0, R_i2w, pos, vel, bg, ba)

    # ---- steps ------------------------------------------------------------------------------
    def propagate_augment(self, frame):
        self.prop.propagate_steps(self.state, frame.gyro, frame.accel, frame.dt)
        self.state.timestamp = float(frame.t)  # propagateUntil ends exactly at t_end
        StateManager.augment_sliding_window_pose(self.state)

    def build_map_server(self, frame):
        st = self.state
        times = st.sw_times()
        ms = {}
        F = frame.pf_w.shape[0]
        for f in range(F):
            fi = FeatureInfo(f, frame.pf_w[f], st.sw_camleft_poses[times[int(frame.anchor_slot[f])]])
            obs = fi.stereo_obs if self.stereo else fi.mono_obs
            for s, t in enumerate(times):
                if frame.obs_mask[f, s]:
                    obs[t] = np.array(frame.obs[f, s, :], dtype=np.float64)
            extra = int(frame.obs_total[f]) - len(obs) if getattr(frame, "obs_total", None) is not None else 0
            for k in range(extra):  # observations at clones that already left the window
                obs[-1.0 - k] = np.zeros(4 if self.stereo else 2)
            ms[f] = fi
        return ms

    def visual_update(self, frame):
        mode = frame.visual_mode
        if mode is None or frame.pf_w.shape[0] == 0:
            return None
        st = self.state
        ms = self.build_map_server(frame)
        times = st.sw_times()
        if mode == "all_obs":
            upd = self.remove_lost
            upd.last_gammas = []
            upd.max_valid_ids = int(frame.max_valid)
            dx = upd.update_with_ids(st, ms, sorted(ms.keys()), self.stereo, keep="cols")
        elif mode in ("keyframe", "sw_marg"):
            upd = self.keyframe if mode == "keyframe" else self.sw_marg
            sel = [times[s] for s in frame.selected_slots]
            dof = 2 if mode == "keyframe" else len(sel) - 1
            dx = upd._update_selected(st, ms, sel, dof, self.stereo)
        else:
            raise ValueError(mode)
        self.last["gammas"] = list(upd.last_gammas)
        return dx

    def marginalize(self, frame):
        times = self.state.sw_times()
        for s in sorted(frame.marg_slots, reverse=True):
            StateManager.marg_sliding_window_pose(self.state, times[s])

    def gnss_update(self, frame):
        if getattr(frame, "gnss", None) is None:
            return None
        g = frame.gnss
        ep = g if isinstance(g, GnssEpoch) else GnssEpoch(**g)
        dx = self.gnss.update_tracked_sys(self.state, ep, np.asarray(frame.R_enu2ecef, float))
        self.last["gnss_gammas"] = list(self.gnss.last_gammas)
        return dx

    def step(self, frame):
        """IngvioFilter.cpp:143-231 order."""
        self.propagate_augment(frame)
        dxv = self.visual_update(frame)
        self.marginalize(frame)
        dxg = self.gnss_update(frame)
        return dxv, dxg

    # ---- read-outs used by parity tests ---------------------------------------------------------
    def cov(self):
        return self.state.cov.copy()

    def pose(self):
        e = self.state.extended_pose
        return e.rot.copy(), e.vec1.copy(), e.vec2.copy()
