"""Whole-frame driver for the track-table tests: the visual part of IngvioFilter::callbackMonoFrame / StereoFrame
(IngvioFilter.cpp:143-205) on the oracle, fed by tracker messages through oracle/ingvio_oracle/map_server.py. The GPU
test runs the same frames through ingvio_b200.map_server.DeviceMapServer and compares state and covariance."""
import numpy as np

import ingvio_oracle as o
from ingvio_oracle import StateManager as SM
from ingvio_oracle import map_server as oms


class OracleFrontEnd:
    """One sequence: OracleFilter + MapServer + Triangulator, non-keyframe (SwMargUpdate) or keyframe mode."""

    def __init__(self, oracle_filter, keyframe=False, max_valid=20, max_lm_feats=0):
        self.f = oracle_filter
        self.max_lm = max_lm_feats
        self.lm = o.LandmarkUpdate(oracle_filter.fp) if max_lm_feats > 0 else None
        if max_lm_feats > 0:
            oracle_filter.state.state_params.max_landmarks = max_lm_feats
        self.ms = oms.MapServer()
        self.tri = o.Triangulator(o.TriParams())
        self.keyframe = keyframe
        self.f.remove_lost.max_valid_ids = max_valid
        self.counts = dict(lost=0, lost_used=0, sel=0)

    def frame(self, fr, n_meas, ids, uv):
        f, st, ms, stereo = self.f, self.f.state, self.ms, self.f.stereo
        f.propagate_augment(fr)
        oms.collect_meas(ms, st, ids[:n_meas], uv[:n_meas], stereo)
        # RemoveLostUpdate::updateState*
        oms.mark_marg_features(ms, st, stereo)
        upd = oms.select_lost(ms, self.tri, st, stereo)
        self.counts["lost"] += len(upd)
        if upd:
            f.remove_lost.last_gammas = []
            f.remove_lost.update_with_ids(st, ms, upd, stereo)
            self.counts["lost_used"] += sum(1 for _ in f.remove_lost.last_gammas)
            for k in upd:
                del ms[k]
        # SwMargUpdate / KeyframeUpdate
        if self.keyframe:
            marg_ts = f.keyframe.get_marg_kfs(st)
            sel_ts, dof, thr, updater = list(marg_ts), 2, 0.3, f.keyframe
        else:
            mt = st.next_marg_time()
            marg_ts = [] if mt == float("inf") else [mt]
            sel_ts = f.sw_marg.select_sw_timestamps(st.sw_camleft_poses, mt) if marg_ts else []
            dof, thr, updater = len(sel_ts) - 1, 0.0, f.sw_marg
        info = dict(marg_slots=[], sel_slots=[], thr=thr, dof_fixed=2 if self.keyframe else 0)
        if marg_ts:
            times = st.sw_times()
            info["marg_slots"] = [times.index(t) for t in marg_ts]
            info["sel_slots"] = [times.index(t) for t in sel_ts]
            for k in ms.ids():      # tri_ok = outcome of THIS frame's attempt (select_seen_at re-triangulates)
                ms[k].tri_ok = False
            upd = oms.select_seen_at(ms, self.tri, st, sel_ts, stereo)
            self.counts["sel"] += len(upd)
            updater._update_selected(st, ms, sel_ts, dof, stereo)
        if self.lm is not None:      # IngvioFilter.cpp:155-161 / :181-187 (the reference runs these every frame)
            self.lm.update_landmark_mono(st, ms)
            self.lm.init_new_landmark_mono(st, ms, lambda ft: oms.triangulate_feature_info(ft, self.tri, st, stereo),
                                           f.fp.max_sw_clones)
        if marg_ts:
            oms.clean_obs_at(ms, marg_ts, stereo)
            oms.change_msckf_anchor(ms, st, marg_ts, thr)
        if self.lm is not None:      # :167-173 (keyframes: the clones about to leave) / :193-194 (next marg time)
            if self.keyframe:
                self.lm.change_landmark_anchor(st, ms, list(marg_ts))
            else:
                self.lm.change_landmark_anchor(st, ms)
        if marg_ts:
            for t in marg_ts:
                SM.marg_sliding_window_pose(st, t)
        oms.erase_invalid_features(ms, 0.2, state=st)
        return info
