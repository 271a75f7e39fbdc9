"""Track table ("map server"): per-feature observation maps, life cycle and the track selections of the updaters.

TEST INFRASTRUCTURE (oracle). CPU restatement of
  /root/reference/ingvio_estimator/src/MapServer.h:31-134, MapServer.cpp:20-83                (MonoMeas / StereoMeas / FeatureInfo)
  /root/reference/ingvio_estimator/src/MapServerManager.cpp:101-273                            (collect*Meas, markMarg*Features)
  /root/reference/ingvio_estimator/src/MapServerManager.cpp:275-341                            (triangulateFeatureInfo*)
  /root/reference/ingvio_estimator/src/MapServerManager.cpp:454-490                            (eraseInvalidFeatures)
  /root/reference/ingvio_estimator/src/RemoveLostUpdate.cpp:45-59,165-166 / :280-294,403-404   (lost-track selection, erase)
  /root/reference/ingvio_estimator/src/SwMargUpdate.cpp:61-85,191-259                          (selected-clone selection, clean, re-anchor)
  /root/reference/ingvio_estimator/src/KeyframeUpdate.cpp:251-327,455-480                      (same, two marginalised keyframes)
and of the wire format feature_tracker/msg/{MonoMeas,StereoMeas}.msg (uint64 id, float64 u0 v0 [u1 v1]).
SURVEY.md section 8f rank 4 ("next" row). `MapServer` is `std::map<int, shared_ptr<FeatureInfo>>`: a dict here, iterated
in ascending key order wherever the reference iterates the map. SLAM-type features (section 8f rank 3): the branches of
collect*Meas (:136, :178), markMarg*Features (:231, :259) and eraseInvalidFeatures (:485) are restated; the landmark algebra
itself lives in landmark_update.py.

Pinned by the reference's own test: tests/test_oracle_map_server.py restates
test/TestMapServer.cpp:184-308 (collectFeatureAndMarg, mono and stereo).
"""
import numpy as np

from .visual_update import MSCKF, SLAM, FeatureInfo


def msg_id_to_key(msg_id):
    """`_id = mono_msg.id` assigns a uint64 to an `int` (MapServer.cpp:24, MapServer.h:120): two's-complement wrap."""
    v = int(msg_id) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


class MapServer(dict):
    """std::map<int, std::shared_ptr<FeatureInfo>> (MapServer.h:134)."""

    def ids(self):
        return sorted(self.keys())


def _obs_of(feat, stereo):
    return feat.stereo_obs if stereo else feat.mono_obs


def collect_meas(map_server, state, ids, uv, stereo=False):
    """MapServerManager::collect{Mono,Stereo}Meas (MapServerManager.cpp:189-219) over one frame message, with
    FeatureInfoManager::collect{Mono,Stereo}Meas (:101-187) per measurement, in message order."""
    t = state.timestamp
    if t not in state.sw_camleft_poses:
        raise RuntimeError("[FeatureInfoManager]: Meas timestamp not in sw!")  # :107-111 (assert)
    for mid, z in zip(ids, uv):
        key = msg_id_to_key(mid)
        if key not in map_server:
            f = FeatureInfo(-1, np.zeros(3), None, tri_ok=False)  # MapServer.h:74-80 (default ctor)
            map_server[key] = f
        f = map_server[key]
        obs = _obs_of(f, stereo)
        if len(obs) == 0:
            obs[t] = np.array(z, dtype=np.float64)
            f.id = key
            f.ftype = MSCKF
            f.is_to_marg = False
            f.is_tri = False
            f.anchor = state.sw_camleft_poses[t]  # resetAnchoredPose(:122)
        else:
            if t in obs:  # "Meas timestamp already in mono obs, skip adding!" (:126-130)
                continue
            if f.ftype == SLAM:   # a landmark in the state only keeps its current observation (:136-137, :178-179)
                obs.clear()
            obs[t] = np.array(z, dtype=np.float64)
            f.is_to_marg = False


def mark_marg_features(map_server, state, stereo=False):
    """MapServerManager::markMarg{Mono,Stereo}Features (:221-273), MSCKF part."""
    from .landmark_update import marg_anchored_landmark_in_state
    t = state.timestamp
    marg_ids = []
    for key in map_server.ids():
        f = map_server[key]
        if t not in _obs_of(f, stereo):
            f.is_to_marg = True
            if f.ftype == SLAM:       # a lost landmark leaves the state and the map at once (:231-246)
                marg_ids.append(key)
    for key in marg_ids:
        marg_anchored_landmark_in_state(state, key)
        del map_server[key]


def triangulate_feature_info(feat, tri, state, stereo=False):
    """FeatureInfoManager::triangulateFeatureInfo{Mono,Stereo} (:275-341). `tri` is an oracle Triangulator."""
    times = state.sw_times()
    obs_map = _obs_of(feat, stereo)
    common = [t for t in times if t in obs_map]  # filterCommonTimestamp (Triangulator.h:96-117)
    obs = [obs_map[t] for t in common]
    poses = [(state.sw_camleft_poses[t].value_linear(), state.sw_camleft_poses[t].value_trans()) for t in common]
    if stereo:
        sp = state.state_params
        flag, pf = tri.triangulate_stereo(obs, poses, (sp.T_cl2cr_R, sp.T_cl2cr_p))
    else:
        flag, pf = tri.triangulate_mono(obs, poses)
    if flag and not np.isnan(pf).any():
        Ra, pa = feat.anchor.value_linear(), feat.anchor.value_trans()
        if (Ra.T @ (pf - pa))[2] <= 0:
            return False
        feat.pf_w = np.array(pf, dtype=np.float64)
        feat.is_tri = True
        feat.tri_ok = True
        return True
    return False


def select_lost(map_server, tri, state, stereo=False):
    """RemoveLostUpdate.cpp:45-59 (mono: >= 4 frames) / :280-294 (stereo: >= 3): returns update_ids after erasing the
    directly marginalised tracks. Call after mark_marg_features."""
    min_obs = 3 if stereo else 4
    update_ids, direct = [], []
    for key in map_server.ids():
        f = map_server[key]
        if f.ftype == MSCKF and f.is_to_marg:
            if triangulate_feature_info(f, tri, state, stereo) and len(_obs_of(f, stereo)) >= min_obs:
                update_ids.append(key)
            else:
                direct.append(key)
    for key in direct:
        del map_server[key]
    return update_ids


def select_seen_at(map_server, tri, state, selected_timestamps, stereo=False):
    """SwMargUpdate.cpp:61-85 / KeyframeUpdate.cpp:455-480: MSCKF tracks observed at every selected clone that
    triangulate."""
    update_ids = []
    for key in map_server.ids():
        f = map_server[key]
        if f.ftype != MSCKF:
            continue
        obs = _obs_of(f, stereo)
        if any(ts not in obs for ts in selected_timestamps):
            continue
        if triangulate_feature_info(f, tri, state, stereo):
            update_ids.append(key)
    return update_ids


def clean_obs_at(map_server, marg_times, stereo=False):
    """SwMargUpdate::clean{Mono,Stereo}ObsAtMargTime (SwMargUpdate.cpp:191-213, :389-411) and the keyframe twin
    (KeyframeUpdate.cpp:251-278): drop the observations at the clones about to leave, erase tracks left empty."""
    to_clean = []
    for key in map_server.ids():
        obs = _obs_of(map_server[key], stereo)
        for t in marg_times:
            obs.pop(t, None)
            if len(obs) == 0:
                to_clean.append(key)
    for key in to_clean:
        map_server.pop(key, None)


def change_msckf_anchor(map_server, state, marg_times, min_depth):
    """SwMargUpdate::changeMSCKFAnchor (SwMargUpdate.cpp:216-259, min_depth 0) / KeyframeUpdate::changeMSCKFAnchor
    (KeyframeUpdate.cpp:280-327, min_depth 0.3): tracks anchored at a clone about to leave move to the newest clone."""
    old = [state.sw_camleft_poses[t] for t in marg_times if t in state.sw_camleft_poses]
    if not old:
        return
    new_anchor = state.sw_camleft_poses[max(state.sw_camleft_poses.keys())]
    to_marg = []
    for key in map_server.ids():
        f = map_server[key]
        if f.ftype != MSCKF:
            continue
        if any(f.anchor is o for o in old):
            if f.is_tri:
                body = new_anchor.value_linear().T @ (f.pf_w - new_anchor.value_trans())
                if body[2] <= min_depth:
                    to_marg.append(key)
                    continue
                f.anchor = new_anchor
            else:
                to_marg.append(key)
    for key in to_marg:
        del map_server[key]


def erase_invalid_features(map_server, min_depth=0.2, state=None):
    """MapServerManager::eraseInvalidFeatures (:454-490); `state` is needed once SLAM features exist (:485-486)."""
    from .landmark_update import marg_anchored_landmark_in_state, sync_feat
    rm = []
    for key in map_server.ids():
        f = map_server[key]
        if not f.is_tri:
            continue
        sync_feat(f)
        if f.anchor is None:
            rm.append(key)
            continue
        body = f.anchor.value_linear().T @ (f.pf_w - f.anchor.value_trans())
        if body[2] <= min_depth:
            rm.append(key)
    for key in rm:
        if map_server[key].ftype == SLAM and state is not None:
            marg_anchored_landmark_in_state(state, key)
        del map_server[key]


def table_snapshot(map_server, state, stereo=False, obs_slots=None):
    """Canonical dump (ascending id) in the array layout of igv_tracks_get, for bit-exact comparisons."""
    times = state.sw_times()
    SW = obs_slots if obs_slots is not None else len(times)
    rho = 4 if stereo else 2
    keys = map_server.ids()
    n = len(keys)
    out = dict(id=np.array(keys, dtype=np.int32), to_marg=np.zeros(n, np.uint8), is_tri=np.zeros(n, np.uint8),
               mask=np.zeros((n, SW), np.uint8), obs=np.zeros((n, SW, rho)), anchor_slot=np.full(n, -1, np.int32),
               pf=np.zeros((n, 3)))
    for i, k in enumerate(keys):
        f = map_server[k]
        out["to_marg"][i] = 1 if f.is_to_marg else 0
        out["is_tri"][i] = 1 if f.is_tri else 0
        if f.is_tri:
            out["pf"][i] = f.pf_w
        ob = _obs_of(f, stereo)
        for s, t in enumerate(times):
            if t in ob:
                out["mask"][i, s] = 1
                out["obs"][i, s] = ob[t]
            if f.anchor is state.sw_camleft_poses[t]:
                out["anchor_slot"][i] = s
    return out
