"""SLAM-landmark path, mono: landmarks kept in the state as 3-dim anchored variables.

TEST INFRASTRUCTURE (oracle). CPU restatement of
  /root/reference/ingvio_estimator/src/LandmarkUpdate.cpp:32-149     updateLandmarkMono
  /root/reference/ingvio_estimator/src/LandmarkUpdate.cpp:273-361    changeLandmarkAnchor (both overloads)
  /root/reference/ingvio_estimator/src/LandmarkUpdate.cpp:363-423    initNewLandmarkMono
  /root/reference/ingvio_estimator/src/LandmarkUpdate.cpp:426-500    calcResJacobianSingleFeatAllMonoObs
  /root/reference/ingvio_estimator/src/LandmarkUpdate.cpp:521-572    calcResJacobianSingleLandmarkMono (epose / ext form)
  /root/reference/ingvio_estimator/src/MapServerManager.cpp:343-387  FeatureInfoManager::changeAnchoredPose
  /root/reference/ingvio_estimator/src/StateManager.cpp:340-353      margAnchoredLandmarkInState
  /root/reference/ingvio_estimator/src/AnchoredLandmark.cpp:227-243  AnchoredLandmark::update (types.AnchoredLandmark)
Pinned by the reference itself: tests/test_ref_pin.py drives the unmodified LandmarkUpdate.cpp (oracle/_ref) and this file
over the same recorded frames with max_lm_feats > 0.  A SLAM feature keeps its value in `feat.landmark` (an
AnchoredLandmark in the state); `feat.pf_w` / `feat.anchor` mirror it for the map-server functions that read them.
"""
import numpy as np

from .lie import skew
from .state_manager import StateManager
from .types import AnchoredLandmark
from .visual_update import MSCKF, SLAM, UpdateBase, _h_proj


def lm_pf(feat):
    return feat.landmark.value_pos_xyz() if getattr(feat, "landmark", None) is not None else feat.pf_w


def sync_feat(feat):
    """Keep the MSCKF-style fields of a SLAM feature in step with its landmark variable."""
    if getattr(feat, "landmark", None) is not None:
        feat.pf_w = feat.landmark.value_pos_xyz().copy()
        feat.anchor = feat.landmark.get_anchored_pose()


def marg_anchored_landmark_in_state(state, lm_id):
    """StateManager.cpp:340-353."""
    if lm_id not in state.anchored_landmarks:
        return
    StateManager.marginalize(state, state.anchored_landmarks[lm_id])
    del state.anchored_landmarks[lm_id]


def change_anchored_pose(feat, state, target_ts):
    """FeatureInfoManager::changeAnchoredPose (MapServerManager.cpp:343-378)."""
    if len(state.sw_camleft_poses) < 2:
        return
    if target_ts not in state.sw_camleft_poses or feat.id not in state.anchored_landmarks:
        return
    if feat.ftype != SLAM:
        return
    lm = feat.landmark
    if not any(p is lm.get_anchored_pose() for p in state.sw_camleft_poses.values()):
        return
    if state.anchored_landmarks[feat.id] is not lm:
        return
    target = state.sw_camleft_poses[target_ts]
    var_order = [lm.get_anchored_pose(), target, lm]
    pf = lm.value_pos_xyz()
    H = np.zeros((3, 15))
    H[:, 0:3] = -skew(pf)
    H[:, 6:9] = skew(pf)
    H[:, 12:15] = np.eye(3)
    StateManager.replace_var_linear(state, lm, var_order, H)
    lm.reset_anchored_pose(target)          # the world position is unchanged (resetAnchoredPose(..., true))
    sync_feat(feat)


class LandmarkUpdate(UpdateBase):
    """LandmarkUpdate.h:36-127 (mono)."""

    def __init__(self, fp):
        super().__init__(fp.chi2_max_dof, fp.chi2_thres)
        self.noise = fp.visual_noise
        self.last_init = []     # (id, accepted) of the last init_new_landmark_mono, for parity tests

    # ---- per-landmark Jacobian at the current IMU pose (LandmarkUpdate.cpp:521-572) ----
    @staticmethod
    def calc_res_jacobian_single_landmark_mono(feat, state):
        pf_w = feat.landmark.value_pos_xyz()
        e, x = state.extended_pose, state.camleft_imu_extrinsics
        R_i2w_T, R_cl2i_T = e.rot.T, x.rot.T
        pf_i = R_i2w_T @ (pf_w - e.vec1)
        pf_cl = R_cl2i_T @ (pf_i - x.vec)
        if state.timestamp not in feat.mono_obs or feat.ftype != SLAM:
            raise RuntimeError("[LandmarkUpdate]: Cannot calc curr slam feature mono res and jacobi!")
        res = feat.mono_obs[state.timestamp] - np.array([pf_cl[0] / pf_cl[2], pf_cl[1] / pf_cl[2]])
        Hp = _h_proj(pf_cl)
        R_w2cl = R_cl2i_T @ R_i2w_T
        H_epose = np.zeros((2, 9))
        H_epose[:, 0:3] = Hp @ R_w2cl @ skew(pf_w)
        H_epose[:, 3:6] = -Hp @ R_w2cl
        H_ext = np.zeros((2, 6))
        H_ext[:, 0:3] = Hp @ R_cl2i_T @ skew(pf_i)
        H_ext[:, 3:6] = -Hp @ R_cl2i_T
        H_anch = np.zeros((2, 6))
        H_anch[:, 0:3] = -Hp @ R_w2cl @ skew(pf_w)
        H_pf = Hp @ R_w2cl
        return res, H_epose, H_ext, H_anch, H_pf

    def update_landmark_mono(self, state, map_server):
        """LandmarkUpdate.cpp:32-149. `_anchored_landmarks` is an unordered_map in the reference: the stacking order is
        unspecified there and irrelevant to the posterior; ascending id here."""
        if len(state.anchored_landmarks) == 0:
            return None
        var_order = [state.extended_pose, state.camleft_imu_extrinsics]
        col_of = {id(state.extended_pose): 0, id(state.camleft_imu_extrinsics): 9}
        rows, col_cnt = [], 15
        self.last_gammas = []
        for lid in sorted(state.anchored_landmarks):
            if lid not in map_server:
                raise RuntimeError("[LandmarkUpdate]: Landmark in state not in map server!")
            feat = map_server[lid]
            if feat.ftype != SLAM:
                raise RuntimeError("[LandmarkUpdate]: Landmark in state not marked SLAM type in map server!")
            if state.timestamp not in feat.mono_obs:
                raise RuntimeError("[LandmarkUpdate]: Landmark in state not tracked to curr time!")
            lm = state.anchored_landmarks[lid]
            anch = lm.get_anchored_pose()
            res, H_epose, H_ext, H_anch, H_pf = self.calc_res_jacobian_single_landmark_mono(feat, state)
            H_chi2 = np.hstack([H_epose, H_ext, H_anch, H_pf])
            if not self.test_chi_squared(state, res, H_chi2, [state.extended_pose, state.camleft_imu_extrinsics, anch, lm],
                                         self.noise, fid=lid):
                continue
            if id(anch) not in col_of:
                col_of[id(anch)] = col_cnt
                col_cnt += 6
                var_order.append(anch)
            if id(lm) not in col_of:
                col_of[id(lm)] = col_cnt
                col_cnt += 3
                var_order.append(lm)
            rows.append((res, H_epose, H_ext, H_anch, H_pf, col_of[id(anch)], col_of[id(lm)]))
        if not rows:
            return None
        H = np.zeros((2 * len(rows), col_cnt))
        r = np.zeros(2 * len(rows))
        for k, (res, He, Hx, Ha, Hf, ca, cl) in enumerate(rows):
            r[2 * k:2 * k + 2] = res
            H[2 * k:2 * k + 2, 0:9] = He
            H[2 * k:2 * k + 2, 9:15] = Hx
            H[2 * k:2 * k + 2, ca:ca + 6] = Ha
            H[2 * k:2 * k + 2, cl:cl + 3] = Hf
        dx = StateManager.ekf_update(state, var_order, H, r, self.noise ** 2.0 * np.eye(r.shape[0]), return_dx=True)
        for lid in state.anchored_landmarks:
            if lid in map_server:
                sync_feat(map_server[lid])
        return dx

    # ---- delayed initialisation rows over the whole window (LandmarkUpdate.cpp:426-500) ----
    @staticmethod
    def calc_res_jacobian_single_feat_all_mono_obs(feat, state):
        times = state.sw_times()
        sw_index = {id(state.sw_camleft_poses[t]): 6 * k for k, t in enumerate(times)}
        ncols = 6 * len(times)
        pf_w = lm_pf(feat)
        anchor = feat.landmark.get_anchored_pose() if getattr(feat, "landmark", None) is not None else feat.anchor
        res, Hx, Hf = [], [], []
        for t in sorted(feat.mono_obs):
            if t not in state.sw_camleft_poses:
                continue
            pose = state.sw_camleft_poses[t]
            R, p = pose.value_linear(), pose.value_trans()
            pf_cm = R.T @ (pf_w - p)
            Hp = _h_proj(pf_cm)
            H_pf2x = np.zeros((3, ncols))
            c = sw_index[id(pose)]
            if pose is not anchor:
                H_pf2x[:, c:c + 3] = R.T @ skew(pf_w)
                ca = sw_index[id(anchor)]
                H_pf2x[:, ca:ca + 3] = -H_pf2x[:, c:c + 3]
            H_pf2x[:, c + 3:c + 6] = -R.T
            if np.isnan(Hp).any() or np.isnan(H_pf2x).any():
                continue
            Hx.append(Hp @ H_pf2x)
            Hf.append(Hp @ R.T)
            res.append(feat.mono_obs[t] - np.array([pf_cm[0] / pf_cm[2], pf_cm[1] / pf_cm[2]]))
        if not res:
            return np.zeros(0), np.zeros((0, ncols)), np.zeros((0, 3))
        return np.concatenate(res), np.vstack(Hx), np.vstack(Hf)

    def init_new_landmark_mono(self, state, map_server, triangulate, min_init_poses):
        """LandmarkUpdate.cpp:363-423. `triangulate(feat) -> bool` is FeatureInfoManager::triangulateFeatureInfoMono."""
        self.last_init = []
        if len(state.sw_camleft_poses) < min_init_poses:
            return
        vac = state.state_params.max_landmarks - len(state.anchored_landmarks)
        if vac <= 0:
            return
        ids = []
        for key in sorted(map_server):
            if len(ids) >= vac:
                break
            f = map_server[key]
            if len(f.mono_obs) < min_init_poses or f.ftype == SLAM:
                continue
            if not triangulate(f):
                continue
            ids.append(key)
        sw_var_type = [state.sw_camleft_poses[t] for t in state.sw_times()]
        for key in ids:
            f = map_server[key]
            lm = AnchoredLandmark()
            lm.reset_anchored_pose(f.anchor)
            lm.set_value_pos_xyz(f.pf_w)
            f.landmark = lm
            res, Hx, Hf = self.calc_res_jacobian_single_feat_all_mono_obs(f, state)
            ok = StateManager.add_variable_delayed(state, lm, sw_var_type, Hx, Hf, res, self.noise, 0.95, True)
            self.last_init.append((key, bool(ok)))
            if not ok:
                f.landmark = None
                continue
            if key in state.anchored_landmarks:
                raise RuntimeError("[LandmarkUpdate]: The id intended to add already in state!")
            state.anchored_landmarks[key] = lm
            f.ftype = SLAM
            sync_feat(f)
        for lid in state.anchored_landmarks:       # the residual EKF of every accepted initialisation moved them all
            if lid in map_server:
                sync_feat(map_server[lid])

    # ---- anchor change before the old clone(s) leave (LandmarkUpdate.cpp:273-361) ----
    @staticmethod
    def change_landmark_anchor(state, map_server, marg_kfs=None):
        if marg_kfs is None:
            mt = state.next_marg_time()
            if mt == float("inf") or mt not in state.sw_camleft_poses:
                return
            marg_kfs = [mt]
        if len(marg_kfs) == 0:
            return
        old = [state.sw_camleft_poses[t] for t in marg_kfs]
        latest = max(state.sw_camleft_poses.keys())
        new_anchor = state.sw_camleft_poses[latest]
        to_marg = []
        for key in sorted(map_server):
            f = map_server[key]
            if f.ftype != SLAM:
                continue
            if any(f.landmark.get_anchored_pose() is o for o in old):
                pf = f.landmark.value_pos_xyz()
                body = new_anchor.value_linear().T @ (pf - new_anchor.value_trans())
                if body[2] <= 0:
                    to_marg.append(key)
                    continue
                change_anchored_pose(f, state, latest)
                if f.landmark.get_anchored_pose() is not new_anchor:
                    to_marg.append(key)
        for key in to_marg:
            marg_anchored_landmark_in_state(state, key)
            del map_server[key]
