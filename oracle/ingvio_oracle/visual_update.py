"""MSCKF visual updaters: chi^2 gate, RemoveLost / Keyframe / SwMarg updates (mono + stereo).

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/Update.cpp:27-149,
RemoveLostUpdate.cpp:40-523, KeyframeUpdate.cpp:43-735, SwMargUpdate.cpp:42-700.

Third-party arithmetic that is NOT under /root/reference, restated from the
published algorithms and unpinned by any reference test ("parity unpinned" at the reference level):
  * Eigen::JacobiSVD full-U left null space  -> numpy.linalg.svd(full_matrices=True) (LAPACK gesdd)
  * SuiteSparse SPQR, SPQR_ORDERING_NATURAL   -> numpy.linalg.qr(mode="complete") (LAPACK geqrf)
  * boost::math::quantile(chi_squared)        -> scipy.stats.chi2.ppf
Only basis-invariant outputs (gate value gamma, dx, P, state) are comparable across implementations.

Triangulation (Triangulator.cpp) is outside the path: every FeatureInfo carries the
triangulated world position `pf_w` and a `tri_ok` flag supplied by the caller.
"""
import numpy as np
from scipy.stats import chi2 as _chi2

from .lie import skew
from .state import State
from .state_manager import StateManager

MSCKF, SLAM = 0, 1


class FeatureInfo:
    """MapServer.h:69-132, reduced to what the updaters read."""

    def __init__(self, fid, pf_w, anchor, tri_ok=True):
        self.id = fid
        self.ftype = MSCKF
        self.is_to_marg = False
        self.is_tri = tri_ok
        self.tri_ok = tri_ok
        self.pf_w = np.array(pf_w, dtype=np.float64).reshape(3)
        self.anchor = anchor          # SE3 object (identity = `is`)
        self.mono_obs = {}            # timestamp -> (2,)
        self.stereo_obs = {}          # timestamp -> (4,)

    def num_of_mono_frames(self):
        return len(self.mono_obs)

    def num_of_stereo_frames(self):
        return len(self.stereo_obs)


class UpdateBase:
    """Update.h:36-97 / Update.cpp:27-149."""

    def __init__(self, max_dof=150, thres=0.95):
        self.thres = thres
        self.chi_squared_table = {i: float(_chi2.ppf(thres, i)) for i in range(1, max_dof + 1)}
        self.last_gammas = []  # (feature id, gamma, dof, accepted) for parity tests

    def whiten_residual(self, state: State, res, H, var_order, noise_or_R):
        small_cov = StateManager.get_marginal_cov(state, var_order)
        if np.isscalar(noise_or_R):
            S = H @ small_cov @ H.T + noise_or_R ** 2.0 * np.eye(H.shape[0])
        else:
            S = H @ small_cov @ H.T + noise_or_R
        # Eigen ldlt().solve -> any SPD solve; S is SPD here
        return float(res @ np.linalg.solve(S, res))

    def _table(self, dof):
        if dof not in self.chi_squared_table:
            for i in range(max(self.chi_squared_table) + 1, dof + 1):
                self.chi_squared_table[i] = float(_chi2.ppf(self.thres, i))
        return self.chi_squared_table[dof]

    def test_chi_squared(self, state, res, H, var_order, noise_or_R, dof=None, fid=None):
        if dof is None:
            dof = res.shape[0]
        elif not np.isscalar(noise_or_R) and dof <= 0:
            return False
        prob = self.whiten_residual(state, res, H, var_order, noise_or_R)
        ok = prob < self._table(dof)
        self.last_gammas.append((fid, prob, dof, bool(ok)))
        return ok


def _h_proj(p):
    H = np.zeros((2, 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        H[0, 0] = 1.0 / p[2]
        H[0, 2] = -p[0] / p[2] ** 2
        H[1, 1] = 1.0 / p[2]
        H[1, 2] = -p[1] / p[2] ** 2
    return H


def _null_project(Hf, H, res, row_cnt):
    U, _, _ = np.linalg.svd(Hf, full_matrices=True)
    V = U[:, U.shape[1] - (row_cnt - 3):]
    return V.T @ H, V.T @ res, V


def _spqr_thin(H_large, res_large, keep_rows):
    """RemoveLostUpdate.cpp:139-160 etc.: Q^T [H, r] with natural ordering, keep the top rows.

    Only Q^T H and Q^T r are consumed by the reference, never Q itself, so the economy factor of
    [H | r] carries everything: its first n columns are Q^T H (upper triangular, zero below row n) and
    its last column is Q^T r with the whole tail of the residual folded into entry n. Keeping more than
    n rows (RemoveLost keeps all of them, :153) only appends rows whose H part is zero, which change
    neither the posterior nor dx; they are returned as that single folded row. This avoids the m x m Q
    that `mode="complete"` would allocate (4 GB at config c5)."""
    m, n = H_large.shape
    if m > n:
        R = np.linalg.qr(np.hstack([H_large, res_large.reshape(-1, 1)]), mode="r")
        rows = n if keep_rows <= n else min(n + 1, R.shape[0])
        return R[:rows, :n], R[:rows, n]
    return H_large, res_large


class RemoveLostUpdate(UpdateBase):
    """RemoveLostUpdate.h:33-75."""

    def __init__(self, fp, max_valid_ids=20):
        super().__init__(fp.chi2_max_dof, fp.chi2_thres)
        self.noise = fp.visual_noise
        self.max_valid_ids = max_valid_ids

    # -- per-feature residual/Jacobian over all observations in the window -------------------------
    def calc_res_jacobian_single_feat_all_obs(self, feat: FeatureInfo, sw_poses, stereo=False,
                                              T_cl2cr=None):
        """RemoveLostUpdate.cpp:169-273 (mono) / :407-523 (stereo)."""
        obs_map = feat.stereo_obs if stereo else feat.mono_obs
        rho = 4 if stereo else 2
        max_rows = rho * len(obs_map)
        max_cols = 6 * len(sw_poses)
        res_t = np.zeros(max_rows)
        H_t = np.zeros((max_rows, max_cols))
        Hf_t = np.zeros((max_rows, 3))
        pf_w = feat.pf_w
        anchor = feat.anchor
        block_index_map = {id(anchor): 0}
        block_var_order = [anchor]
        row_cnt, col_cnt = 0, 6
        for t in sorted(obs_map.keys()):
            if t not in sw_poses:
                continue
            z = obs_map[t]
            pose = sw_poses[t]
            R = pose.value_linear()
            p = pose.value_trans()
            pf_c = R.T @ (pf_w - p)
            Hp = _h_proj(pf_c)
            if stereo:
                Rc, pc = T_cl2cr
                pf_r = Rc @ pf_c + pc
                Hp_r = _h_proj(pf_r)
            flag = False
            if id(pose) not in block_index_map:
                flag = True
                block_index_map[id(pose)] = col_cnt
                block_var_order.append(pose)
                col_cnt += 6
            H_pf2x = np.zeros((3, max_cols))
            c0 = block_index_map[id(pose)]
            if pose is not anchor:
                H_pf2x[:, c0:c0 + 3] = R.T @ skew(pf_w)
                H_pf2x[:, 0:3] = -H_pf2x[:, c0:c0 + 3]
            H_pf2x[:, c0 + 3:c0 + 6] = -R.T
            H_pf2pf = R.T
            if np.isnan(Hp).any() or np.isnan(H_pf2x).any():
                if flag:
                    del block_index_map[id(pose)]
                    block_var_order.pop()
                    col_cnt -= 6
                continue
            H_t[row_cnt:row_cnt + 2, :] = Hp @ H_pf2x
            Hf_t[row_cnt:row_cnt + 2, :] = Hp @ H_pf2pf
            with np.errstate(divide="ignore", invalid="ignore"):
                if stereo:
                    H_t[row_cnt + 2:row_cnt + 4, :] = Hp_r @ Rc @ H_pf2x
                    Hf_t[row_cnt + 2:row_cnt + 4, :] = Hp_r @ Rc @ H_pf2pf
                    pred = np.array([pf_c[0] / pf_c[2], pf_c[1] / pf_c[2], pf_r[0] / pf_r[2], pf_r[1] / pf_r[2]])
                else:
                    pred = np.array([pf_c[0] / pf_c[2], pf_c[1] / pf_c[2]])
            res_t[row_cnt:row_cnt + rho] = np.asarray(z, dtype=np.float64) - pred
            row_cnt += rho
        res_t = res_t[:row_cnt]
        H_t = H_t[:row_cnt, :col_cnt]
        Hf_t = Hf_t[:row_cnt, :]
        H_block, res_block, _ = _null_project(Hf_t, H_t, res_t, row_cnt)
        return block_var_order, block_index_map, res_block, H_block

    def update_state(self, state: State, map_server: dict, stereo=False):
        """RemoveLostUpdate.cpp:40-167 (mono) / :276-405 (stereo)."""
        self.last_gammas = []
        t_now = state.timestamp
        for f in map_server.values():  # markMarg{Mono,Stereo}Features (MapServerManager.cpp:219-273)
            obs = f.stereo_obs if stereo else f.mono_obs
            if t_now not in obs:
                f.is_to_marg = True
        update_ids, direct = [], []
        for fid in sorted(map_server.keys()):
            f = map_server[fid]
            if f.ftype == MSCKF and f.is_to_marg:
                nobs = f.num_of_stereo_frames() if stereo else f.num_of_mono_frames()
                if f.tri_ok and nobs >= (3 if stereo else 4):
                    update_ids.append(fid)
                else:
                    direct.append(fid)
        for fid in direct:
            del map_server[fid]
        if not update_ids:
            return None
        dx = self.update_with_ids(state, map_server, update_ids, stereo)
        for fid in update_ids:
            del map_server[fid]
        return dx

    def update_with_ids(self, state: State, map_server: dict, update_ids, stereo=False, keep="rows"):
        """Body of RemoveLostUpdate.cpp:62-163 after the track selection. keep="rows" is the
        reference's topRows(row_cnt) (:153); keep="cols" is the Keyframe/SwMarg rule (same posterior)."""
        rho = 4 if stereo else 2
        sw_poses = state.sw_camleft_poses
        max_cols = 6 * len(sw_poses)
        max_rows = sum(rho * (map_server[i].num_of_stereo_frames() if stereo else
                              map_server[i].num_of_mono_frames()) - 3 for i in update_ids)
        res_large = np.zeros(max_rows)
        H_large = np.zeros((max_rows, max_cols))
        sw_index_map, sw_var_order = {}, []
        valid, row_cnt, col_cnt = 0, 0, 0
        T = (state.state_params.T_cl2cr_R, state.state_params.T_cl2cr_p)
        for fid in update_ids:
            f = map_server[fid]
            bvo, bim, res_b, H_b = self.calc_res_jacobian_single_feat_all_obs(f, sw_poses, stereo, T)
            dof = (len(f.stereo_obs) if stereo else len(f.mono_obs)) - 1
            if not self.test_chi_squared(state, res_b, H_b, bvo, self.noise, dof, fid=fid):
                continue
            rowblk = np.zeros((res_b.shape[0], max_cols))
            # std::map<shared_ptr,int> iterates in pointer order; any order gives the same posterior.
            for pose in bvo:
                sub = bim[id(pose)]
                if id(pose) not in sw_index_map:
                    sw_index_map[id(pose)] = col_cnt
                    sw_var_order.append(pose)
                    col_cnt += 6
                c = sw_index_map[id(pose)]
                rowblk[:, c:c + 6] = H_b[:, sub:sub + 6]
            H_large[row_cnt:row_cnt + H_b.shape[0], :] = rowblk
            res_large[row_cnt:row_cnt + res_b.shape[0]] = res_b
            row_cnt += res_b.shape[0]
            valid += 1
            if valid >= self.max_valid_ids:
                break
        H_large = H_large[:row_cnt, :col_cnt]
        res_large = res_large[:row_cnt]
        self.last_stack = (H_large.copy(), res_large.copy(), list(sw_var_order))
        H_thin, res_thin = _spqr_thin(H_large, res_large, row_cnt if keep == "rows" else col_cnt)
        dx = None
        if res_thin.shape[0] > 0:
            dx = StateManager.ekf_update(state, sw_var_order, H_thin, res_thin,
                                         self.noise ** 2 * np.eye(res_thin.shape[0]), return_dx=True)
        return dx


class _SelectedUpdateBase(UpdateBase):
    """Shared body of KeyframeUpdate / SwMargUpdate (the two files are textually parallel)."""

    def __init__(self, fp):
        super().__init__(fp.chi2_max_dof, fp.chi2_thres)
        self.noise = fp.visual_noise

    def calc_res_jacobian_single_feat_selected_obs(self, feat, sw_poses, sw_var_order, sw_index_map,
                                                   selected_timestamps, stereo=False, T_cl2cr=None):
        """KeyframeUpdate.cpp:160-249 / :330-436; SwMargUpdate.cpp:499-700."""
        rho = 4 if stereo else 2
        num_cols = 6 * len(sw_var_order)
        num_rows = rho * len(selected_timestamps)
        res_t = np.zeros(num_rows)
        H_t = np.zeros((num_rows, num_cols))
        Ha_t = np.zeros((num_rows, 6))
        Hf_t = np.zeros((num_rows, 3))
        pf_w = feat.pf_w
        anchor = feat.anchor
        obs_map = feat.stereo_obs if stereo else feat.mono_obs
        row_cnt = 0
        for t in selected_timestamps:
            if t not in sw_poses or t not in obs_map:
                continue
            z = obs_map[t]
            pose = sw_poses[t]
            R = pose.value_linear()
            p = pose.value_trans()
            pf_c = R.T @ (pf_w - p)
            Hp = _h_proj(pf_c)
            if stereo:
                Rc, pc = T_cl2cr
                pf_r = Rc @ pf_c + pc
                Hp_r = _h_proj(pf_r)
            H_pf2x = np.zeros((3, num_cols))
            H_pf2a = np.zeros((3, 6))
            c0 = sw_index_map[id(pose)]
            if pose is not anchor:
                H_pf2x[:, c0:c0 + 3] = R.T @ skew(pf_w)
                H_pf2a[:, 0:3] = -H_pf2x[:, c0:c0 + 3]
            H_pf2x[:, c0 + 3:c0 + 6] = -R.T
            H_pf2pf = R.T
            if np.isnan(Hp).any() or np.isnan(H_pf2x).any():
                continue
            H_t[row_cnt:row_cnt + 2, :] = Hp @ H_pf2x
            Ha_t[row_cnt:row_cnt + 2, :] = Hp @ H_pf2a
            Hf_t[row_cnt:row_cnt + 2, :] = Hp @ H_pf2pf
            if stereo:
                H_t[row_cnt + 2:row_cnt + 4, :] = Hp_r @ Rc @ H_pf2x
                Ha_t[row_cnt + 2:row_cnt + 4, :] = Hp_r @ Rc @ H_pf2a
                Hf_t[row_cnt + 2:row_cnt + 4, :] = Hp_r @ Rc @ H_pf2pf
                pred = np.array([pf_c[0] / pf_c[2], pf_c[1] / pf_c[2], pf_r[0] / pf_r[2], pf_r[1] / pf_r[2]])
            else:
                pred = np.array([pf_c[0] / pf_c[2], pf_c[1] / pf_c[2]])
            res_t[row_cnt:row_cnt + rho] = np.asarray(z, dtype=np.float64) - pred
            row_cnt += rho
        # `if (row_cnt < res_block.rows())` compares against the empty output argument
        # (KeyframeUpdate.cpp:230, SwMargUpdate.cpp:569): never true, so skipped observations
        # leave zero rows and nothing is shrunk.
        U, _, _ = np.linalg.svd(Hf_t, full_matrices=True)
        V = U[:, U.shape[1] - (row_cnt - 3):]
        return anchor, V.T @ res_t, V.T @ H_t, V.T @ Ha_t

    def _update_selected(self, state: State, map_server: dict, selected_timestamps, dof, stereo):
        """KeyframeUpdate.cpp:438-585 (mono) / :587-735 (stereo); SwMargUpdate.cpp:42-189 / :216-365."""
        self.last_gammas = []
        if len(selected_timestamps) == 0:
            return None
        sw_poses = state.sw_camleft_poses
        sw_var_order = [sw_poses[t] for t in selected_timestamps]
        sw_index_map = {id(sw_poses[t]): 6 * i for i, t in enumerate(selected_timestamps)}
        update_ids = []
        for fid in sorted(map_server.keys()):
            f = map_server[fid]
            if f.ftype != MSCKF:
                continue
            obs = f.stereo_obs if stereo else f.mono_obs
            if any(t not in obs for t in selected_timestamps):
                continue
            if f.tri_ok:
                update_ids.append(fid)
        if not update_ids:
            return None
        rho = 4 if stereo else 2
        max_cols = 6 * len(sw_poses)
        max_rows = len(update_ids) * (rho * len(selected_timestamps) - 3)
        res_large = np.zeros(max_rows)
        H_large = np.zeros((max_rows, max_cols))
        row_cnt = 0
        col_cnt = 6 * len(sw_var_order)
        T = (state.state_params.T_cl2cr_R, state.state_params.T_cl2cr_p)
        for fid in update_ids:
            f = map_server[fid]
            anchor, res_b, H_b, Ha_b = self.calc_res_jacobian_single_feat_selected_obs(
                f, sw_poses, sw_var_order, sw_index_map, selected_timestamps, stereo, T)
            flag = False
            if id(anchor) not in sw_index_map:
                flag = True
                sw_var_order.append(anchor)
                sw_index_map[id(anchor)] = col_cnt
                col_cnt += 6
                H_b = np.hstack([H_b, np.zeros((H_b.shape[0], 6))])
            ca = sw_index_map[id(anchor)]
            # NOTE: assignment, not accumulation (KeyframeUpdate.cpp:523, SwMargUpdate.cpp:127):
            # when the anchor is itself a selected clone its translation columns are overwritten by 0.
            if H_b.shape[1] < col_cnt:
                H_b = np.hstack([H_b, np.zeros((H_b.shape[0], col_cnt - H_b.shape[1]))])
            H_b[:, ca:ca + 6] = Ha_b
            if not self.test_chi_squared(state, res_b, H_b[:, :col_cnt], sw_var_order, self.noise,
                                         dof, fid=fid):
                if flag:
                    sw_var_order.pop()
                    del sw_index_map[id(anchor)]
                    col_cnt -= 6
                continue
            H_large[row_cnt:row_cnt + H_b.shape[0], :col_cnt] = H_b[:, :col_cnt]
            res_large[row_cnt:row_cnt + res_b.shape[0]] = res_b
            row_cnt += res_b.shape[0]
        H_large = H_large[:row_cnt, :col_cnt]
        res_large = res_large[:row_cnt]
        H_thin, res_thin = _spqr_thin(H_large, res_large, col_cnt)  # topRows(col_cnt)
        return StateManager.ekf_update(state, sw_var_order, H_thin, res_thin,
                                       self.noise ** 2 * np.eye(res_thin.shape[0]), return_dx=True)


class KeyframeUpdate(_SelectedUpdateBase):
    """KeyframeUpdate.h:33-130. `_select_cnt` is a class static in the reference (KeyframeUpdate.cpp:41)."""

    def __init__(self, fp):
        super().__init__(fp)
        self.max_sw_poses = fp.max_sw_clones
        self.select_cnt = 0
        self._timestamp = -1.0
        self._kfs = []

    def get_marg_kfs(self, state: State):
        """KeyframeUpdate.cpp:43-116."""
        n = len(state.sw_camleft_poses)
        if n < self.max_sw_poses or self.max_sw_poses < 3:
            return []
        if state.timestamp == self._timestamp and len(self._kfs) > 0:
            return list(self._kfs)
        if n > self.max_sw_poses:
            raise RuntimeError("[KeyframeUpdate]: Current sw poses larger than max size!")
        self._timestamp = state.timestamp
        rem = self.max_sw_poses - 2
        idx1 = 2 + self.select_cnt
        self.select_cnt = (self.select_cnt + 1) % rem
        times_desc = sorted(state.sw_camleft_poses.keys(), reverse=True)
        self._kfs = [times_desc[idx1], times_desc[1]]
        return list(self._kfs)

    def update_state(self, state, map_server, stereo=False):
        sel = self.get_marg_kfs(state)
        return self._update_selected(state, map_server, sel, 2, stereo)  # dof=2, KeyframeUpdate.cpp:525-526

    def marg_sw_pose(self, state):
        """KeyframeUpdate.cpp:118-129."""
        for t in self.get_marg_kfs(state):
            StateManager.marg_sliding_window_pose(state, t)


class SwMargUpdate(_SelectedUpdateBase):
    """SwMargUpdate.h:33-110."""

    def __init__(self, fp):
        super().__init__(fp)
        self.frame_select_interval = fp.frame_select_interval

    def select_sw_timestamps(self, sw_poses, marg_time):
        """SwMargUpdate.cpp:475-497."""
        if marg_time == float("inf") or marg_time not in sw_poses:
            return []
        sel = [marg_time]
        cnt = 1
        for t in sorted(sw_poses.keys()):
            if t <= marg_time:
                continue
            if cnt % self.frame_select_interval == 0:
                sel.append(t)
            cnt += 1
        return sel

    def update_state(self, state, map_server, stereo=False):
        marg_time = state.next_marg_time()
        if marg_time == float("inf"):
            return None
        sel = self.select_sw_timestamps(state.sw_camleft_poses, marg_time)
        return self._update_selected(state, map_server, sel, len(sel) - 1, stereo)  # SwMargUpdate.cpp:129-130

    def marg_sw_pose(self, state):
        StateManager.marg_sliding_window_pose(state)
