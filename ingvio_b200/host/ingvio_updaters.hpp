// ingvio_updaters.hpp -- C++ host mirror of the reference's visual updaters on top of the fused C-ABI calls.
//
// Reference: RemoveLostUpdate.h:32-77 / .cpp:40-167,276-405; SwMargUpdate.h:33-110 / .cpp:42-259,389-497;
// KeyframeUpdate.h:33-130 / .cpp:43-129,251-327,438-735; Triangulator.h:35-60. Same class names, constructors
// (from IngvioParams), method names and argument lists as the reference's Updater interfaces, so IngvioFilter's
// frame callbacks (IngvioFilter.cpp:143-205) compile unchanged against them. Each updateState* body is the device
// chain  igv_tracks_gather -> igv_triangulate -> igv_tracks_commit_tri -> igv_msckf_update [-> igv_tracks_erase]
// instead of the per-feature Eigen / SuiteSparse loop; the host Type objects are refreshed from the handle's mean
// mirror afterwards (StateManager::sync_mean_from_device). `Triangulator` only carries its parameters: the solve runs
// inside the chain (igv_triangulate).
//
// Header-only, plain C++17.
#pragma once
#include "ingvio_map_server.hpp"

namespace ingvio {

// The subset of IngvioParams (IngvioParams.h) these classes read; defaults = config/sportsfield/ingvio_mono.yaml.
struct IngvioParams {
  int _chi2_max_dof = 150;
  double _chi2_thres = 0.95;
  double _visual_noise = 0.12;
  int _frame_select_interval = 28;
  int _max_sw_clones = 20;
  double _trans_thres = 0.1, _huber_epsilon = 0.01, _conv_precision = 5e-7, _init_damping = 1e-3;
  int _outer_loop_max_iter = 10, _inner_loop_max_iter = 10;
  double _max_depth = 60.0, _min_depth = 0.2;
  // GNSS (IngvioParams.h:86-113)
  double _psr_noise_amp = 1.0, _dopp_noise_amp = 1.0, _init_cov_yof = 0.015;
  int _is_adjust_yof = 0, _is_gnss_chi2_test = 0, _is_gnss_strong_reject = 0;
};

class Triangulator {   // Triangulator.h:35-60
 public:
  Triangulator() {}
  explicit Triangulator(const IngvioParams& f) {
    _prm.trans_thres = f._trans_thres; _prm.huber_epsilon = f._huber_epsilon; _prm.conv_precision = f._conv_precision;
    _prm.init_damping = f._init_damping; _prm.outer_loop_max_iter = f._outer_loop_max_iter;
    _prm.inner_loop_max_iter = f._inner_loop_max_iter; _prm.max_depth = f._max_depth; _prm.min_depth = f._min_depth;
  }
  const igv_tri_params& params() const { return _prm; }
 private:
  igv_tri_params _prm{0.1, 0.01, 5e-7, 1e-3, 10, 10, 60.0, 0.2};   // Triangulator.h:38-46 defaults
};

namespace detail {

inline std::vector<int> slots_of(const std::shared_ptr<State>& state, const std::vector<double>& times) {
  std::vector<int> slots;
  for (double t : times) {
    int s = 0;
    bool found = false;
    for (auto& it : state->_sw_camleft_poses) { if (it.first == t) { found = true; break; } ++s; }
    if (!found) { std::printf("[SwMargUpdate]: selected timestamp not in sw!\n"); std::exit(EXIT_FAILURE); }   // SwMargUpdate.cpp:424-428
    slots.push_back(s);
  }
  return slots;
}

// One fused visual update over the tracks the table selects. Returns the number of selected tracks.
inline int fused_visual_update(const UpdateBase& upd, std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server,
                               const std::shared_ptr<Triangulator>& tri, int rule, const std::vector<int>& sel_slots, int min_obs,
                               int dof_fixed, int vis_mode, double noise, int max_valid, bool erase_after) {
  igv_batch* h = StateManager::handle(state);
  map_server->bind(state);
  const int F = state->_max_feats, SW = state->_state_params._max_sw_poses + 1;
  const int rho = state->_state_params._cam_nums == 2 ? 4 : 2;
  std::vector<int> entry(F), anchor(F), dof(F);
  int n_sel = 0;
  std::vector<double> obs((std::size_t)F * SW * rho), pf(3 * (std::size_t)F);
  std::vector<unsigned char> mask_all((std::size_t)F * SW), mask_upd((std::size_t)F * SW), ok(F), tri_ok(F);
  igv_track_gather_args g{};
  g.rule = rule; g.n_selected = (int)sel_slots.size(); g.selected_slots = sel_slots.empty() ? nullptr : sel_slots.data();
  g.min_obs = min_obs; g.dof_fixed = dof_fixed; g.n_feats = F; g.obs_slots = SW;
  g.track_entry = entry.data(); g.n_sel = &n_sel; g.obs = obs.data(); g.mask_all = mask_all.data(); g.mask_upd = mask_upd.data();
  g.anchor_slot = anchor.data(); g.chi2_dof = dof.data(); g.feat_ok = ok.data();
  StateManager::check(state, igv_tracks_gather(h, &g), true);
  if (n_sel == 0) return 0;                       // "if (update_ids.size() == 0) return;" (after the erase of direct_marg_ids: none here)
  igv_tri_args t{};
  t.n_feats = F; t.obs = obs.data(); t.obs_mask = mask_all.data(); t.obs_slots = SW; t.anchor_slot = anchor.data();
  t.prm = (tri ? tri : std::make_shared<Triangulator>())->params();
  t.pf_out = pf.data(); t.ok_out = tri_ok.data();
  StateManager::check(state, igv_triangulate(h, &t), true);
  StateManager::check(state, igv_tracks_commit_tri(h, F, entry.data(), pf.data(), tri_ok.data(), ok.data()), true);
  upd.upload(state);                              // this updater's chi^2 table (Update.cpp:27-34)
  igv_msckf_args a{};
  a.mode = vis_mode; a.n_feats = F; a.pf_w = pf.data(); a.anchor_slot = anchor.data(); a.obs = obs.data();
  a.obs_mask = mask_upd.data(); a.chi2_dof = dof.data(); a.obs_slots = SW; a.noise = noise; a.max_valid = max_valid;
  a.feat_ok = ok.data();
  StateManager::check(state, igv_msckf_update(h, &a), true);
  if (erase_after) StateManager::check(state, igv_tracks_erase(h, F, entry.data()), true);
  StateManager::sync_mean_from_device(state);     // every Type::update(dx) of StateManager::ekfUpdate (:425)
  map_server->touch();
  return n_sel;
}

}  // namespace detail

// ---- RemoveLostUpdate (RemoveLostUpdate.h:32-77) --------------------------------------------------------------------
class RemoveLostUpdate : public UpdateBase {
 public:
  explicit RemoveLostUpdate(const IngvioParams& f) : UpdateBase(f._chi2_max_dof, f._chi2_thres), _max_valid_ids(20), _noise(f._visual_noise) {}
  void updateStateMono(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) {
    MapServerManager::markMargMonoFeatures(map_server, state);                                      // RemoveLostUpdate.cpp:44
    detail::fused_visual_update(*this, state, map_server, tri, IGV_TRK_LOST, {}, 4, 0, IGV_VIS_ALL_OBS, _noise, _max_valid_ids, true);
  }
  void updateStateStereo(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) {
    MapServerManager::markMargStereoFeatures(map_server, state);                                    // :280
    detail::fused_visual_update(*this, state, map_server, tri, IGV_TRK_LOST, {}, 3, 0, IGV_VIS_ALL_OBS, _noise, _max_valid_ids, true);
  }
 protected:
  int _max_valid_ids;
  double _noise;
};

// ---- SwMargUpdate (SwMargUpdate.h:33-110) ---------------------------------------------------------------------------
class SwMargUpdate : public UpdateBase {
 public:
  explicit SwMargUpdate(const IngvioParams& f)
      : UpdateBase(f._chi2_max_dof, f._chi2_thres), _noise(f._visual_noise), _frame_select_interval(f._frame_select_interval) {}
  void updateStateMono(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) { update(state, map_server, tri); }
  void updateStateStereo(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) { update(state, map_server, tri); }
  void cleanMonoObsAtMargTime(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) { clean(state, map_server); }     // :191-213
  void cleanStereoObsAtMargTime(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) { clean(state, map_server); }   // :389-411
  void changeMSCKFAnchor(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) {                                        // :216-259
    const double marg_time = state->nextMargTime();
    if (marg_time == std::numeric_limits<double>::infinity() || !state->_sw_camleft_poses.count(marg_time)) return;
    map_server->bind(state);
    const std::vector<int> s = detail::slots_of(state, {marg_time});
    StateManager::check(state, igv_tracks_change_anchor(StateManager::handle(state), 1, s.data(), 0.0));
    map_server->touch();
  }
  void margSwPose(std::shared_ptr<State> state) {                                                                                      // :261-268
    const double marg_time = state->nextMargTime();
    if (marg_time == std::numeric_limits<double>::infinity()) return;
    StateManager::margSlidingWindowPose(state, marg_time);
  }
  void selectSwTimestamps(const std::map<double, std::shared_ptr<SE3>>& sw_poses, const double& marg_time,
                          std::vector<double>& selected_timestamps) const {                                                           // :475-497
    selected_timestamps.clear();
    if (marg_time == std::numeric_limits<double>::infinity() || !sw_poses.count(marg_time)) return;
    int cnt = 1;
    selected_timestamps.push_back(marg_time);
    for (const auto& item : sw_poses) {
      if (item.first <= marg_time) continue;
      if (cnt % _frame_select_interval == 0) selected_timestamps.push_back(item.first);
      ++cnt;
    }
  }
 protected:
  double _noise;
  int _frame_select_interval;
  void update(std::shared_ptr<State>& state, std::shared_ptr<MapServer>& map_server, std::shared_ptr<Triangulator>& tri) {          // :42-189
    const double marg_time = state->nextMargTime();
    if (marg_time == std::numeric_limits<double>::infinity()) return;
    std::vector<double> sel;
    selectSwTimestamps(state->_sw_camleft_poses, marg_time, sel);
    detail::fused_visual_update(*this, state, map_server, tri, IGV_TRK_SEEN_AT, detail::slots_of(state, sel), 0, 0, IGV_VIS_SELECTED,
                                _noise, 0, false);                                                  // dof = #selected-1 (:129-130)
  }
  void clean(std::shared_ptr<State>& state, std::shared_ptr<MapServer>& map_server) {
    const double marg_time = state->nextMargTime();
    if (marg_time == std::numeric_limits<double>::infinity()) return;
    map_server->bind(state);
    const std::vector<int> s = detail::slots_of(state, {marg_time});
    StateManager::check(state, igv_tracks_clean_obs(StateManager::handle(state), 1, s.data()));
    map_server->touch();
  }
};

// ---- KeyframeUpdate (KeyframeUpdate.h:33-130) -----------------------------------------------------------------------
class KeyframeUpdate : public UpdateBase {
 public:
  explicit KeyframeUpdate(const IngvioParams& f) : UpdateBase(f._chi2_max_dof, f._chi2_thres), _noise(f._visual_noise), _max_sw_poses(f._max_sw_clones) {}
  void getMargKfs(const std::shared_ptr<State> state, std::vector<double>& marg_kfs) {                                               // :43-116
    if ((int)state->_sw_camleft_poses.size() < _max_sw_poses || _max_sw_poses < 3) { marg_kfs.clear(); return; }
    if (state->_timestamp == _timestamp && _kfs.size() > 0) { marg_kfs = _kfs; return; }
    if ((int)state->_sw_camleft_poses.size() > _max_sw_poses) {
      std::printf("[KeyframeUpdate]: Current sw poses larger than max size!\n");
      std::exit(EXIT_FAILURE);
    }
    _timestamp = state->_timestamp;
    _kfs.clear();
    const int rem = _max_sw_poses - 2;
    const int idx1 = 2 + _select_cnt;
    _select_cnt = (_select_cnt + 1) % rem;
    auto item1 = state->_sw_camleft_poses.rbegin();
    for (int i = 0; i < idx1; ++i) ++item1;
    auto item2 = state->_sw_camleft_poses.rbegin();
    ++item2;
    _kfs.push_back(item1->first);
    _kfs.push_back(item2->first);
    marg_kfs = _kfs;
  }
  void margSwPose(std::shared_ptr<State> state) {                                                                                      // :118-129
    std::vector<double> kfs;
    getMargKfs(state, kfs);
    for (const double& t : kfs) StateManager::margSlidingWindowPose(state, t);
  }
  void changeMSCKFAnchor(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) {                                        // :280-327
    std::vector<double> kfs;
    getMargKfs(state, kfs);
    if (kfs.empty()) return;
    map_server->bind(state);
    const std::vector<int> s = detail::slots_of(state, kfs);
    StateManager::check(state, igv_tracks_change_anchor(StateManager::handle(state), (int)s.size(), s.data(), 0.3));
    map_server->touch();
  }
  void updateStateMono(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) { update(state, map_server, tri); }
  void updateStateStereo(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server, std::shared_ptr<Triangulator> tri) { update(state, map_server, tri); }
  void cleanMonoObsAtMargTime(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) { clean(state, map_server); }     // :251-278
  void cleanStereoObsAtMargTime(std::shared_ptr<State> state, std::shared_ptr<MapServer> map_server) { clean(state, map_server); }
 protected:
  double _noise;
  int _max_sw_poses;
  double _timestamp = -1.0;
  std::vector<double> _kfs;
  static inline int _select_cnt = 0;              // a class static in the reference (KeyframeUpdate.cpp:41)
  void update(std::shared_ptr<State>& state, std::shared_ptr<MapServer>& map_server, std::shared_ptr<Triangulator>& tri) {          // :438-735
    std::vector<double> sel;
    getMargKfs(state, sel);
    if (sel.empty()) return;
    detail::fused_visual_update(*this, state, map_server, tri, IGV_TRK_SEEN_AT, detail::slots_of(state, sel), 0, 2, IGV_VIS_SELECTED,
                                _noise, 0, false);                                                  // dof = 2 (:525-526)
  }
  void clean(std::shared_ptr<State>& state, std::shared_ptr<MapServer>& map_server) {
    std::vector<double> kfs;
    getMargKfs(state, kfs);
    if (kfs.empty()) return;
    map_server->bind(state);
    const std::vector<int> s = detail::slots_of(state, kfs);
    StateManager::check(state, igv_tracks_clean_obs(StateManager::handle(state), (int)s.size(), s.data()));
    map_server->touch();
  }
};

}  // namespace ingvio
