// ingvio_map_server.hpp -- C++ host mirror of the reference's map server on top of the C-ABI track table.
//
// Reference: ingvio_estimator/src/MapServer.h:31-134 (MonoMeas, StereoMeas, FeatureInfo,
// `typedef std::map<int, std::shared_ptr<FeatureInfo>> MapServer`), MapServerManager.h / .cpp:101-273, :454-490
// (collect{Mono,Stereo}Meas, markMarg{Mono,Stereo}Features, eraseInvalidFeatures) and the message layout
// feature_tracker/msg/{MonoMeas,StereoMeas,MonoFrame,StereoFrame}.msg. Same class and method names, argument meaning
// and error behaviour; the per-feature observation maps live in the handle's device table (igv_tracks_*), and
// `FeatureInfo` objects are read-only views refreshed from igv_tracks_get whenever the table changed.
// MSCKF-type features only (SLAM landmarks: SURVEY.md section 8f rank 3).
//
// Header-only, plain C++17; include after / instead of ingvio_host.hpp.
#pragma once
#include <array>
#include <cstdint>

#include "ingvio_host.hpp"

namespace feature_tracker {   // the wire format, field for field
struct Header { double stamp = 0.0; double toSec() const { return stamp; } };   // std_msgs/Header: header.stamp.toSec()
struct MonoMeas { std::uint64_t id = 0; double u0 = 0, v0 = 0; };
struct StereoMeas { std::uint64_t id = 0; double u0 = 0, v0 = 0, u1 = 0, v1 = 0; };
struct MonoFrame { Header header; std::vector<MonoMeas> mono_features; };
struct StereoFrame { Header header; std::vector<StereoMeas> stereo_features; };
}  // namespace feature_tracker

namespace ingvio {

class MapServer;
class MapServerManager;

// MapServer.h:69-132 -- a view of one table entry.
class FeatureInfo {
 public:
  enum FeatureType { MSCKF = 0, SLAM };
  const int& getId() const { return _id; }
  const FeatureType& getFeatureType() const { return _ftype; }
  const bool& isToMarg() const { return _isToMarg; }
  const bool& isTri() const { return _isTri; }
  bool hasMonoObsAt(double timestamp) const { return _mono_obs.count(timestamp) != 0; }
  bool hasStereoObsAt(double timestamp) const { return _stereo_obs.count(timestamp) != 0; }
  std::array<double, 2> monoMeasAt(double timestamp) const {       // zero when absent (MapServer.cpp:56-62)
    auto it = _mono_obs.find(timestamp);
    return it == _mono_obs.end() ? std::array<double, 2>{0, 0} : it->second;
  }
  std::array<double, 4> stereoMeasAt(double timestamp) const {     // MapServer.cpp:70-76
    auto it = _stereo_obs.find(timestamp);
    return it == _stereo_obs.end() ? std::array<double, 4>{0, 0, 0, 0} : it->second;
  }
  int numOfMonoFrames() const { return (int)_mono_obs.size(); }
  int numOfStereoFrames() const { return (int)_stereo_obs.size(); }
  const std::shared_ptr<SE3> anchor() const { return _anchor; }
  const Vec3d& valuePosXyz() const { return _pf; }                 // _landmark->valuePosXyz()

  int _id = -1;
  FeatureType _ftype = MSCKF;
  bool _isToMarg = false, _isTri = false;
  std::shared_ptr<SE3> _anchor;
  Vec3d _pf;
  std::map<double, std::array<double, 2>> _mono_obs;
  std::map<double, std::array<double, 4>> _stereo_obs;
};

// std::map<int, std::shared_ptr<FeatureInfo>> whose contents live on the device.
class MapServer {
 public:
  explicit MapServer(int max_tracks = 512) : _max_tracks(max_tracks) {}
  std::size_t size() { refresh(); return _view.size(); }
  bool count(int id) { refresh(); return _view.count(id) != 0; }
  std::shared_ptr<FeatureInfo> at(int id) { refresh(); return _view.at(id); }
  std::map<int, std::shared_ptr<FeatureInfo>>::const_iterator begin() { refresh(); return _view.begin(); }
  std::map<int, std::shared_ptr<FeatureInfo>>::const_iterator end() { refresh(); return _view.end(); }
  std::map<int, std::shared_ptr<FeatureInfo>>::const_iterator find(int id) { refresh(); return _view.find(id); }

  // Attach to the State whose handle holds the table (done by the first MapServerManager / updater call).
  void bind(const std::shared_ptr<State>& state) {
    if (_state.lock() == state) return;
    if (!_state.expired()) throw std::runtime_error("[MapServer]: already bound to another State");
    igv_batch* h = StateManager::handle(state);
    if (igv_tracks_capacity(h) == 0) StateManager::check(state, igv_tracks_create(h, _max_tracks), true);
    _state = state;
    _stereo = state->_state_params._cam_nums == 2;
    _dirty = true;
  }
  void touch() { _dirty = true; }   // the table was changed through the C-ABI directly

 private:
  friend class MapServerManager;
  void refresh() {
    auto state = _state.lock();
    if (!state) { _view.clear(); return; }
    igv_batch* h = StateManager::handle(state);
    // anything enqueued on the handle since the last dump (a marginalised clone, a fused update) may have changed it
    if (!_dirty && igv_launch_count(h) == _launches_at_dump) return;
    _view.clear();
    _dirty = false;
    const int T = igv_tracks_capacity(h), rho = _stereo ? 4 : 2;
    std::vector<double> times;
    std::vector<std::shared_ptr<SE3>> poses;
    for (auto& it : state->_sw_camleft_poses) { times.push_back(it.first); poses.push_back(it.second); }   // slot order
    const int SW = std::max<int>(1, (int)times.size());
    std::vector<int> id(T), anchor(T);
    std::vector<unsigned char> used(T), to_marg(T), is_tri(T);
    std::vector<unsigned long long> mask(T);
    std::vector<double> pf(3 * (std::size_t)T), obs((std::size_t)T * SW * rho);
    igv_track_dump d{};
    d.obs_slots = SW; d.id = id.data(); d.used = used.data(); d.to_marg = to_marg.data(); d.is_tri = is_tri.data();
    d.slot_mask = mask.data(); d.anchor_slot = anchor.data(); d.pf = pf.data(); d.obs = obs.data();
    StateManager::check(state, igv_tracks_get(h, &d), true);
    _launches_at_dump = igv_launch_count(h);
    for (int t = 0; t < T; ++t) {
      if (!used[t]) continue;
      auto f = std::make_shared<FeatureInfo>();
      f->_id = id[t];
      f->_isToMarg = to_marg[t] != 0;
      f->_isTri = is_tri[t] != 0;
      for (int i = 0; i < 3; ++i) f->_pf[i] = pf[3 * (std::size_t)t + i];
      if (anchor[t] >= 0 && anchor[t] < (int)poses.size()) f->_anchor = poses[anchor[t]];
      for (int s = 0; s < (int)times.size(); ++s) {
        if (!((mask[t] >> s) & 1ull)) continue;
        const double* z = &obs[((std::size_t)t * SW + s) * rho];
        if (_stereo) f->_stereo_obs[times[s]] = {z[0], z[1], z[2], z[3]};
        else f->_mono_obs[times[s]] = {z[0], z[1]};
      }
      _view[f->_id] = f;
    }
  }
  int _max_tracks;
  bool _stereo = false, _dirty = true;
  long long _launches_at_dump = -1;
  std::weak_ptr<State> _state;
  std::map<int, std::shared_ptr<FeatureInfo>> _view;
};

// MapServerManager.h (all static)
class MapServerManager {
 public:
  MapServerManager() = delete;

  // MapServerManager.cpp:189-203 (+ FeatureInfoManager::collectMonoMeas :101-143)
  static void collectMonoMeas(std::shared_ptr<MapServer> map_server, std::shared_ptr<State> state,
                              const feature_tracker::MonoFrame& mono_frame_msg) {
    const int n = (int)mono_frame_msg.mono_features.size();
    std::vector<unsigned long long> ids(std::max(n, 1));
    std::vector<double> uv(2 * (std::size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) {
      const auto& m = mono_frame_msg.mono_features[i];
      ids[i] = m.id; uv[2 * i] = m.u0; uv[2 * i + 1] = m.v0;
    }
    collect(map_server, state, n, ids.data(), uv.data(), false);
  }
  // MapServerManager.cpp:205-219 (+ FeatureInfoManager::collectStereoMeas :145-187)
  static void collectStereoMeas(std::shared_ptr<MapServer> map_server, std::shared_ptr<State> state,
                                const feature_tracker::StereoFrame& stereo_frame_msg) {
    const int n = (int)stereo_frame_msg.stereo_features.size();
    std::vector<unsigned long long> ids(std::max(n, 1));
    std::vector<double> uv(4 * (std::size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i) {
      const auto& m = stereo_frame_msg.stereo_features[i];
      ids[i] = m.id; uv[4 * i] = m.u0; uv[4 * i + 1] = m.v0; uv[4 * i + 2] = m.u1; uv[4 * i + 3] = m.v1;
    }
    collect(map_server, state, n, ids.data(), uv.data(), true);
  }
  // MapServerManager.cpp:221-246 / :248-273 (MSCKF part)
  static void markMargMonoFeatures(std::shared_ptr<MapServer> map_server, std::shared_ptr<State> state) { mark(map_server, state); }
  static void markMargStereoFeatures(std::shared_ptr<MapServer> map_server, std::shared_ptr<State> state) { mark(map_server, state); }
  // MapServerManager.cpp:454-490
  static void eraseInvalidFeatures(std::shared_ptr<MapServer> map_server, std::shared_ptr<State> state) {
    map_server->bind(state);
    StateManager::check(state, igv_tracks_erase_invalid(StateManager::handle(state), 0.2));
    map_server->_dirty = true;
  }
  // MapServerManager.cpp:396-416
  static void mapStatistics(const std::shared_ptr<MapServer> map_server) {
    std::printf("[MapServerManager]: Num of msckf feats = %d Num of slam feats = 0\n", (int)map_server->size());
  }

 private:
  static void collect(std::shared_ptr<MapServer>& map_server, std::shared_ptr<State>& state, int n, const unsigned long long* ids,
                      const double* uv, bool stereo) {
    map_server->bind(state);
    if (stereo != (state->_state_params._cam_nums == 2)) throw std::runtime_error("[MapServerManager]: mono/stereo message on the other kind of State");
    if (state->_sw_camleft_poses.find(state->_timestamp) == state->_sw_camleft_poses.end() ||
        state->_sw_camleft_poses.rbegin()->first != state->_timestamp) {
      std::printf("[FeatureInfoManager]: Meas timestamp not in sw!\n");   // MapServerManager.cpp:107-111 (assert)
      return;
    }
    if (n == 0) return;
    StateManager::check(state, igv_tracks_collect(StateManager::handle(state), &n, n, ids, uv));
    map_server->_dirty = true;
  }
  static void mark(std::shared_ptr<MapServer>& map_server, std::shared_ptr<State>& state) {
    map_server->bind(state);
    StateManager::check(state, igv_tracks_mark_lost(StateManager::handle(state)));
    map_server->_dirty = true;
  }
};

}  // namespace ingvio
