// ingvio_filter.hpp -- C++ host mirror of the reference's IMU propagator front and of the frame callbacks.
//
// Reference: ImuPropagator.h:40-140 / ImuPropagator.cpp:232-341 (ImuCtrl, storeImu, propagateUntil,
// propagateAugmentAtEnd, propagateToExpectedPoseAndAugment) and IngvioFilter.h:57-140 / IngvioFilter.cpp:124-250,
// :252-379, :381-407 (callbackMonoFrame, callbackStereoFrame, callbackIMU) without ROS transport, visualisation, the
// GNSS front end and the SLAM-landmark branch (max_landmark_features = 0 in every shipped config).
//
// ImuPropagator keeps the reference's sample buffer and resolves propagateUntil's loop (:244-271: skip samples older than
// the state, stop after t_end, dt = stamp - state time, skip dt < 1e-6, a last partial step with the last used sample,
// erase what was consumed) into one list of (gyro, accel, dt) steps, which goes to the device in ONE call
// (igv_propagate_imu: mean and covariance, ImuPropagator.cpp:98-162 + StateManager.cpp:42-119 per step).
// IngvioFilter is the reference's callback sequence written against the mirrored classes only.
//
// Header-only, plain C++17.
#pragma once
#include <deque>

#include "ingvio_updaters.hpp"

namespace ingvio {

class ImuCtrl {   // ImuPropagator.h:40-85
 public:
  ImuCtrl() {}
  ImuCtrl(double t, const Vec3d& gyro, const Vec3d& accel) : _timestamp(t), _accel_raw(accel), _gyro_raw(gyro) {}
  double _timestamp = -1;
  Vec3d _accel_raw, _gyro_raw;
};

class ImuPropagator {
 public:
  ImuPropagator() {}
  explicit ImuPropagator(const IngvioParams&) {}
  // The static initialisation from the first samples (ImuPropagator.cpp:30-96: gravity direction -> initial attitude) is
  // ROS-side start-up logic and stays with the caller; gravity itself is a StateParams value of the handle.
  void storeImu(const ImuCtrl& imu_ctrl) { _imu_ctrl_buffer.push_back(imu_ctrl); }
  std::size_t bufferSize() const { return _imu_ctrl_buffer.size(); }

  // ImuPropagator.cpp:232-292
  void propagateUntil(std::shared_ptr<State> state, double t_end) {
    if (!_has_gravity_set || t_end <= state->_timestamp) return;
    if (_imu_ctrl_buffer.size() == 0) return;
    if (_imu_ctrl_buffer[0]._timestamp > t_end) return;
    int propa_cnt = 0;
    ImuCtrl last_imu_ctrl = _imu_ctrl_buffer[_imu_ctrl_buffer.size() - 1];
    double t = state->_timestamp;                    // state->_timestamp as stateAndCovTransition advances it (+= dt)
    std::vector<double> gyro, accel, dts;
    auto push = [&](const ImuCtrl& c, double dt) {
      for (int i = 0; i < 3; ++i) { gyro.push_back(c._gyro_raw[i]); accel.push_back(c._accel_raw[i]); }
      dts.push_back(dt);
      t += dt;
    };
    for (std::size_t i = 0; i < _imu_ctrl_buffer.size(); ++i) {
      const double ctrl_time = _imu_ctrl_buffer[i]._timestamp;
      if (ctrl_time < t) { ++propa_cnt; continue; }
      if (ctrl_time > t_end) break;
      ++propa_cnt;
      const double dt = ctrl_time - t;
      if (dt < 1e-6) continue;
      last_imu_ctrl = _imu_ctrl_buffer[i];
      push(_imu_ctrl_buffer[i], dt);
    }
    if (t < t_end) {
      const double dt_last = t_end - t;
      if (dt_last > 1e-06) push(last_imu_ctrl, dt_last);
      else t = t_end;
    }
    if (!dts.empty()) {
      StateManager::check(state, igv_propagate_imu(StateManager::handle(state), (int)dts.size(), gyro.data(), accel.data(), dts.data()), true);
      StateManager::sync_mean_from_device(state);
    }
    state->_timestamp = t;
    _imu_ctrl_buffer.erase(_imu_ctrl_buffer.begin(), _imu_ctrl_buffer.begin() + propa_cnt);
  }

  // ImuPropagator.cpp:294-314
  void propagateAugmentAtEnd(std::shared_ptr<State> state, double t_end) {
    if (!_has_gravity_set) return;
    propagateUntil(state, t_end);
    if (state->_timestamp < t_end) { std::printf("[ImuPropagator]: Cannot propa to t_end due to no imu ctrl!\n"); return; }
    else if (state->_timestamp > t_end) { std::printf("[IMUPropagator]: Cannot propa because t_end < curr state time!\n"); return; }
    StateManager::augmentSlidingWindowPose(state);
  }

 protected:
  bool _has_gravity_set = true;
  std::deque<ImuCtrl> _imu_ctrl_buffer;
};

// The options of IngvioParams the callbacks read (IngvioParams.h), next to the updater / triangulator values above.
struct FilterOptions {
  bool _is_key_frame = true;      // config: is_key_frame (1 in every shipped config)
  int _max_tracks = 512;          // capacity of the device track table (no counterpart in the reference)
};

class IngvioFilter {   // IngvioFilter.h:57-140, the estimator side
 public:
  IngvioFilter(const StateParams& sp, const IngvioParams& fp, const FilterOptions& opt, int max_feats = 400)
      : _filter_params(fp), _options(opt) {
    _state = std::make_shared<State>(sp, max_feats, 1);                    // IngvioFilter.cpp:76-88
    _imu_propa = std::make_shared<ImuPropagator>(fp);
    _tri = std::make_shared<Triangulator>(fp);
    _map_server = std::make_shared<MapServer>(opt._max_tracks);
    _remove_lost_update = std::make_shared<RemoveLostUpdate>(fp);
    _keyframe_update = std::make_shared<KeyframeUpdate>(fp);
    _sw_marg_update = std::make_shared<SwMargUpdate>(fp);
  }
  IngvioFilter(const IngvioFilter&) = delete;

  // In the reference the initial attitude comes from ImuPropagator::getInitQuat inside callbackIMU (:397-405); here the
  // caller supplies the initial state.
  void initState(double t0, const Mat3& R_i2w, const Vec3d& p, const Vec3d& v, const Vec3d& bg, const Vec3d& ba) {
    _state->initStateAndCov(t0, R_i2w, p, v, bg, ba);
    _hasInitState = true;
  }

  void callbackIMU(const ImuCtrl& imu_msg) {                                // IngvioFilter.cpp:381-407
    if (!_hasImageCome) return;
    _imu_propa->storeImu(imu_msg);
  }

  void callbackMonoFrame(const feature_tracker::MonoFrame& mono_frame) {    // IngvioFilter.cpp:124-250
    if (!_hasImageCome) { _hasImageCome = true; return; }
    if (!_hasInitState) return;
    const double target_time = mono_frame.header.toSec();
    if (_state->_timestamp >= target_time) return;
    _imu_propa->propagateAugmentAtEnd(_state, target_time);
    if (_state->_timestamp < target_time) return;
    MapServerManager::collectMonoMeas(_map_server, _state, mono_frame);
    _remove_lost_update->updateStateMono(_state, _map_server, _tri);
    if (_options._is_key_frame) {
      _keyframe_update->updateStateMono(_state, _map_server, _tri);
      _keyframe_update->cleanMonoObsAtMargTime(_state, _map_server);
      _keyframe_update->changeMSCKFAnchor(_state, _map_server);
      _keyframe_update->margSwPose(_state);
    } else {
      _sw_marg_update->updateStateMono(_state, _map_server, _tri);
      _sw_marg_update->cleanMonoObsAtMargTime(_state, _map_server);
      _sw_marg_update->changeMSCKFAnchor(_state, _map_server);
      _sw_marg_update->margSwPose(_state);
    }
    MapServerManager::eraseInvalidFeatures(_map_server, _state);
  }

  void callbackStereoFrame(const feature_tracker::StereoFrame& stereo_frame) {   // IngvioFilter.cpp:252-379
    if (!_hasImageCome) { _hasImageCome = true; return; }
    if (!_hasInitState) return;
    const double target_time = stereo_frame.header.toSec();
    if (_state->_timestamp >= target_time) return;
    _imu_propa->propagateAugmentAtEnd(_state, target_time);
    if (_state->_timestamp < target_time) return;
    MapServerManager::collectStereoMeas(_map_server, _state, stereo_frame);
    _remove_lost_update->updateStateStereo(_state, _map_server, _tri);
    if (_options._is_key_frame) {
      _keyframe_update->updateStateStereo(_state, _map_server, _tri);
      _keyframe_update->cleanStereoObsAtMargTime(_state, _map_server);
      _keyframe_update->changeMSCKFAnchor(_state, _map_server);
      _keyframe_update->margSwPose(_state);
    } else {
      _sw_marg_update->updateStateStereo(_state, _map_server, _tri);
      _sw_marg_update->cleanStereoObsAtMargTime(_state, _map_server);
      _sw_marg_update->changeMSCKFAnchor(_state, _map_server);
      _sw_marg_update->margSwPose(_state);
    }
    MapServerManager::eraseInvalidFeatures(_map_server, _state);
  }

  std::shared_ptr<State> state() const { return _state; }
  std::shared_ptr<MapServer> mapServer() const { return _map_server; }
  std::shared_ptr<ImuPropagator> imuPropagator() const { return _imu_propa; }

 protected:
  IngvioParams _filter_params;
  FilterOptions _options;
  bool _hasImageCome = false, _hasInitState = false;
  std::shared_ptr<State> _state;
  std::shared_ptr<ImuPropagator> _imu_propa;
  std::shared_ptr<Triangulator> _tri;
  std::shared_ptr<MapServer> _map_server;
  std::shared_ptr<RemoveLostUpdate> _remove_lost_update;
  std::shared_ptr<KeyframeUpdate> _keyframe_update;
  std::shared_ptr<SwMargUpdate> _sw_marg_update;
};

}  // namespace ingvio
