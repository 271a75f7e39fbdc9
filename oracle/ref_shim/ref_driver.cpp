// TEST INFRASTRUCTURE (oracle/): drives the REFERENCE's own classes -- compiled unmodified from /root/reference against
// the stand-in headers of oracle/ref_shim/include -- over a recorded stream (IMU samples + tracker messages, the format of
// tests/test_cpp_updaters.py), in the call order of IngvioFilter::callbackMonoFrame / callbackStereoFrame
// (/root/reference/ingvio_estimator/src/IngvioFilter.cpp:124-205, :252-334; the ROS plumbing, the SLAM-landmark branch
// with max_lm_feats = 0 and the GNSS branch are not on this path), and writes the state and covariance after every
// frame.  The numpy oracle (and the CUDA path) are compared with this output: it is what pins the oracle to the reference.
//   usage: ref_driver <input.bin> <output.bin>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include "IngvioParams.h"
#include "State.h"
#include "StateManager.h"
#include "ImuPropagator.h"
#include "MapServer.h"
#include "MapServerManager.h"
#include "Triangulator.h"
#include "RemoveLostUpdate.h"
#include "SwMargUpdate.h"
#include "KeyframeUpdate.h"
#include "LandmarkUpdate.h"

using namespace ingvio;

namespace {

struct Reader {
  std::vector<double> d; std::size_t pos = 0;
  explicit Reader(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    d.resize(n / sizeof(double));
    if (std::fread(d.data(), sizeof(double), d.size(), f) != d.size()) { std::fprintf(stderr, "short read\n"); std::exit(2); }
    std::fclose(f);
  }
  double next() { if (pos >= d.size()) { std::fprintf(stderr, "input exhausted\n"); std::exit(2); } return d[pos++]; }
  int nexti() { return (int)next(); }
  Eigen::Matrix3d mat3() { Eigen::Matrix3d M; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M(i, j) = next(); return M; }   // row-major in the file
  Eigen::Vector3d vec3() { Eigen::Vector3d v; for (int i = 0; i < 3; ++i) v(i) = next(); return v; }
};

void put_rot(std::vector<double>& x, std::size_t at, const Eigen::Matrix3d& R) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) x[at + 3 * i + j] = R(i, j);
}
void put_vec(std::vector<double>& x, std::size_t at, const Eigen::Vector3d& v) { for (int i = 0; i < 3; ++i) x[at + i] = v(i); }

}  // namespace

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  Reader in(argv[1]);
  const int n_frames = in.nexti(), K = in.nexti(), M = in.nexti(), rho = in.nexti(), keyframe = in.nexti(), SW = in.nexti();
  const int max_feats = in.nexti(), select_interval = in.nexti(), max_tracks = in.nexti();
  (void)max_feats; (void)max_tracks;
  IngvioParams fp;
  fp._cam_nums = rho == 4 ? 2 : 1;
  fp._max_sw_clones = SW;
  fp._is_key_frame = keyframe;
  // SLAM landmarks (LandmarkUpdate, mono): off in the recorded format; IGV_REF_MAX_LM > 0 switches the branch of
  // IngvioFilter.cpp:155-173 / :181-194 on
  const char* lm_env = std::getenv("IGV_REF_MAX_LM");
  fp._max_lm_feats = lm_env ? std::atoi(lm_env) : 0;
  fp._enable_gnss = 0;
  fp._noise_g = in.next(); fp._noise_a = in.next(); fp._noise_bg = in.next(); fp._noise_ba = in.next();
  fp._noise_clockbias = in.next(); fp._noise_cb_rw = in.next();
  const Eigen::Vector3d gravity = in.vec3();
  fp._init_gravity = -gravity(2);
  fp._max_imu_buffer_size = 100000;
  fp._init_imu_buffer_sp = -1;            // gravity (0, 0, -g) and identity attitude are given, not estimated
  const Eigen::Matrix3d R_cl2i = in.mat3(); const Eigen::Vector3d p_cl2i = in.vec3();
  const Eigen::Matrix3d R_cl2cr = in.mat3(); const Eigen::Vector3d p_cl2cr = in.vec3();
  fp._T_cl2i.linear() = R_cl2i; fp._T_cl2i.translation() = p_cl2i;
  Eigen::Isometry3d T_cl2cr = Eigen::Isometry3d::Identity();
  T_cl2cr.linear() = R_cl2cr; T_cl2cr.translation() = p_cl2cr;
  fp._T_cr2i = fp._T_cl2i * T_cl2cr.inverse();   // StateParams forms T_cl2cr = T_cr2i^-1 T_cl2i (State.cpp:31)
  fp._init_cov_rot = in.next(); fp._init_cov_pos = in.next(); fp._init_cov_vel = in.next(); fp._init_cov_bg = in.next();
  fp._init_cov_ba = in.next(); fp._init_cov_ext_rot = in.next(); fp._init_cov_ext_pos = in.next();
  fp._init_cov_rcv_clockbias = 2.0; fp._init_cov_rcv_clockbias_randomwalk = 1.0; fp._init_cov_yof = 0.015;
  fp._visual_noise = in.next(); fp._chi2_thres = in.next();
  fp._chi2_max_dof = 160;
  fp._frame_select_interval = select_interval;
  // Triangulator defaults of the shipped configs (config/sportsfield/ingvio_mono.yaml) = Triangulator.h:85-93
  fp._trans_thres = 0.1; fp._huber_epsilon = 0.01; fp._conv_precision = 5e-7; fp._init_damping = 1e-3;
  fp._outer_loop_max_iter = 10; fp._inner_loop_max_iter = 10; fp._max_depth = 60.0; fp._min_depth = 0.2;
  fp._max_baseline_ratio = 40.0;
  const Eigen::Matrix3d R0 = in.mat3();
  const Eigen::Vector3d p0 = in.vec3(), v0 = in.vec3(), bg0 = in.vec3(), ba0 = in.vec3();

  auto state = std::make_shared<State>(fp);
  auto imu_propa = std::make_shared<ImuPropagator>(fp);
  auto tri = std::make_shared<Triangulator>(fp);
  auto map_server = std::make_shared<MapServer>();
  auto remove_lost = std::make_shared<RemoveLostUpdate>(fp);
  auto sw_marg = std::make_shared<SwMargUpdate>(fp);
  auto kf_update = std::make_shared<KeyframeUpdate>(fp);
  auto lm_update = std::make_shared<LandmarkUpdate>(fp);
  const bool use_lm = fp._max_lm_feats > 0 && rho == 2;
  state->initStateAndCov(0.0, Eigen::Quaterniond(R0), p0, v0, bg0, ba0);
  // the quaternion round trip must not perturb the attitude the other implementations start from
  state->_extended_pose->setValueLinearByMat(R0);

  FILE* out = std::fopen(argv[2], "wb");
  if (!out) { std::perror(argv[2]); return 2; }
  const std::size_t xs = 39 + 12 * (std::size_t)(SW + 1);
  std::vector<double> gyro(3 * K), accel(3 * K), dt(K), ids(M), uv((std::size_t)M * rho), x(xs);
  double t_prev = 0.0;
  for (int k = 0; k < n_frames; ++k) {
    const double t = in.next();
    for (auto& v : gyro) v = in.next();
    for (auto& v : accel) v = in.next();
    for (auto& v : dt) v = in.next();
    const int n_meas = in.nexti();
    for (auto& v : ids) v = in.next();
    for (auto& v : uv) v = in.next();
    // IMU stream: sample j of the interval is stamped at the END of its step, the last one at the image time
    double acc_dt = 0.0, tot = 0.0;
    for (int j = 0; j < K; ++j) tot += dt[j];
    for (int j = 0; j < K; ++j) {
      acc_dt += dt[j];
      ImuCtrl c;
      c._timestamp = (j == K - 1) ? t : t_prev + (t - t_prev) * (acc_dt / tot);
      c._gyro_raw = Eigen::Vector3d(gyro[3 * j], gyro[3 * j + 1], gyro[3 * j + 2]);
      c._accel_raw = Eigen::Vector3d(accel[3 * j], accel[3 * j + 1], accel[3 * j + 2]);
      imu_propa->storeImu(c);
    }
    imu_propa->propagateAugmentAtEnd(state, t);                       // IngvioFilter.cpp:143
    if (state->_timestamp != t) { std::fprintf(stderr, "frame %d: state time %.9f != image time %.9f\n", k, state->_timestamp, t); return 1; }
    if (rho == 2) {
      auto msg = std::make_shared<feature_tracker::MonoFrame>();
      msg->header.stamp = ros::Time(t);
      for (int i = 0; i < n_meas; ++i) {
        feature_tracker::MonoMeas m; m.id = (uint64_t)ids[i]; m.u0 = uv[2 * i]; m.v0 = uv[2 * i + 1];
        msg->mono_features.push_back(m);
      }
      MapServerManager::collectMonoMeas(map_server, state, msg);      // :147
      remove_lost->updateStateMono(state, map_server, tri);           // :149
      if (keyframe) {
        kf_update->updateStateMono(state, map_server, tri);           // :153
        if (use_lm) {
          lm_update->updateLandmarkMono(state, map_server);           // :157
          lm_update->initNewLandmarkMono(state, map_server, tri, fp._max_sw_clones);   // :159-160
        }
        kf_update->cleanMonoObsAtMargTime(state, map_server);         // :163
        kf_update->changeMSCKFAnchor(state, map_server);              // :165
        if (use_lm) {
          std::vector<double> marg_kfs;
          kf_update->getMargKfs(state, marg_kfs);                     // :169-170
          lm_update->changeLandmarkAnchor(state, map_server, marg_kfs);   // :172
        }
        kf_update->margSwPose(state);                                 // :175
      } else {
        sw_marg->updateStateMono(state, map_server, tri);             // :179
        if (use_lm) {
          lm_update->updateLandmarkMono(state, map_server);           // :183
          lm_update->initNewLandmarkMono(state, map_server, tri, fp._max_sw_clones);   // :185-186
        }
        sw_marg->cleanMonoObsAtMargTime(state, map_server);           // :189
        sw_marg->changeMSCKFAnchor(state, map_server);                // :191
        if (use_lm) lm_update->changeLandmarkAnchor(state, map_server);   // :193-194
        sw_marg->margSwPose(state);                                   // :196
      }
    } else {
      auto msg = std::make_shared<feature_tracker::StereoFrame>();
      msg->header.stamp = ros::Time(t);
      for (int i = 0; i < n_meas; ++i) {
        feature_tracker::StereoMeas m; m.id = (uint64_t)ids[i];
        m.u0 = uv[4 * i]; m.v0 = uv[4 * i + 1]; m.u1 = uv[4 * i + 2]; m.v1 = uv[4 * i + 3];
        msg->stereo_features.push_back(m);
      }
      MapServerManager::collectStereoMeas(map_server, state, msg);    // :275
      remove_lost->updateStateStereo(state, map_server, tri);         // :277
      if (keyframe) {
        kf_update->updateStateStereo(state, map_server, tri);
        kf_update->cleanStereoObsAtMargTime(state, map_server);
        kf_update->changeMSCKFAnchor(state, map_server);
        kf_update->margSwPose(state);
      } else {
        sw_marg->updateStateStereo(state, map_server, tri);
        sw_marg->cleanStereoObsAtMargTime(state, map_server);
        sw_marg->changeMSCKFAnchor(state, map_server);
        sw_marg->margSwPose(state);
      }
    }
    MapServerManager::eraseInvalidFeatures(map_server, state);        // :199
    t_prev = t;

    std::fill(x.begin(), x.end(), 0.0);
    put_rot(x, 0, state->_extended_pose->valueLinearAsMat());
    put_vec(x, 9, state->_extended_pose->valueTrans1());
    put_vec(x, 12, state->_extended_pose->valueTrans2());
    put_vec(x, 15, state->_bg->value());
    put_vec(x, 18, state->_ba->value());
    put_rot(x, 21, state->_camleft_imu_extrinsics->valueLinearAsMat());
    put_vec(x, 30, state->_camleft_imu_extrinsics->valueTrans());
    for (const auto& g : state->_gnss) x[33 + (int)g.first] = g.second->value();
    std::size_t s = 0;
    for (const auto& c : state->_sw_camleft_poses) {
      if (39 + 12 * s + 12 > xs) break;
      put_rot(x, 39 + 12 * s, c.second->valueLinearAsMat());
      put_vec(x, 39 + 12 * s + 9, c.second->valueTrans());
      ++s;
    }
    const Eigen::MatrixXd P = StateManager::getFullCov(state);
    const double hdr[3] = {(double)state->curr_cov_size(), (double)state->_sw_camleft_poses.size(), (double)map_server->size()};
    std::fwrite(hdr, sizeof(double), 3, out);
    std::fwrite(x.data(), sizeof(double), x.size(), out);
    std::fwrite(P.data(), sizeof(double), (std::size_t)(P.rows() * P.cols()), out);   // column-major
    if (use_lm) {   // landmarks in the state: count, then (id, covariance index, world xyz) in ascending id
      std::map<int, std::shared_ptr<AnchoredLandmark>> lms(state->_anchored_landmarks.begin(), state->_anchored_landmarks.end());
      const double nl = (double)lms.size();
      std::fwrite(&nl, sizeof(double), 1, out);
      for (const auto& it : lms) {
        const double rec[5] = {(double)it.first, (double)it.second->idx(), it.second->valuePosXyz()(0), it.second->valuePosXyz()(1), it.second->valuePosXyz()(2)};
        std::fwrite(rec, sizeof(double), 5, out);
      }
    }
  }
  std::fclose(out);
  std::printf("FRAMES DONE %d\n", n_frames);
  return 0;
}
