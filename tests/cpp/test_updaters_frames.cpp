// Frame driver for the C++ mirror of the estimator (ingvio_b200/host/ingvio_filter.hpp: IngvioFilter with its ImuPropagator,
// MapServer, RemoveLostUpdate, SwMargUpdate / KeyframeUpdate): reads a recorded stream (IMU samples + tracker messages,
// written by tests/test_cpp_updaters.py), feeds it through the reference's callbacks -- callbackIMU per sample,
// callbackMonoFrame / callbackStereoFrame per image (IngvioFilter.cpp:124-250, :252-379, :381-407) -- and writes the state
// and covariance after each frame for the Python side to compare with the oracle.
//   usage: test_updaters_frames <input.bin> <output.bin>
#include <cstdio>
#include <cstring>

#include "../../ingvio_b200/host/ingvio_filter.hpp"

using namespace ingvio;

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
  void take(double* out, int n) { for (int i = 0; i < n; ++i) out[i] = next(); }
  Mat3 mat3() { Mat3 M; take(M.m, 9); return M; }
  Vec3d vec3() { Vec3d v; take(v.v, 3); return v; }
};

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  Reader in(argv[1]);
  const int n_frames = in.nexti(), K = in.nexti(), M = in.nexti(), rho = in.nexti(), keyframe = in.nexti(), SW = in.nexti(),
            max_feats = in.nexti(), select_interval = in.nexti(), max_tracks = in.nexti();
  StateParams sp;
  sp._cam_nums = rho == 4 ? 2 : 1;
  sp._max_sw_poses = SW;
  sp._enable_gnss = false;
  sp._noise_g = in.next(); sp._noise_a = in.next(); sp._noise_bg = in.next(); sp._noise_ba = in.next();
  sp._noise_clockbias = in.next(); sp._noise_cb_rw = in.next();
  in.take(sp.gravity, 3);
  sp._T_cl2i_R = in.mat3(); sp._T_cl2i_p = in.vec3();
  sp._T_cl2cr_R = in.mat3(); sp._T_cl2cr_p = in.vec3();
  sp._init_cov_rot = in.next(); sp._init_cov_pos = in.next(); sp._init_cov_vel = in.next(); sp._init_cov_bg = in.next();
  sp._init_cov_ba = in.next(); sp._init_cov_ext_rot = in.next(); sp._init_cov_ext_pos = in.next();
  IngvioParams fp;
  fp._visual_noise = in.next(); fp._chi2_thres = in.next();
  fp._frame_select_interval = select_interval; fp._max_sw_clones = SW; fp._chi2_max_dof = 160;
  const Mat3 R0 = in.mat3(); const Vec3d p0 = in.vec3(), v0 = in.vec3(), bg0 = in.vec3(), ba0 = in.vec3();

  FilterOptions opt;
  opt._is_key_frame = keyframe != 0;
  opt._max_tracks = max_tracks;
  IngvioFilter filter(sp, fp, opt, max_feats);
  {   // the first image only arms the filter (IngvioFilter.cpp:129-133); IMU samples before it are dropped (:393)
    feature_tracker::MonoFrame first; feature_tracker::StereoFrame first_s;
    if (rho == 2) filter.callbackMonoFrame(first); else filter.callbackStereoFrame(first_s);
  }
  filter.initState(0.0, R0, p0, v0, bg0, ba0);
  auto state = filter.state();
  auto map_server = filter.mapServer();
  igv_batch* h = StateManager::handle(state);

  FILE* out = std::fopen(argv[2], "wb");
  if (!out) { std::perror(argv[2]); return 2; }
  std::vector<double> gyro(3 * K), accel(3 * K), dt(K), x(igv_state_size(h));
  double t_prev = 0.0;
  for (int k = 0; k < n_frames; ++k) {
    const double t = in.next();
    in.take(gyro.data(), 3 * K); in.take(accel.data(), 3 * K); in.take(dt.data(), K);
    const int n_meas = in.nexti();
    feature_tracker::MonoFrame mono; feature_tracker::StereoFrame stereo;
    mono.header.stamp = stereo.header.stamp = t;
    std::vector<double> ids(M), uv((std::size_t)M * rho);
    in.take(ids.data(), M); in.take(uv.data(), M * rho);
    for (int i = 0; i < n_meas; ++i) {
      if (rho == 2) { feature_tracker::MonoMeas m; m.id = (std::uint64_t)ids[i]; m.u0 = uv[2 * i]; m.v0 = uv[2 * i + 1]; mono.mono_features.push_back(m); }
      else { feature_tracker::StereoMeas m; m.id = (std::uint64_t)ids[i]; m.u0 = uv[4 * i]; m.v0 = uv[4 * i + 1]; m.u1 = uv[4 * i + 2]; m.v1 = uv[4 * i + 3]; stereo.stereo_features.push_back(m); }
    }
    // sensor_msgs/Imu stream: sample j of this interval is stamped at the END of its step; the last one at the image time
    double acc_dt = 0.0, tot = 0.0;
    for (int j = 0; j < K; ++j) tot += dt[j];
    for (int j = 0; j < K; ++j) {
      acc_dt += dt[j];
      const double stamp = (j == K - 1) ? t : t_prev + (t - t_prev) * (acc_dt / tot);
      Vec3d w, a;
      for (int i = 0; i < 3; ++i) { w[i] = gyro[3 * j + i]; a[i] = accel[3 * j + i]; }
      filter.callbackIMU(ImuCtrl(stamp, w, a));
    }
    if (rho == 2) filter.callbackMonoFrame(mono); else filter.callbackStereoFrame(stereo);
    t_prev = t;
    if (state->_timestamp != t) { std::fprintf(stderr, "frame %d: state time %.9f != image time %.9f\n", k, state->_timestamp, t); return 1; }
    if (filter.imuPropagator()->bufferSize() != 0) { std::fprintf(stderr, "frame %d: %zu IMU samples left in the buffer\n", k, filter.imuPropagator()->bufferSize()); return 1; }
    StateManager::check(state, igv_state_get(h, x.data()), true);
    const Matrix P = StateManager::getFullCov(state);
    const double hdr[3] = {(double)state->curr_cov_size(), (double)state->_sw_camleft_poses.size(), (double)map_server->size()};
    std::fwrite(hdr, sizeof(double), 3, out);
    std::fwrite(x.data(), sizeof(double), x.size(), out);
    std::fwrite(P.data(), sizeof(double), (std::size_t)P.rows() * P.cols(), out);
    if (!StateManager::checkStateContinuity(state)) { std::fprintf(stderr, "state continuity broken at frame %d\n", k); return 1; }
  }
  std::fclose(out);
  std::printf("FRAMES DONE %d\n", n_frames);
  return 0;
}
