// GNSS Updater interface of the C++ mirror (ingvio_b200/host/ingvio_gnss.hpp: GnssUpdate::checkYofStatus /
// updateTrackedSys / addNewTrackedSys with the reference's signatures, GnssUpdate.h:52-67) driven on one recorded epoch
// written by tests/test_cpp_gnss.py; writes state and covariance after each call for comparison with the oracle.
//   usage: test_gnss_update <input.bin> <output.bin>
#include <cstdio>
#include <cstdlib>

#include "../../ingvio_b200/host/ingvio_gnss.hpp"

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
  Mat3 mat3() { Mat3 M; for (double& v : M.m) v = next(); return M; }
  Vec3d vec3() { Vec3d v; for (int i = 0; i < 3; ++i) v[i] = next(); return v; }
};

static void dump(FILE* out, std::shared_ptr<State> state) {
  igv_batch* h = StateManager::handle(state);
  std::vector<double> x(igv_state_size(h));
  StateManager::check(state, igv_state_get(h, x.data()), true);
  const Matrix P = StateManager::getFullCov(state);
  const double hdr[2] = {(double)state->curr_cov_size(), (double)state->_gnss.size()};
  std::fwrite(hdr, sizeof(double), 2, out);
  std::fwrite(x.data(), sizeof(double), x.size(), out);
  std::fwrite(P.data(), sizeof(double), (std::size_t)P.rows() * P.cols(), out);
  // the host-side Type objects must agree with the device mean (the mirror keeps both)
  const double yof_dev = x[33 + State::YOF];
  if (state->_gnss.count(State::YOF) && std::fabs(state->_gnss.at(State::YOF)->value() - yof_dev) > 1e-12) {
    std::fprintf(stderr, "host / device yaw offset differ\n"); std::exit(1);
  }
}

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  Reader in(argv[1]);
  StateParams sp;
  sp._cam_nums = 1; sp._max_sw_poses = 4; sp._enable_gnss = true;
  IngvioParams fp;
  fp._psr_noise_amp = in.next(); fp._dopp_noise_amp = in.next();
  fp._is_adjust_yof = in.nexti(); fp._is_gnss_chi2_test = in.nexti(); fp._is_gnss_strong_reject = in.nexti();
  fp._init_cov_yof = in.next(); fp._chi2_thres = in.next(); fp._chi2_max_dof = 160;
  const Mat3 R0 = in.mat3(); const Vec3d p0 = in.vec3(), v0 = in.vec3();
  auto state = std::make_shared<State>(sp, 8, 32);
  Vec3d z; 
  state->initStateAndCov(0.0, R0, p0, v0, z, z);
  const int n_g = in.nexti();
  for (int i = 0; i < n_g; ++i) { const int gt = in.nexti(); const double val = in.next(), cov = in.next(); StateManager::addGNSSVariable(state, gt, val, cov); }
  auto aligner = std::make_shared<GvioAligner>();
  { const Mat3 Re = in.mat3(); const Vec3d anchor = in.vec3(); const double yaw = in.next(); aligner->setAlignment(Re, anchor, yaw); }
  GnssUpdate upd(fp);
  upd.checkYofStatus(state, aligner);                 // adds the yaw offset (GnssUpdate.cpp:33-43)
  const int N = state->curr_cov_size();
  {   // a dense prior so that every cross term of the update is exercised
    std::vector<double> P((std::size_t)N * N);
    for (double& v : P) v = in.next();
    StateManager::check(state, igv_cov_set(StateManager::handle(state), P.data(), N), true);
  }
  std::vector<double> iono(8);
  for (double& v : iono) v = in.next();
  const int S = in.nexti();
  GnssMeas meas(S);
  for (auto& s : meas) {
    s.sys = in.nexti();
    for (double& v : s.pos) v = in.next();
    for (double& v : s.vel) v = in.next();
    s.dt = in.next(); s.ddt = in.next(); s.tgd = in.next();
    s.psr = in.next(); s.dopp = in.next(); s.freq = in.next();
    s.ura = in.next(); s.psr_std = in.next(); s.dopp_std = in.next();
    s.ttx_doy = in.next(); s.ttx_sow = in.next();
  }
  SppMeas spp;
  for (double& v : spp.posSpp) v = in.next();
  for (double& v : spp.velSpp) v = in.next();

  FILE* out = std::fopen(argv[2], "wb");
  if (!out) { std::perror(argv[2]); return 2; }
  dump(out, state);
  upd.updateTrackedSys(state, meas, aligner, iono);
  dump(out, state);
  upd.addNewTrackedSys(state, meas, spp, aligner, iono);
  dump(out, state);
  // the linear-replacement and independent-add entry points of StateManager exist with the reference's signatures
  auto extra = std::make_shared<Scalar>();
  Matrix c1(1, 1); c1(0, 0) = 0.25;
  StateManager::addVariableIndependent(state, extra, c1);
  Matrix Hrep(1, 3); Hrep(0, 0) = 0.5; Hrep(0, 1) = -1.0; Hrep(0, 2) = 2.0;
  StateManager::replaceVarLinear(state, extra, {state->_bg}, Hrep);
  dump(out, state);
  if (!StateManager::checkStateContinuity(state)) { std::fprintf(stderr, "state continuity broken\n"); return 1; }
  std::fclose(out);
  std::printf("GNSS DONE %d\n", (int)state->_gnss.size());
  return 0;
}
