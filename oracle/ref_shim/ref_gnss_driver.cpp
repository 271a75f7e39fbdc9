// TEST INFRASTRUCTURE (oracle/): drives the REFERENCE's own GNSS path -- GnssUpdate::checkYofStatus / updateTrackedSys /
// addNewTrackedSys (GnssUpdate.cpp:33-476), GnssManager, StateManager::addVariableDelayed / ekfUpdate and gnss_comm's
// sat_states / psr_res / dopp_res / eph2pos / geph2pos / atmosphere models, all compiled unmodified from /root/reference
// against the stand-in headers of oracle/ref_shim/include -- on one recorded epoch (broadcast ephemerides + raw L1
// observations), and writes state and covariance after each call.  Pins the oracle's GNSS restatement
// (oracle/ingvio_oracle/gnss_update.py, gnss_comm.py) to the reference.
//   usage: ref_gnss_driver <input.bin> <output.bin>
// `private` members of State (the covariance) and `protected` members of GvioAligner are reached by re-declaring the
// access keywords in THIS translation unit only; the reference's own units are compiled as they are.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#define private public
#define protected public
#include "State.h"
#include "GvioAligner.h"
#undef private
#undef protected
#include "StateManager.h"
#include "GnssUpdate.h"
#include "GnssManager.h"
#include <gnss_comm/gnss_utility.hpp>
#include <gnss_comm/gnss_spp.hpp>

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
  Eigen::Matrix3d mat3() { Eigen::Matrix3d M; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M(i, j) = next(); return M; }
  Eigen::Vector3d vec3() { Eigen::Vector3d v; for (int i = 0; i < 3; ++i) v(i) = next(); return v; }
};

void dump(FILE* out, std::shared_ptr<State> state) {
  const Eigen::MatrixXd P = StateManager::getFullCov(state);
  std::vector<double> x(39, 0.0);
  const Eigen::Matrix3d& R = state->_extended_pose->valueLinearAsMat();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) x[3 * i + j] = R(i, j);
  for (int i = 0; i < 3; ++i) { x[9 + i] = state->_extended_pose->valueTrans1()(i); x[12 + i] = state->_extended_pose->valueTrans2()(i); }
  for (int i = 0; i < 3; ++i) { x[15 + i] = state->_bg->value()(i); x[18 + i] = state->_ba->value()(i); }
  double gidx[6];
  for (int g = 0; g < 6; ++g) gidx[g] = -1.0;
  for (const auto& g : state->_gnss) { x[33 + (int)g.first] = g.second->value(); gidx[(int)g.first] = g.second->idx(); }
  const double hdr[2] = {(double)P.rows(), (double)state->_gnss.size()};
  std::fwrite(hdr, sizeof(double), 2, out);
  std::fwrite(gidx, sizeof(double), 6, out);
  std::fwrite(x.data(), sizeof(double), x.size(), out);
  std::fwrite(P.data(), sizeof(double), (std::size_t)(P.rows() * P.cols()), out);
}
}  // namespace

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  Reader in(argv[1]);
  IngvioParams fp;
  fp._cam_nums = 1; fp._max_sw_clones = 4; fp._max_lm_feats = 0; fp._is_key_frame = 0; fp._enable_gnss = 1;
  fp._noise_g = 0.004; fp._noise_a = 0.08; fp._noise_bg = 2e-4; fp._noise_ba = 8e-3; fp._noise_clockbias = 2.0; fp._noise_cb_rw = 0.2;
  fp._init_cov_rot = 0; fp._init_cov_pos = 0; fp._init_cov_vel = 0.25; fp._init_cov_bg = 0.01; fp._init_cov_ba = 0.01;
  fp._init_cov_ext_rot = 0.018; fp._init_cov_ext_pos = 0.002; fp._init_cov_rcv_clockbias = 2.0; fp._init_cov_rcv_clockbias_randomwalk = 1.0;
  fp._psr_noise_amp = in.next(); fp._dopp_noise_amp = in.next();
  fp._is_adjust_yof = in.nexti(); fp._is_gnss_chi2_test = in.nexti(); fp._is_gnss_strong_reject = in.nexti();
  fp._init_cov_yof = in.next(); fp._chi2_thres = in.next(); fp._chi2_max_dof = 160;
  fp._visual_noise = 0.12; fp._frame_select_interval = 2;
  fp._gv_align_batch_size = 1; fp._gv_align_max_iter = 1; fp._gv_align_conv_epsilon = 1e-3; fp._gv_align_vel_thres = 0.1;
  fp._use_fix_time_offset = 1; fp._gnss_local_offset = 0.0;
  const Eigen::Matrix3d R0 = in.mat3();
  const Eigen::Vector3d p0 = in.vec3(), v0 = in.vec3();
  auto state = std::make_shared<State>(fp);
  state->initStateAndCov(0.0, Eigen::Quaterniond(R0), p0, v0, Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero());
  state->_extended_pose->setValueLinearByMat(R0);
  const int n_g = in.nexti();
  for (int i = 0; i < n_g; ++i) {
    const int gt = in.nexti(); const double val = in.next(), cov = in.next();
    StateManager::addGNSSVariable(state, static_cast<State::GNSSType>(gt), val, cov);
  }
  auto aligner = std::make_shared<GvioAligner>(fp);
  {
    const Eigen::Matrix3d Re = in.mat3(); const Eigen::Vector3d anchor = in.vec3(); const double yaw = in.next();
    aligner->_T_enu2ecef = Eigen::Isometry3d::Identity();
    aligner->_T_enu2ecef.linear() = Re;
    aligner->_T_enu2ecef.translation() = anchor;
    aligner->_yaw_offset = yaw;
    aligner->_isAligned = true;
  }
  GnssUpdate upd(fp);
  upd.checkYofStatus(state, aligner);
  const int N = state->curr_cov_size();
  for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) state->_cov(i, j) = in.next();   // column-major dense prior
  std::vector<double> iono(8);
  for (double& v : iono) v = in.next();
  const double t_obs = in.next();        // receiver sampling time, integer seconds (GPST-scale time_t)
  const int S = in.nexti();
  GnssMeas meas;
  const uint32_t sys_of[4] = {SYS_GPS, SYS_GLO, SYS_GAL, SYS_BDS};
  for (int i = 0; i < S; ++i) {
    const int k = in.nexti();
    const int prn = in.nexti();
    const double toe_time = in.next();
    double rec[24];
    for (double& v : rec) v = in.next();
    const double psr = in.next(), dopp = in.next(), freq = in.next(), ura = in.next(), psr_std = in.next(), dopp_std = in.next();
    gnss_comm::ObsPtr obs(new gnss_comm::Obs());
    obs->time.time = (time_t)t_obs; obs->time.sec = 0.0;
    obs->sat = gnss_comm::sat_no(sys_of[k], (uint32_t)prn);
    obs->freqs = {freq}; obs->psr = {psr}; obs->dopp = {dopp}; obs->psr_std = {psr_std}; obs->dopp_std = {dopp_std};
    obs->CN0 = {45.0}; obs->LLI = {0}; obs->code = {1}; obs->cp = {0.0}; obs->cp_std = {0.0}; obs->status = {1};
    gnss_comm::EphemBasePtr eb;
    if (k == 1) {
      gnss_comm::GloEphemPtr g(new gnss_comm::GloEphem());
      for (int c = 0; c < 3; ++c) { g->pos[c] = rec[c]; g->vel[c] = rec[3 + c]; g->acc[c] = rec[6 + c]; }
      g->tau_n = rec[9]; g->gamma = rec[10]; g->delta_tau_n = 0.0; g->freqo = 0; g->age = 0;
      eb = g;
    } else {
      gnss_comm::EphemPtr e(new gnss_comm::Ephem());
      e->A = rec[0]; e->e = rec[1]; e->i0 = rec[2]; e->OMG0 = rec[3]; e->omg = rec[4]; e->M0 = rec[5]; e->delta_n = rec[6];
      e->OMG_dot = rec[7]; e->i_dot = rec[8]; e->cuc = rec[9]; e->cus = rec[10]; e->crc = rec[11]; e->crs = rec[12];
      e->cic = rec[13]; e->cis = rec[14]; e->af0 = rec[15]; e->af1 = rec[16]; e->af2 = rec[17]; e->toe_tow = rec[18];
      e->tgd[0] = rec[19]; e->tgd[1] = 0.0; e->A_dot = 0.0; e->n_dot = 0.0; e->week = 0; e->iodc = 0; e->code = 0;
      e->toc.time = (time_t)(toe_time - rec[20]); e->toc.sec = 0.0;      // rec[20] = time_diff(toe, toc)
      eb = e;
    }
    eb->sat = obs->sat; eb->toe.time = (time_t)toe_time; eb->toe.sec = 0.0; eb->ttr = eb->toe; eb->health = 0; eb->ura = ura; eb->iode = 0;
    meas.first.push_back(obs);
    meas.second.push_back(eb);
  }
  Eigen::Matrix<double, 7, 1> spp_pos; Eigen::Vector4d spp_vel;
  for (int i = 0; i < 7; ++i) spp_pos(i) = in.next();
  for (int i = 0; i < 4; ++i) spp_vel(i) = in.next();
  SppMeas spp(meas.first[0]->time, spp_pos, spp_vel);
  const int n_add = in.nexti();          // systems to add, one addNewTrackedSys call each (the reference iterates an
  std::vector<int> add_order(n_add);     // unordered_set: the order is made explicit by masking the SPP solution)
  for (int& g : add_order) g = in.nexti();

  if (std::getenv("IGV_REF_DEBUG")) {   // per-satellite states and residuals at the prior, for cross-checking the oracle
    auto ss = gnss_comm::sat_states(meas.first, meas.second);
    Eigen::Matrix<double, 7, 1> xyzt;
    xyzt.block<3, 1>(0, 0) = aligner->getTenu2ecef() * GnssManager::calcTw2enu(state->_gnss.at(State::GNSSType::YOF)->value()) * state->_extended_pose->valueTrans1();
    xyzt.block<4, 1>(3, 0) = GnssManager::getClockbiasVec(state);
    Eigen::VectorXd rp; Eigen::MatrixXd J; std::vector<Eigen::Vector2d> atm, azel;
    gnss_comm::psr_res(xyzt, meas.first, ss, iono, rp, J, atm, azel);
    Eigen::Matrix<double, 4, 1> dv;
    dv.block<3, 1>(0, 0) = aligner->getRenu2ecef() * GnssManager::calcRw2enu(state->_gnss.at(State::GNSSType::YOF)->value()) * state->_extended_pose->valueTrans2();
    dv(3, 0) = state->_gnss.at(State::GNSSType::FS)->value();
    Eigen::VectorXd rv; Eigen::MatrixXd Jv;
    gnss_comm::dopp_res(dv, xyzt.block<3, 1>(0, 0), meas.first, ss, rv, Jv);
    for (size_t i = 0; i < ss.size(); ++i)
      std::printf("DBGV %zu vel %.9f %.9f %.9f ddt %.12e resv %.9f\n", i, ss[i]->vel(0), ss[i]->vel(1), ss[i]->vel(2), ss[i]->ddt, rv(i));
    for (size_t i = 0; i < ss.size(); ++i)
      std::printf("DBG %zu pos %.6f %.6f %.6f dt %.12e tgd %.3e ttx %.6f res %.6f ion %.6f tro %.6f el %.9f\n", i, ss[i]->pos(0), ss[i]->pos(1), ss[i]->pos(2),
                  ss[i]->dt, ss[i]->tgd, (double)(ss[i]->ttx.time - (time_t)t_obs) + ss[i]->ttx.sec, rp(i), atm[i](0), atm[i](1), azel[i](1));
  }
  FILE* out = std::fopen(argv[2], "wb");
  if (!out) { std::perror(argv[2]); return 2; }
  dump(out, state);
  upd.updateTrackedSys(state, meas, aligner, iono);
  dump(out, state);
  for (int g : add_order) {
    SppMeas one = spp;                   // only this system is "new" in the SPP solution of this call
    for (int i = 0; i < 4; ++i) if (i != g) one.posSpp(3 + i) = 0.0;
    if (g != 4) one.velSpp(3) = 0.0;
    upd.addNewTrackedSys(state, meas, one, aligner, iono);
    dump(out, state);
  }
  std::fclose(out);
  std::printf("GNSS REF DONE %d\n", (int)state->_gnss.size());
  return 0;
}
