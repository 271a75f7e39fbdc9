// GnssUpdate with the reference's interface (GnssUpdate.h:35-105) over the C-ABI: checkYofStatus, removeUntrackedSys,
// updateTrackedSys, addNewTrackedSys keep their names, argument meaning and early-return conditions
// (/root/reference/ingvio_estimator/src/GnssUpdate.cpp:33-476); the bodies are the device chain
//   igv_gnss_residuals (gnss_comm::psr_res / dopp_res at the filter's receiver state) -> igv_gnss_update, or
//   igv_gnss_residuals (with the SPP initial values) -> igv_gnss_add_new_tracked_sys per new system.
// gnss_comm's Obs / Ephem pointer pairs become GnssMeas below: one record per satellite at the output boundary of
// gnss_comm::sat_states (igv_sat_states produces it from ephemeris records); GvioAligner keeps the five getters the
// update reads. Header-only, C++17, links only against libingvio_b200.so.
#pragma once
#include <array>
#include <unordered_set>

#include "ingvio_updaters.hpp"

namespace ingvio {

// One satellite of an epoch: gnss_comm::SatState (pos / vel / dt / ddt / tgd / ttx) + the L1 observation of its Obs.
struct GnssSat {
  int sys = 0;                       // State::GNSSType GPS..BDS (= gnss_comm::sys2idx)
  double pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0};
  double dt = 0, ddt = 0, tgd = 0;   // satellite clock [s], drift [s/s], group delay [s]
  double psr = 0, dopp = 0, freq = 0;            // L1 pseudo-range [m], Doppler [Hz], carrier [Hz] (<= 0: no L1 observation)
  double ura = 1, psr_std = 1, dopp_std = 1;     // ephemeris ura, Obs::psr_std / dopp_std of the L1 observation
  double ttx_doy = 1, ttx_sow = 0;               // transmit time: day of year, GPS seconds of week
};
typedef std::vector<GnssSat> GnssMeas;

struct SppMeas {                      // GnssSync.h:37-55
  std::array<double, 7> posSpp{};     // ECEF position + 4 receiver clock biases [m]
  std::array<double, 4> velSpp{};     // ECEF velocity + clock drift
};

// GvioAligner.h:54-68 -- only what GnssUpdate reads; the batch alignment itself is outside the hot path.
class GvioAligner {
 public:
  bool isAlign() const { return _isAligned; }
  double getYawOffset() const { return _yaw_offset; }
  const Mat3& getRenu2ecef() const { return _R_enu2ecef; }
  Mat3 getRecef2enu() const { Mat3 T; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T(i, j) = _R_enu2ecef(j, i); return T; }
  void getTenu2ecef(double T12[12]) const { for (int i = 0; i < 9; ++i) T12[i] = _R_enu2ecef.m[i]; for (int i = 0; i < 3; ++i) T12[9 + i] = _anchor_ecef[i]; }
  void setAlignment(const Mat3& R_enu2ecef, const Vec3d& anchor_ecef, double yaw_offset) {
    _R_enu2ecef = R_enu2ecef; _anchor_ecef = anchor_ecef; _yaw_offset = yaw_offset; _isAligned = true;
  }
 private:
  bool _isAligned = false;
  double _yaw_offset = 0.0;
  Mat3 _R_enu2ecef;
  Vec3d _anchor_ecef;
};

class GnssManager {   // GnssManager.cpp:60-134
 public:
  static bool checkGnssStates(const std::shared_ptr<State>& state) {
    if (!state->_gnss.count(State::YOF) || !state->_gnss.count(State::FS)) return false;
    for (int i = 0; i < 4; ++i) if (state->_gnss.count(i)) return true;
    return false;
  }
  static std::array<double, 4> getClockbiasVec(const std::shared_ptr<State>& state) {
    std::array<double, 4> cb{};
    for (auto& g : state->_gnss) if (g.first != State::YOF && g.first != State::FS) cb[g.first] = g.second->value();
    return cb;
  }
};

class GnssUpdate : public UpdateBase {
 public:
  GnssUpdate(const IngvioParams& fp)   // GnssUpdate.h:43-50
      : UpdateBase(fp._chi2_max_dof, fp._chi2_thres), _psr_noise_amp(fp._psr_noise_amp), _dopp_noise_amp(fp._dopp_noise_amp),
        _is_adjust_yof(fp._is_adjust_yof), _is_gnss_chi2_test(fp._is_gnss_chi2_test),
        _is_gnss_strong_reject(fp._is_gnss_strong_reject), _init_cov_yof(fp._init_cov_yof) {}

  // GnssUpdate.cpp:33-43
  void checkYofStatus(std::shared_ptr<State> state, std::shared_ptr<GvioAligner> gvio_aligner) {
    if (!state->_state_params._enable_gnss || !gvio_aligner->isAlign()) return;
    if (!state->_gnss.count(State::YOF))
      StateManager::addGNSSVariable(state, State::YOF, gvio_aligner->getYawOffset(), _init_cov_yof);   // state->_state_params._init_cov_yof
  }

  // GnssUpdate.cpp:66-82
  void removeUntrackedSys(std::shared_ptr<State> state, const GnssMeas& gnss_meas) {
    if (!state->_state_params._enable_gnss) return;
    getSysInGnssMeas(gnss_meas);
    std::vector<int> to_marg;
    for (auto& g : state->_gnss) if (!_gnss_sys.count(g.first)) to_marg.push_back(g.first);
    for (int g : to_marg) StateManager::margGNSSVariable(state, g);
  }

  // GnssUpdate.cpp:84-293
  void updateTrackedSys(std::shared_ptr<State> state, const GnssMeas& gnss_meas, std::shared_ptr<GvioAligner> gvio_aligner,
                        const std::vector<double>& iono_params) {
    if (!state->_state_params._enable_gnss || !gvio_aligner->isAlign()) return;
    if (gnss_meas.size() <= 0 || iono_params.size() != 8) return;
    if (!GnssManager::checkGnssStates(state)) return;
    Residuals r;
    residuals(state, gnss_meas, gvio_aligner, iono_params, nullptr, r);
    igv_batch* h = StateManager::handle(state);
    upload(state);
    igv_gnss_args a{};
    a.n_sats = (int)gnss_meas.size();
    a.unit = r.unit.data(); a.res_pos = r.res_pos.data(); a.res_vel = r.res_vel.data();
    a.sigma_psr = r.sig_psr.data(); a.sigma_dopp = r.sig_dopp.data(); a.sys = r.sys.data();
    a.R_enu2ecef = gvio_aligner->getRenu2ecef().m;
    a.is_adjust_yof = _is_adjust_yof; a.chi2_test = _is_gnss_chi2_test; a.strong_reject = _is_gnss_strong_reject;
    StateManager::check(state, igv_gnss_update(h, &a));
    StateManager::sync_mean_from_device(state);
  }

  // GnssUpdate.cpp:317-476. The reference iterates an unordered_set; here new systems are added in GNSSType order.
  void addNewTrackedSys(std::shared_ptr<State> state, const GnssMeas& gnss_meas, const SppMeas& spp_meas,
                        std::shared_ptr<GvioAligner> gvio_aligner, const std::vector<double>& iono_params) {
    if (!state->_state_params._enable_gnss || !gvio_aligner->isAlign()) return;
    if (gnss_meas.size() <= 0 || iono_params.size() != 8) return;
    if (!state->_gnss.count(State::YOF)) return;
    getSysInSppMeas(spp_meas);
    std::vector<int> sys_to_add;
    for (int g = 0; g <= State::FS; ++g) if (_spp_sys.count(g) && !state->_gnss.count(g)) sys_to_add.push_back(g);   // calcSysToAdd
    if (sys_to_add.empty()) return;
    // xyzt / dopp with the SPP values of the systems to add (:351-370); NaN = the state's own entry
    double ci[5];
    for (double& v : ci) v = std::numeric_limits<double>::quiet_NaN();
    for (int g : sys_to_add) ci[g] = (g == State::FS) ? spp_meas.velSpp[3] : spp_meas.posSpp[3 + g];
    igv_batch* h = StateManager::handle(state);
    const Mat3 R_e2n = gvio_aligner->getRecef2enu();
    for (int g : sys_to_add) {
      Residuals r;
      residuals(state, gnss_meas, gvio_aligner, iono_params, ci, r);   // re-evaluated: the previous addition moved the state
      const double value = ci[g];
      int accepted = 0;
      const int n0 = state->curr_cov_size();
      igv_gnss_new_sys_args a{};
      a.n_sats = (int)gnss_meas.size(); a.gtype = g; a.value = &value;
      a.unit = r.unit.data(); a.res_pos = r.res_pos.data(); a.res_vel = r.res_vel.data();
      a.sigma_psr = r.sig_psr.data(); a.sigma_dopp = r.sig_dopp.data(); a.sys = r.sys.data();
      a.R_enu2ecef = gvio_aligner->getRenu2ecef().m; a.R_ecef2enu = R_e2n.m;
      a.is_adjust_yof = _is_adjust_yof; a.chi2_mult = 0.95; a.prior_cov_if_rejected = 1.0; a.accepted_out = &accepted;
      const igv_status st = igv_gnss_add_new_tracked_sys(h, &a);
      StateManager::check(state, st);
      if (st != IGV_OK) continue;
      if (igv_dim(h) == n0) continue;                    // nothing was added (no yaw offset in the state)
      if (!accepted) {                                   // single filter: drop the decoupled placeholder = "continue" of :430,:470
        StateManager::check(state, igv_marginalize(h, n0));
        std::printf("[StateManager]: Cannot add variable due to chi2 test failure!\n");
        continue;
      }
      auto var = std::make_shared<Scalar>();
      var->setValue(value);
      StateManager::registerVariable(state, var, n0);
      state->_gnss[g] = var;
      StateManager::sync_mean_from_device(state);
    }
  }

 protected:
  double _psr_noise_amp, _dopp_noise_amp;
  int _is_adjust_yof, _is_gnss_chi2_test, _is_gnss_strong_reject;
  double _init_cov_yof;
  std::unordered_set<int> _gnss_sys, _spp_sys;

  void getSysInGnssMeas(const GnssMeas& gnss_meas) {   // :46-64
    _gnss_sys.clear();
    _gnss_sys.insert(State::YOF);
    bool flag = false;
    for (const auto& s : gnss_meas) if (!_gnss_sys.count(s.sys)) { flag = true; _gnss_sys.insert(s.sys); }
    if (flag) _gnss_sys.insert(State::FS);
  }
  void getSysInSppMeas(const SppMeas& spp) {            // :295-305
    _spp_sys.clear();
    if (std::fabs(spp.velSpp[3]) > 1e-3) _spp_sys.insert(State::FS);
    for (int i = 0; i < 4; ++i) if (std::fabs(spp.posSpp[3 + i]) > 1e-3) _spp_sys.insert(i);
  }

  struct Residuals { std::vector<double> unit, res_pos, res_vel, sig_psr, sig_dopp; std::vector<int> sys; };
  void residuals(std::shared_ptr<State> state, const GnssMeas& m, std::shared_ptr<GvioAligner> al, const std::vector<double>& iono,
                 const double* clock_init, Residuals& r) {
    const size_t S = m.size();
    std::vector<double> sat_pos(3 * S), sat_vel(3 * S), sat_clk(3 * S), obs(3 * S), obs_std(3 * S), ttx(2 * S);
    r.sys.resize(S); r.unit.assign(3 * S, 0.0); r.res_pos.assign(S, 0.0); r.res_vel.assign(S, 0.0);
    r.sig_psr.assign(S, 0.0); r.sig_dopp.assign(S, 0.0);
    for (size_t i = 0; i < S; ++i) {
      for (int k = 0; k < 3; ++k) { sat_pos[3 * i + k] = m[i].pos[k]; sat_vel[3 * i + k] = m[i].vel[k]; }
      sat_clk[3 * i] = m[i].dt; sat_clk[3 * i + 1] = m[i].ddt; sat_clk[3 * i + 2] = m[i].tgd;
      obs[3 * i] = m[i].psr; obs[3 * i + 1] = m[i].dopp; obs[3 * i + 2] = m[i].freq;
      obs_std[3 * i] = m[i].ura; obs_std[3 * i + 1] = m[i].psr_std; obs_std[3 * i + 2] = m[i].dopp_std;
      ttx[2 * i] = m[i].ttx_doy; ttx[2 * i + 1] = m[i].ttx_sow;
      r.sys[i] = m[i].sys;
    }
    double T12[12];
    al->getTenu2ecef(T12);
    igv_gnss_res_args a{};
    a.n_sats = (int)S;
    a.sat_pos = sat_pos.data(); a.sat_vel = sat_vel.data(); a.sat_clk = sat_clk.data(); a.obs = obs.data();
    a.obs_std = obs_std.data(); a.ttx = ttx.data(); a.sys = r.sys.data(); a.T_enu2ecef = T12; a.iono = iono.data();
    a.psr_noise_amp = _psr_noise_amp; a.dopp_noise_amp = _dopp_noise_amp;
    a.unit = r.unit.data(); a.res_pos = r.res_pos.data(); a.res_vel = r.res_vel.data();
    a.sigma_psr = r.sig_psr.data(); a.sigma_dopp = r.sig_dopp.data();
    a.clock_init = clock_init;
    StateManager::check(state, igv_gnss_residuals(StateManager::handle(state), &a));
  }
};

}  // namespace ingvio
