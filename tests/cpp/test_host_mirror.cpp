// C++ tests of the host mirror (ingvio_b200/host/ingvio_host.hpp) on top of libingvio_b200.so.
// Each case restates a gtest of the reference (file:line in the comment) with its tolerance; expected
// values are recomputed here from dense closed forms, exactly as the reference's tests do.
// Needs a CUDA device. Built and run by tests/test_cpp_host_mirror.py.
#include <cmath>
#include <cstdio>
#include <functional>

#include "../../ingvio_b200/host/ingvio_host.hpp"

using namespace ingvio;

static unsigned long long g_seed = 88172645463325252ull;
static double urand() {  // xorshift, uniform in [-1, 1]
  g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17;
  return 2.0 * ((g_seed >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}
static Matrix randm(int r, int c) { Matrix M(r, c); for (double& x : M.a) x = urand(); return M; }
static Matrix mul(const Matrix& A, const Matrix& B) {
  Matrix C(A.rows(), B.cols());
  for (int i = 0; i < A.rows(); ++i) for (int k = 0; k < A.cols(); ++k) for (int j = 0; j < B.cols(); ++j) C(i, j) += A(i, k) * B(k, j);
  return C;
}
static Matrix tr(const Matrix& A) { Matrix T(A.cols(), A.rows()); for (int i = 0; i < A.rows(); ++i) for (int j = 0; j < A.cols(); ++j) T(j, i) = A(i, j); return T; }
static Matrix inv(Matrix A) {
  const int n = A.rows();
  Matrix I = Matrix::Identity(n);
  for (int j = 0; j < n; ++j) {
    int p = j;
    for (int i = j + 1; i < n; ++i) if (std::fabs(A(i, j)) > std::fabs(A(p, j))) p = i;
    for (int k = 0; k < n; ++k) { std::swap(A(j, k), A(p, k)); std::swap(I(j, k), I(p, k)); }
    const double d = 1.0 / A(j, j);
    for (int k = 0; k < n; ++k) { A(j, k) *= d; I(j, k) *= d; }
    for (int i = 0; i < n; ++i) if (i != j) { const double f = A(i, j); for (int k = 0; k < n; ++k) { A(i, k) -= f * A(j, k); I(i, k) -= f * I(j, k); } }
  }
  return I;
}
static double dist(const Matrix& A, const Matrix& B) { double s = 0; for (size_t i = 0; i < A.a.size(); ++i) s += (A.a[i] - B.a[i]) * (A.a[i] - B.a[i]); return std::sqrt(s); }
static double nrm(const Matrix& A) { double s = 0; for (double x : A.a) s += x * x; return std::sqrt(s); }
static Mat3 rand_rot() {
  double q[4], n = 0; for (double& x : q) { x = urand(); n += x * x; } n = std::sqrt(n); for (double& x : q) x /= n;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  Mat3 R; const double m[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                               2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
  for (int i = 0; i < 9; ++i) R.m[i] = m[i];
  return R;
}
static Vec3d rand_v() { Vec3d v; for (int i = 0; i < 3; ++i) v[i] = urand(); return v; }

static int g_fail = 0;
#define CHECK(cond, ...)                                                   \
  do {                                                                     \
    if (!(cond)) { ++g_fail; std::printf("  FAILED %s:%d: %s | ", __FILE__, __LINE__, #cond); std::printf(__VA_ARGS__); std::printf("\n"); } \
  } while (0)
static void run(const char* name, const std::function<void()>& f) {
  const int before = g_fail;
  f();
  std::printf("[%s] %s\n", g_fail == before ? "  OK  " : "FAILED", name);
}

static std::shared_ptr<State> fixture(Matrix* Phi, Matrix* G) {  // TestStateManager.cpp:160-187
  StateParams p;
  p._max_sw_poses = 8;
  p._init_cov_vel = 0.25; p._init_cov_bg = 0.01; p._init_cov_ba = 0.01; p._init_cov_ext_rot = 1.8e-2; p._init_cov_ext_pos = 2e-3;
  p._noise_g = 0.004; p._noise_a = 0.08; p._noise_bg = 2e-4; p._noise_ba = 8e-3; p._noise_clockbias = 0.2; p._noise_cb_rw = 0.2;
  p._T_cl2i_R = rand_rot(); p._T_cl2i_p = rand_v();
  auto state = std::make_shared<State>(p);
  state->initStateAndCov(0.0, rand_rot(), rand_v(), rand_v(), rand_v(), rand_v());
  StateManager::addGNSSVariable(state, State::GPS, 20.0, 4.0);
  StateManager::addGNSSVariable(state, State::YOF, 123.0, 1.0);
  StateManager::addGNSSVariable(state, State::FS, 2.0, 1.0);
  StateManager::addGNSSVariable(state, State::BDS, 16.0, 4.0);
  if (Phi) *Phi = randm(15, 15);
  if (G) *G = randm(15, 12);
  return state;
}

int main() {
  run("chi2_quantile == boost/scipy quantile (Update.cpp:27-34)", [] {
    const int dofs[] = {1, 2, 3, 10, 19, 66, 150};
    const double ref[] = {3.841458820694124, 5.991464547107979, 7.814727903251179, 18.307038053275146, 30.14352720564616,
                          85.96490744123096, 179.58063415418053};
    for (int i = 0; i < 7; ++i) CHECK(std::fabs(chi2_quantile(dofs[i], 0.95) - ref[i]) < 1e-9 * ref[i], "dof %d: %.15g", dofs[i], chi2_quantile(dofs[i], 0.95));
    CHECK(std::fabs(chi2_quantile(7, 0.99) - 18.475306906582357) < 1e-9 * 18.5, "0.99 quantile");
  });

  run("testState.StateAddMargProp (TestStateManager.cpp:53-156, 1e-10)", [] {
    StateParams p;
    p._noise_clockbias = 0.2;  // value StateParams holds after State.cpp:51-52
    auto state = std::make_shared<State>(p);
    CHECK(StateManager::checkStateContinuity(state) && state->curr_cov_size() == 21, "base dim %d", state->curr_cov_size());
    StateManager::addGNSSVariable(state, State::GPS, 20.0, 4.0);
    CHECK(state->curr_cov_size() == 22 && state->curr_err_variable_size() == 5, "after GPS");
    StateManager::addGNSSVariable(state, State::BDS, 16.0, 4.0);
    StateManager::addGNSSVariable(state, State::YOF, 123.0, 1.0);
    StateManager::addGNSSVariable(state, State::FS, 2.0, 1.0);
    CHECK(state->curr_cov_size() == 25 && state->curr_err_variable_size() == 8, "after 4 adds");
    StateManager::margGNSSVariable(state, State::GPS);
    CHECK(state->curr_cov_size() == 24 && state->curr_err_variable_size() == 7, "after marg");
    StateManager::addGNSSVariable(state, State::GLO, 3.0, 6.0);
    CHECK(state->curr_cov_size() == 25 && StateManager::checkStateContinuity(state), "after GLO");
    CHECK(state->_gnss.at(State::BDS)->idx() == 21 && state->_gnss.at(State::YOF)->idx() == 22 && state->_gnss.at(State::FS)->idx() == 23 &&
              state->_gnss.at(State::GLO)->idx() == 24, "insertion order");
    const Matrix cov = StateManager::getFullCov(state);
    const Matrix Phi_imu = randm(15, 15), G_imu = randm(15, 12);
    const double dt = 1.5;
    Matrix Q(14, 14);
    const double s[4] = {p._noise_g, p._noise_a, p._noise_bg, p._noise_ba};
    for (int i = 0; i < 12; ++i) Q(i, i) = s[i / 3] * s[i / 3];
    Q(12, 12) = p._noise_clockbias * p._noise_clockbias; Q(13, 13) = p._noise_cb_rw * p._noise_cb_rw;
    Matrix Phi = Matrix::Identity(25);
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) Phi(i, j) = Phi_imu(i, j);
    Phi(21, 23) = dt; Phi(24, 23) = dt;
    Matrix G(25, 14);
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 12; ++j) G(i, j) = G_imu(i, j);
    G(21, 12) = 1; G(23, 13) = 1; G(24, 12) = 1;
    Matrix ref = mul(mul(Phi, cov), tr(Phi));
    const Matrix PG = mul(Phi, G);
    const Matrix add = mul(mul(PG, Q), tr(PG));
    for (size_t i = 0; i < ref.a.size(); ++i) ref.a[i] += dt * add.a[i];
    StateManager::propagateStateCov(state, Phi_imu, G_imu, dt);
    const double e = dist(StateManager::getFullCov(state), ref);
    CHECK(e < 1e-10 * std::max(1.0, nrm(ref)), "propagateStateCov err %.3e", e);
    const Matrix small = StateManager::getMarginalCov(state, {state->_extended_pose, state->_ba, state->_gnss.at(State::YOF), state->_gnss.at(State::GLO)});
    const Matrix full = StateManager::getFullCov(state);
    CHECK(small.rows() == 14 && std::fabs(small(13, 12) - full(24, 22)) < 1e-12 && std::fabs(small(10, 2) - full(13, 2)) < 1e-12, "getMarginalCov");
  });

  run("StateUpdateTest.augmentPose (TestStateManager.cpp:195-255, 1e-8)", [] {
    Matrix Phi, G;
    auto state = fixture(&Phi, &G);
    state->_timestamp = 1.0;
    const double dts[2] = {1.5, 0.5}, ts[2] = {2.5, 3.0};
    for (int rep = 0; rep < 2; ++rep) {
      StateManager::propagateStateCov(state, Phi, G, dts[rep]);
      state->_timestamp = ts[rep];
      const Matrix cov1 = StateManager::getFullCov(state);
      const int n = cov1.rows();
      Matrix J(n + 6, n);
      for (int i = 0; i < n; ++i) J(i, i) = 1.0;
      for (int i = 0; i < 6; ++i) J(n + i, i) = 1.0;
      const Mat3& C = state->_extended_pose->valueLinearAsMat();
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { J(n + i, 15 + j) = C(i, j); J(n + 3 + i, 18 + j) = C(i, j); }
      StateManager::augmentSlidingWindowPose(state);
      const Matrix ref = mul(mul(J, cov1), tr(J));
      const double e = dist(StateManager::getFullCov(state), ref);
      CHECK(e < 1e-8 * std::max(1.0, nrm(ref)), "augment err %.3e", e);
      CHECK(state->curr_cov_size() == n + 6 && StateManager::checkStateContinuity(state), "dims");
      auto cl = state->_sw_camleft_poses.at(ts[rep]);
      const Mat3 Rc = C * state->_camleft_imu_extrinsics->valueLinearAsMat();
      double er = 0; for (int i = 0; i < 9; ++i) er += std::fabs(Rc.m[i] - cl->valueLinearAsMat().m[i]);
      CHECK(er < 1e-8, "clone rotation");
    }
    StateManager::margSlidingWindowPose(state, 2.5);
    CHECK(state->curr_cov_size() == 31 && state->_sw_camleft_poses.begin()->second->idx() == 25 && StateManager::checkStateContinuity(state), "marg clone");
  });

  run("StateUpdateTest.stateBoxPlus (TestStateManager.cpp:396-455, 1e-8)", [] {
    auto state = fixture(nullptr, nullptr);
    state->_timestamp = 2.5;
    StateManager::augmentSlidingWindowPose(state);
    const int n = state->curr_cov_size();
    Vector dx(n);
    for (double& x : dx.a) x = urand();
    const Mat3 R0 = state->_extended_pose->valueLinearAsMat();
    const Vec3d p0 = state->_extended_pose->valueTrans1(), v0 = state->_extended_pose->valueTrans2(), bg0 = state->_bg->value();
    auto cl = state->_sw_camleft_poses.at(2.5);
    const Mat3 Rc0 = cl->valueLinearAsMat(); const Vec3d pc0 = cl->valueTrans();
    const double gps0 = state->_gnss.at(State::GPS)->value();
    StateManager::boxPlus(state, dx);
    Vec3d th, d1, d2; for (int i = 0; i < 3; ++i) { th[i] = dx(i); d1[i] = dx(3 + i); d2[i] = dx(6 + i); }
    const Mat3 G0 = GammaFunc(th, 0), G1 = GammaFunc(th, 1);
    const Mat3 Rn = G0 * R0; const Vec3d pn = G0 * p0 + G1 * d1, vn = G0 * v0 + G1 * d2;
    double e = 0;
    for (int i = 0; i < 9; ++i) e += std::fabs(Rn.m[i] - state->_extended_pose->valueLinearAsMat().m[i]);
    for (int i = 0; i < 3; ++i) e += std::fabs(pn[i] - state->_extended_pose->valueTrans1()[i]) + std::fabs(vn[i] - state->_extended_pose->valueTrans2()[i]) +
                                     std::fabs(bg0[i] + dx(9 + i) - state->_bg->value()[i]);
    CHECK(e < 1e-8, "SE23/bg retraction %.3e", e);
    const int ci = cl->idx();
    for (int i = 0; i < 3; ++i) { th[i] = dx(ci + i); d1[i] = dx(ci + 3 + i); }
    const Mat3 Gc0 = GammaFunc(th, 0); const Mat3 Rcn = Gc0 * Rc0; const Vec3d pcn = Gc0 * pc0 + GammaFunc(th, 1) * d1;
    e = 0;
    for (int i = 0; i < 9; ++i) e += std::fabs(Rcn.m[i] - cl->valueLinearAsMat().m[i]);
    for (int i = 0; i < 3; ++i) e += std::fabs(pcn[i] - cl->valueTrans()[i]);
    CHECK(e < 1e-8, "clone retraction %.3e", e);
    CHECK(std::fabs(state->_gnss.at(State::GPS)->value() - gps0 - dx(state->_gnss.at(State::GPS)->idx())) < 1e-12, "scalar");
  });

  run("StateUpdateTest.stateCovUpdate (TestStateManager.cpp:478-557, 1e-8) + whitenResidual (Update.cpp:36-56)", [] {
    Matrix Phi, G;
    auto state = fixture(&Phi, &G);
    state->_timestamp = 1.0;
    StateManager::propagateStateCov(state, Phi, G, 1.5);
    state->_timestamp = 2.5;
    StateManager::augmentSlidingWindowPose(state);
    std::vector<std::shared_ptr<Type>> var_order = {state->_extended_pose, state->_gnss.at(State::GPS), state->_gnss.at(State::BDS), state->_gnss.at(State::FS)};
    CHECK(StateManager::calcSubVarSize(var_order) == 12 && StateManager::checkSubOrder(state, var_order), "var_order");
    Vector res(6); for (double& x : res.a) x = urand();
    Matrix H(6, 12);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) H(i, j) = urand();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { H(i, 3 + j) = urand(); H(3 + i, 6 + j) = urand(); }
    H(0, 9) = H(1, 9) = H(2, 10) = 1.0; for (int i = 3; i < 6; ++i) H(i, 11) = 1.0;
    const int n = state->curr_cov_size();
    Matrix HL(6, n);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 9; ++j) HL(i, j) = H(i, j);
    HL(0, state->_gnss.at(State::GPS)->idx()) = HL(1, state->_gnss.at(State::GPS)->idx()) = 1.0;
    HL(2, state->_gnss.at(State::BDS)->idx()) = 1.0;
    for (int i = 3; i < 6; ++i) HL(i, state->_gnss.at(State::FS)->idx()) = 1.0;
    const Matrix P0 = StateManager::getFullCov(state);
    Matrix R = Matrix::Identity(6); for (double& x : R.a) x *= 0.5;
    Matrix S = mul(mul(HL, P0), tr(HL)); for (int i = 0; i < 6; ++i) S(i, i) += 0.5;
    const Matrix Si = inv(S);
    // whitenResidual with sigma^2 = 0.5
    UpdateBase ub(20, 0.95);
    double gam_ref = 0; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) gam_ref += res(i) * Si(i, j) * res(j);
    const double gam = ub.whitenResidual(state, res, H, var_order, std::sqrt(0.5));
    CHECK(std::fabs(gam - gam_ref) < 1e-9 * std::max(1.0, gam_ref), "gamma %.12g vs %.12g", gam, gam_ref);
    const Vec3d p_before = state->_extended_pose->valueTrans1();
    StateManager::ekfUpdate(state, var_order, H, res, R);
    const Matrix K = mul(mul(P0, tr(HL)), Si);
    Matrix ref = mul(K, HL);
    for (auto& x : ref.a) x = -x;
    for (int i = 0; i < n; ++i) ref(i, i) += 1.0;
    ref = mul(ref, P0);
    const double e = dist(StateManager::getFullCov(state), ref);
    CHECK(e < 1e-8 * std::max(1.0, nrm(ref)), "ekfUpdate err %.3e", e);
    double moved = 0; for (int i = 0; i < 3; ++i) moved += std::fabs(state->_extended_pose->valueTrans1()[i] - p_before[i]);
    CHECK(moved > 0, "boxPlus applied to the host Type objects");
  });

  run("AddDelayedTest.addVar (TestStateManager.cpp:681-724, 1e-8)", [] {
    Matrix Phi, G;
    auto state = fixture(&Phi, &G);
    StateManager::propagateStateCov(state, Phi, G, 1.5);
    StateManager::margGNSSVariable(state, State::GPS);
    UpdateBase ub(40, 0.95);
    ub.upload(state);
    const int n = state->curr_cov_size(), m = 5;
    std::vector<std::shared_ptr<Type>> order = {state->_extended_pose, state->_gnss.at(State::YOF)};
    Matrix Hx = randm(m, 10), Hf(m, 1, 1.0);
    Vector res(m); for (double& x : res.a) x = 0.01 * urand();
    const Matrix P0 = StateManager::getFullCov(state);
    // dense reference: Householder split of Hf by hand
    Matrix Q = Matrix::Identity(m);
    {  // Q = I - 2 v v^T / (v^T v) with v = Hf - |Hf| e0 ... any orthogonal Q with Q^T Hf = [rho;0] works
      Vector v(m); const double nh = std::sqrt((double)m);
      for (int i = 0; i < m; ++i) v(i) = 1.0; v(0) += nh;
      double vv = 0; for (int i = 0; i < m; ++i) vv += v(i) * v(i);
      for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Q(i, j) -= 2.0 * v(i) * v(j) / vv;
    }
    const Matrix HxQ = mul(tr(Q), Hx);
    Matrix HfQ = mul(tr(Q), Hf);
    Vector rQ(m); for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) rQ(i) += Q(j, i) * res(j);
    Matrix HL(m, n);
    for (int i = 0; i < m; ++i) { for (int j = 0; j < 9; ++j) HL(i, j) = HxQ(i, j); HL(i, state->_gnss.at(State::YOF)->idx()) = HxQ(i, 9); }
    const double noise = 0.3, rho = HfQ(0, 0);
    Matrix Pa(n + 1, n + 1);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Pa(i, j) = P0(i, j);
    Matrix h0(1, n); for (int j = 0; j < n; ++j) h0(0, j) = HL(0, j);
    const Matrix Ph = mul(P0, tr(h0));
    double s = noise * noise; for (int j = 0; j < n; ++j) s += h0(0, j) * Ph(j, 0);
    for (int i = 0; i < n; ++i) { Pa(i, n) = -Ph(i, 0) / rho; Pa(n, i) = Pa(i, n); }
    Pa(n, n) = s / (rho * rho);
    Matrix Hup(m - 1, n + 1);
    for (int i = 1; i < m; ++i) for (int j = 0; j < n; ++j) Hup(i - 1, j) = HL(i, j);
    Matrix Su = mul(mul(Hup, Pa), tr(Hup)); for (int i = 0; i < m - 1; ++i) Su(i, i) += noise * noise;
    const Matrix K = mul(mul(Pa, tr(Hup)), inv(Su));
    Matrix ref = mul(K, Hup); for (auto& x : ref.a) x = -x; for (int i = 0; i <= n; ++i) ref(i, i) += 1.0; ref = mul(ref, Pa);
    auto cb = std::make_shared<Scalar>(); cb->setValue(3.0);
    Matrix Hx_in = Hx, Hf_in = Hf; Vector res_in = res;
    const bool ok = StateManager::addVariableDelayed(state, cb, order, Hx_in, Hf_in, res_in, noise, 1e6, true, State::GPS);
    CHECK(ok && state->curr_cov_size() == n + 1 && cb->idx() == n, "accepted");
    const double e = dist(StateManager::getFullCov(state), ref);
    CHECK(e < 1e-8 * std::max(1.0, nrm(ref)), "addVariableDelayed err %.3e", e);
    // chi^2 failure leaves the state untouched (StateManager.cpp:617-621)
    auto cb2 = std::make_shared<Scalar>();
    Vector bad(m); for (int i = 0; i < m; ++i) bad(i) = 50.0 * (i - 2);
    Matrix Hx2 = Hx, Hf2 = Hf;
    const Matrix Pb = StateManager::getFullCov(state);
    const bool ok2 = StateManager::addVariableDelayed(state, cb2, order, Hx2, Hf2, bad, noise, 0.95, true, State::GAL);
    CHECK(!ok2 && state->curr_cov_size() == n + 1 && dist(StateManager::getFullCov(state), Pb) == 0.0, "rejection must not change the state");
  });

  std::printf("%s (%d failed checks)\n", g_fail ? "SOME TESTS FAILED" : "ALL TESTS PASSED", g_fail);
  return g_fail ? 1 : 0;
}
