// ingvio_host.hpp -- C++ host mirror of the reference's filter-core interface on top of the C-ABI.
//
// The reference's seam is C++: `State` owns the Type objects, `StateManager` (all static, friend of
// State) owns every touch of the covariance (ingvio_estimator/src/State.h:72-136,
// StateManager.h:33-128). This header keeps those class names, method names, argument meaning and
// error behaviour, with the covariance living in an igv_batch (B = 1) instead of an Eigen::MatrixXd,
// so the body of each StateManager method is exactly the stub INTEGRATION.md describes.
// The reference passes Eigen types; Eigen is not available here, so `Matrix` / `Vector` below are
// minimal column-major stand-ins with the same data layout as Eigen::MatrixXd / VectorXd
// (a maintainer passes `M.data()` of the Eigen objects instead).
//
// Header-only, plain C++17 (g++), links against libingvio_b200.so only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ingvio_b200.h"

namespace ingvio {

// ---- minimal dense types (column-major like Eigen) ----------------------------------------------
struct Vector {
  std::vector<double> a;
  Vector() {}
  explicit Vector(int n, double v = 0.0) : a(n, v) {}
  int rows() const { return (int)a.size(); }
  double& operator()(int i) { return a[i]; }
  double operator()(int i) const { return a[i]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
};
struct Matrix {
  int r = 0, c = 0;
  std::vector<double> a;
  Matrix() {}
  Matrix(int r_, int c_, double v = 0.0) : r(r_), c(c_), a((size_t)r_ * c_, v) {}
  static Matrix Identity(int n) { Matrix I(n, n); for (int i = 0; i < n; ++i) I(i, i) = 1.0; return I; }
  int rows() const { return r; }
  int cols() const { return c; }
  double& operator()(int i, int j) { return a[i + (size_t)j * r]; }
  double operator()(int i, int j) const { return a[i + (size_t)j * r]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
};
struct Mat3 {  // row-major 3x3
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& operator()(int i, int j) { return m[3 * i + j]; }
  double operator()(int i, int j) const { return m[3 * i + j]; }
};
struct Vec3d { double v[3] = {0, 0, 0}; double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };

inline Mat3 operator*(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return C;
}
inline Vec3d operator*(const Mat3& A, const Vec3d& x) {
  Vec3d y;
  for (int i = 0; i < 3; ++i) y[i] = A(i, 0) * x[0] + A(i, 1) * x[1] + A(i, 2) * x[2];
  return y;
}
inline Vec3d operator+(const Vec3d& a, const Vec3d& b) { Vec3d c; for (int i = 0; i < 3; ++i) c[i] = a[i] + b[i]; return c; }

// AuxGammaFunc.cpp:28-35, :46-113
inline Mat3 skew(const Vec3d& v) { Mat3 S; const double s[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}; for (int i = 0; i < 9; ++i) S.m[i] = s[i]; return S; }
inline Mat3 GammaFunc(const Vec3d& vec, int m) {
  const double th = std::sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  Mat3 out;
  if (std::fabs(th) < 1e-6) {
    const double f = (m == 3) ? 1.0 / 6.0 : (m == 2 ? 0.5 : 1.0);
    for (int i = 0; i < 9; ++i) out.m[i] = 0.0;
    out(0, 0) = out(1, 1) = out(2, 2) = f;
    return out;
  }
  Vec3d n; for (int i = 0; i < 3; ++i) n[i] = vec[i] / th;
  const Mat3 nx = skew(n), nx2 = nx * nx;
  const double s = std::sin(th), c = std::cos(th);
  double f0, f1, f2;
  if (m == 1) { f0 = 1; f1 = (1 - c) / th; f2 = (th - s) / th; }
  else if (m == 2) { f0 = 0.5; f1 = (th - s) / (th * th); f2 = (th * th + 2 * c - 2) / (2 * th * th); }
  else if (m == 3) { const double t3 = th * th * th; f0 = 1.0 / 6; f1 = (th * th + 2 * c - 2) / (2 * t3); f2 = (t3 - 6 * th + 6 * s) / (6 * t3); }
  else { f0 = 1; f1 = s; f2 = 1 - c; }
  for (int i = 0; i < 9; ++i) out.m[i] = f1 * nx.m[i] + f2 * nx2.m[i];
  out(0, 0) += f0; out(1, 1) += f0; out(2, 2) += f0;
  return out;
}

// ---- Type hierarchy (VecState.h:32-134, PoseState.h:30-230) ----------------------------------------
class Type {
 public:
  explicit Type(int size) : _size(size) {}
  virtual ~Type() {}
  int idx() const { return _idx; }
  int size() const { return _size; }
  void set_cov_idx(int i) { _idx = i; }
  virtual void update(const Vector& dx) = 0;
 protected:
  int _idx = -1, _size;
};
class Vec3 : public Type {
 public:
  Vec3() : Type(3) {}
  const Vec3d& value() const { return _vec; }
  void setValue(const Vec3d& v) { _vec = v; }
  void update(const Vector& dx) override { for (int i = 0; i < 3; ++i) _vec[i] += dx(_idx + i); }  // VecState.cpp:27-31
 private:
  Vec3d _vec;
};
class Scalar : public Type {
 public:
  Scalar() : Type(1) {}
  double value() const { return _scalar; }
  void setValue(double v) { _scalar = v; }
  void update(const Vector& dx) override { _scalar += dx(_idx); }  // VecState.cpp:43-47
 private:
  double _scalar = 0.0;
};
class SE3 : public Type {
 public:
  SE3() : Type(6) {}
  const Mat3& valueLinearAsMat() const { return _rot; }
  const Vec3d& valueTrans() const { return _vec; }
  void setValueLinearByMat(const Mat3& R) { _rot = R; }
  void setValueTrans(const Vec3d& p) { _vec = p; }
  void update(const Vector& dx) override {  // PoseState.cpp:79-88
    Vec3d th, dp;
    for (int i = 0; i < 3; ++i) { th[i] = dx(_idx + i); dp[i] = dx(_idx + 3 + i); }
    const Mat3 G0 = GammaFunc(th, 0);
    _rot = G0 * _rot;
    _vec = G0 * _vec + GammaFunc(th, 1) * dp;
  }
 private:
  Mat3 _rot; Vec3d _vec;
};
class SE23 : public Type {
 public:
  SE23() : Type(9) {}
  const Mat3& valueLinearAsMat() const { return _rot; }
  const Vec3d& valueTrans1() const { return _vec1; }
  const Vec3d& valueTrans2() const { return _vec2; }
  void setValueLinearByMat(const Mat3& R) { _rot = R; }
  void setValueTrans1(const Vec3d& p) { _vec1 = p; }
  void setValueTrans2(const Vec3d& v) { _vec2 = v; }
  void update(const Vector& dx) override {  // PoseState.cpp:174-186
    Vec3d th, d1, d2;
    for (int i = 0; i < 3; ++i) { th[i] = dx(_idx + i); d1[i] = dx(_idx + 3 + i); d2[i] = dx(_idx + 6 + i); }
    const Mat3 G0 = GammaFunc(th, 0), G1 = GammaFunc(th, 1);
    _rot = G0 * _rot;
    _vec1 = G0 * _vec1 + G1 * d1;
    _vec2 = G0 * _vec2 + G1 * d2;
  }
 private:
  Mat3 _rot; Vec3d _vec1, _vec2;
};

// ---- chi^2 quantile without Boost (replaces boost::math::quantile(chi_squared), Update.cpp:27-34) ------
namespace detail {
inline double gamma_p(double a, double x) {  // regularised lower incomplete gamma P(a,x)
  if (x <= 0) return 0.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {  // series
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 1000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (std::fabs(del) < std::fabs(sum) * 1e-16) break; }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;  // continued fraction (Lentz)
  for (int i = 1; i < 1000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c; if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}
}  // namespace detail
inline double chi2_quantile(int dof, double p) {
  const double a = 0.5 * dof;
  // Wilson-Hilferty start, then Newton on P(a, x/2) = p with bisection safeguards
  const double z = [&] {  // inverse normal CDF (Acklam), enough for a starting point
    const double q = p - 0.5;
    if (std::fabs(q) < 0.42) { const double r = q * q; return q * (((-25.44106049637 * r + 41.39119773534) * r - 18.61500062529) * r + 2.50662823884) / ((((3.13082909833 * r - 21.06224101826) * r + 23.08336743743) * r - 8.47351093090) * r + 1.0); }
    double r = p < 0.5 ? p : 1 - p; r = std::log(-std::log(r));
    double x = 0.3374754822726147 + r * (0.9761690190917186 + r * (0.1607979714918209 + r * (0.0276438810333863 + r * (0.0038405729373609 + r * (0.0003951896511919 + r * (0.0000321767881768 + r * (0.0000002888167364 + r * 0.0000003960315187)))))));
    return p < 0.5 ? -x : x;
  }();
  double x = dof * std::pow(1.0 - 2.0 / (9.0 * dof) + z * std::sqrt(2.0 / (9.0 * dof)), 3.0);
  if (!(x > 0)) x = 0.5 * dof;
  double lo = 0.0, hi = std::max(4.0 * dof + 50.0, 2 * x);
  for (int it = 0; it < 200; ++it) {
    const double f = detail::gamma_p(a, 0.5 * x) - p;
    if (f > 0) hi = x; else lo = x;
    const double pdf = std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a)) * 0.5;
    double xn = x - f / pdf;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-14 * std::max(1.0, x)) { x = xn; break; }
    x = xn;
  }
  return x;
}

// ---- State (State.h:36-136, State.cpp:25-167) --------------------------------------------------------
struct StateParams {
  int _cam_nums = 2, _max_sw_poses = 20;
  double _noise_g = 0.005, _noise_a = 0.05, _noise_bg = 0.001, _noise_ba = 0.01, _noise_clockbias = 2.0, _noise_cb_rw = 0.2;
  double _init_cov_rot = 0, _init_cov_pos = 0, _init_cov_vel = 0, _init_cov_bg = 0, _init_cov_ba = 0;
  double _init_cov_ext_rot = 0, _init_cov_ext_pos = 0;
  bool _enable_gnss = true;
  Mat3 _T_cl2i_R; Vec3d _T_cl2i_p;
  Mat3 _T_cl2cr_R; Vec3d _T_cl2cr_p;                   // left -> right camera (State.cpp:33), identity / zero by default
  double gravity[3] = {0, 0, -9.8};
};

class StateManager;
class State {
 public:
  enum GNSSType { GPS = 0, GLO, GAL, BDS, FS, YOF };
  explicit State(const StateParams& p, int max_feats = 400, int max_sats = 40) : _state_params(p), _max_feats(max_feats) {
    igv_config cfg{1, 21 + 6 + 6 * (p._max_sw_poses + 1), p._max_sw_poses + 1, max_feats, max_sats, p._cam_nums == 2, 0, nullptr};
    if (igv_create(&cfg, &_gpu) != IGV_OK) throw std::runtime_error("igv_create failed (no CUDA device / library?)");
    igv_params ip{};
    ip.noise_g = p._noise_g; ip.noise_a = p._noise_a; ip.noise_bg = p._noise_bg; ip.noise_ba = p._noise_ba;
    ip.noise_clockbias = p._noise_clockbias; ip.noise_cb_rw = p._noise_cb_rw;
    for (int i = 0; i < 3; ++i) ip.gravity[i] = p.gravity[i];
    for (int i = 0; i < 9; ++i) ip.T_cl2cr_R[i] = p._T_cl2cr_R.m[i];
    for (int i = 0; i < 3; ++i) ip.T_cl2cr_p[i] = p._T_cl2cr_p[i];
    igv_set_params(_gpu, &ip);
    _extended_pose = std::make_shared<SE23>(); _bg = std::make_shared<Vec3>(); _ba = std::make_shared<Vec3>();
    _camleft_imu_extrinsics = std::make_shared<SE3>();
    int idx = 0;
    for (auto v : std::vector<std::shared_ptr<Type>>{_extended_pose, _bg, _ba, _camleft_imu_extrinsics}) {
      v->set_cov_idx(idx); _err_variables.push_back(v); idx += v->size();
    }
    _camleft_imu_extrinsics->setValueLinearByMat(p._T_cl2i_R);
    _camleft_imu_extrinsics->setValueTrans(p._T_cl2i_p);
    // State.cpp:88: cov = 1e-6 I until initStateAndCov overwrites the diagonal
    double diag[21]; for (double& d : diag) d = 1e-6;
    push_mean(diag);
  }
  ~State() { igv_destroy(_gpu); }
  State(const State&) = delete;

  // State.cpp:126-167
  void initStateAndCov(double t0, const Mat3& R_i2w, const Vec3d& pos, const Vec3d& vel, const Vec3d& bg, const Vec3d& ba) {
    _timestamp = t0;
    const StateParams& s = _state_params;
    const double sd[7] = {s._init_cov_rot, s._init_cov_pos, s._init_cov_vel, s._init_cov_bg, s._init_cov_ba, s._init_cov_ext_rot, s._init_cov_ext_pos};
    double diag[21];
    for (int i = 0; i < 21; ++i) diag[i] = sd[i / 3] * sd[i / 3];
    _extended_pose->setValueLinearByMat(R_i2w); _extended_pose->setValueTrans1(pos); _extended_pose->setValueTrans2(vel);
    _bg->setValue(bg); _ba->setValue(ba);
    _camleft_imu_extrinsics->setValueLinearByMat(s._T_cl2i_R); _camleft_imu_extrinsics->setValueTrans(s._T_cl2i_p);
    push_mean(diag);
  }
  int curr_cov_size() const { return igv_dim(_gpu); }
  int curr_err_variable_size() const { return (int)_err_variables.size(); }
  double nextMargTime() const {  // State.h:85-95
    double t = std::numeric_limits<double>::infinity();
    if ((int)_sw_camleft_poses.size() > _state_params._max_sw_poses) for (auto& it : _sw_camleft_poses) t = std::min(t, it.first);
    return t;
  }

  double _timestamp = -1;
  StateParams _state_params;
  int _max_feats;                                     // capacity of one fused visual update (igv_config::max_feats)
  std::shared_ptr<SE23> _extended_pose;
  std::shared_ptr<Vec3> _bg, _ba;
  std::shared_ptr<SE3> _camleft_imu_extrinsics;
  std::unordered_map<int, std::shared_ptr<Scalar>> _gnss;
  std::map<double, std::shared_ptr<SE3>> _sw_camleft_poses;

 private:
  friend class StateManager;
  friend class UpdateBase;
  void push_mean(const double* diag21) {
    double zeros3[3] = {0, 0, 0};
    (void)zeros3;
    if (igv_state_init(_gpu, _extended_pose->valueLinearAsMat().m, _extended_pose->valueTrans1().v, _extended_pose->valueTrans2().v,
                       _bg->value().v, _ba->value().v, _camleft_imu_extrinsics->valueLinearAsMat().m,
                       _camleft_imu_extrinsics->valueTrans().v, diag21) != IGV_OK)
      throw std::runtime_error(igv_last_error(_gpu));
  }
  igv_batch* _gpu = nullptr;                          // replaces Eigen::MatrixXd _cov
  std::vector<std::shared_ptr<Type>> _err_variables;
};

// ---- StateManager (StateManager.h:33-128) ---------------------------------------------------------------
class StateManager {
 public:
  StateManager() = delete;
  static void check(const std::shared_ptr<State>& s, igv_status st, bool fatal = false) {
    if (st == IGV_OK) return;
    std::printf("[StateManager]: %s\n", igv_last_error(s->_gpu));
    if (fatal) std::exit(EXIT_FAILURE);  // the reference exits at these sites (e.g. StateManager.cpp:157-161)
  }
  static bool checkStateContinuity(const std::shared_ptr<State>& state) {  // StateManager.cpp:27-40
    int idx = 0;
    for (auto& v : state->_err_variables) { if (v->idx() != idx) return false; idx += v->size(); }
    return idx == state->curr_cov_size();
  }
  // Phi 15x15, G 15x12, column-major (StateManager.cpp:42-119)
  static void propagateStateCov(std::shared_ptr<State> state, const Matrix& Phi_imu, const Matrix& G_imu, double dt) {
    check(state, igv_propagate_cov(state->_gpu, Phi_imu.data(), G_imu.data(), &dt));
  }
  static Matrix getFullCov(std::shared_ptr<State> state) {  // :121-126
    const int n = state->curr_cov_size();
    Matrix cov(n, n);
    check(state, igv_cov_get(state->_gpu, cov.data(), n));
    return cov;
  }
  static Matrix getMarginalCov(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>>& vars) {  // :128-153
    std::vector<int> idx, size; int n = 0;
    for (auto& v : vars) { idx.push_back(v->idx()); size.push_back(v->size()); n += v->size(); }
    Matrix small(n, n);
    check(state, igv_cov_get_blocks(state->_gpu, (int)idx.size(), idx.data(), size.data(), small.data()));
    return small;
  }
  static void marginalize(std::shared_ptr<State> state, std::shared_ptr<Type> marg) {  // :155-192
    auto it = std::find(state->_err_variables.begin(), state->_err_variables.end(), marg);
    if (it == state->_err_variables.end()) {
      std::printf("[StateManager]: Marg is not in the current state!\n");
      std::exit(EXIT_FAILURE);
    }
    check(state, igv_marginalize(state->_gpu, marg->idx()), true);
    const int s = marg->idx(), sz = marg->size();
    std::vector<std::shared_ptr<Type>> rem;
    for (auto& v : state->_err_variables) if (v != marg) { if (v->idx() > s) v->set_cov_idx(v->idx() - sz); rem.push_back(v); }
    marg->set_cov_idx(-1);
    state->_err_variables = rem;
  }
  // StateManager.cpp:194-214: a new variable with its own covariance block, decoupled from the rest
  static void addVariableIndependent(std::shared_ptr<State> state, std::shared_ptr<Type> new_state, const Matrix& new_state_cov_block) {
    if (std::find(state->_err_variables.begin(), state->_err_variables.end(), new_state) != state->_err_variables.end()) {
      std::printf("[StateManager]: Variable already in the state, cannot be added!\n");
      return;
    }
    new_state->set_cov_idx(state->curr_cov_size());
    check(state, igv_add_variable_independent(state->_gpu, new_state->size(), new_state_cov_block.data()));
    state->_err_variables.push_back(new_state);
  }
  // StateManager.cpp:632-693: target_var becomes a linear function H of the variables in dependence_order
  static void replaceVarLinear(std::shared_ptr<State> state, const std::shared_ptr<Type> target_var,
                               const std::vector<std::shared_ptr<Type>>& dependence_order, const Matrix& H) {
    if (std::find(state->_err_variables.begin(), state->_err_variables.end(), target_var) == state->_err_variables.end()) {
      std::printf("[StateManager]: Target var not in state, cannot linearly replace!\n");
      return;
    }
    std::vector<int> idx, size;
    for (auto& v : dependence_order) { idx.push_back(v->idx()); size.push_back(v->size()); }
    check(state, igv_replace_var_linear(state->_gpu, target_var->idx(), target_var->size(), (int)idx.size(), idx.data(), size.data(), H.data()));
  }
  // a variable the device already appended at covariance index `idx` (fused device-side additions)
  static void registerVariable(std::shared_ptr<State> state, std::shared_ptr<Type> var, int idx) {
    var->set_cov_idx(idx);
    state->_err_variables.push_back(var);
  }
  static void addGNSSVariable(std::shared_ptr<State> state, int gtype, double value, double cov) {  // :216-231
    if (state->_gnss.count(gtype)) std::printf("[StateManager]: GNSS variable already in the state, adding operation will rewrite such var!\n");
    auto s = std::make_shared<Scalar>();
    s->setValue(value);
    s->set_cov_idx(state->curr_cov_size());
    check(state, igv_add_gnss_variable(state->_gpu, gtype, &value, cov));
    state->_gnss[gtype] = s;
    state->_err_variables.push_back(s);
  }
  static void margGNSSVariable(std::shared_ptr<State> state, int gtype) {  // :233-242
    if (!state->_gnss.count(gtype)) { std::printf("[StateManager]: GNSS variable not in the state, no need to marg!\n"); return; }
    marginalize(state, state->_gnss.at(gtype));
    state->_gnss.erase(gtype);
  }
  static void boxPlus(std::shared_ptr<State> state, const Vector& dx) {  // :244-251
    for (auto& v : state->_err_variables) v->update(dx);
  }
  static void augmentSlidingWindowPose(std::shared_ptr<State> state) {  // :253-296
    if (state->_sw_camleft_poses.count(state->_timestamp)) { std::printf("[StateManager]: Curr pose already in the sw, cannot clone!\n"); return; }
    auto clone = std::make_shared<SE3>();
    const Mat3& R = state->_extended_pose->valueLinearAsMat();
    clone->setValueLinearByMat(R * state->_camleft_imu_extrinsics->valueLinearAsMat());
    clone->setValueTrans(R * state->_camleft_imu_extrinsics->valueTrans() + state->_extended_pose->valueTrans1());
    clone->set_cov_idx(state->curr_cov_size());
    check(state, igv_augment_clone_cov(state->_gpu, R.m, clone->valueLinearAsMat().m, clone->valueTrans().v));
    state->_sw_camleft_poses[state->_timestamp] = clone;
    state->_err_variables.push_back(clone);
  }
  static void margSlidingWindowPose(std::shared_ptr<State> state, double marg_time) {  // :316-326
    if (!state->_sw_camleft_poses.count(marg_time)) { std::printf("[StateManager]: Marg pose time not exists! Cannot marg!\n"); return; }
    marginalize(state, state->_sw_camleft_poses.at(marg_time));
    state->_sw_camleft_poses.erase(marg_time);
  }
  static void margSlidingWindowPose(std::shared_ptr<State> state) {  // :328-338
    const double t = state->nextMargTime();
    if (t == std::numeric_limits<double>::infinity()) { std::printf("[StateManager]: Auto marg pose gives inf time! Cannot marg!\n"); return; }
    margSlidingWindowPose(state, t);
  }
  static bool checkSubOrder(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>>& sub) {  // :428-445
    for (auto& v : sub) if (std::find(state->_err_variables.begin(), state->_err_variables.end(), v) == state->_err_variables.end()) return false;
    return true;
  }
  static int calcSubVarSize(const std::vector<std::shared_ptr<Type>>& sub) { int n = 0; for (auto& v : sub) if (v) n += v->size(); return n; }
  // H (rows x n), R (rows x rows) column-major, as Eigen passes them (StateManager.cpp:359-426)
  static void ekfUpdate(std::shared_ptr<State> state, const std::vector<std::shared_ptr<Type>>& var_order, const Matrix& H,
                        const Vector& res, const Matrix& R) {
    std::vector<int> idx, size;
    for (auto& v : var_order) { idx.push_back(v->idx()); size.push_back(v->size()); }
    Vector dx(state->curr_cov_size());
    check(state, igv_ekf_update(state->_gpu, (int)idx.size(), idx.data(), size.data(), H.rows(), H.data(), H.rows(), res.data(),
                                R.data(), IGV_R_FULL, dx.data()));
    int flags = 0;
    igv_get_flags(state->_gpu, &flags, 1);
    if (flags & IGV_FLAG_NEG_DIAG) std::printf("[StateManager]: EKF Update and found negative diag cov elements! \n");  // :413-421
    boxPlus(state, dx);
  }
  // StateManager.cpp:547-630 for a scalar new variable (every call site in scope, GnssUpdate.cpp:430,470)
  static bool addVariableDelayed(std::shared_ptr<State> state, std::shared_ptr<Scalar> var_new,
                                 const std::vector<std::shared_ptr<Type>>& var_old_order, Matrix& H_old, Matrix& H_new, Vector& res,
                                 double noise_iso_meas, double chi2_mult_factor, bool do_chi2 = true, int gtype = -1) {
    if (std::find(state->_err_variables.begin(), state->_err_variables.end(), std::static_pointer_cast<Type>(var_new)) != state->_err_variables.end()) {
      std::printf("[StateManager]: New var already in state! Cannot perform add var delayed inv!\n");
      return false;
    }
    if (H_new.rows() <= H_new.cols()) { std::printf("[StateManager]: H_new rows should be larger than H_new cols!\n"); return false; }
    std::vector<int> idx, size;
    for (auto& v : var_old_order) { idx.push_back(v->idx()); size.push_back(v->size()); }
    int accepted = 0;
    const double value = var_new->value();
    const int n0 = state->curr_cov_size();
    Vector dx(n0 + 1);
    igv_status st = igv_add_variable_delayed(state->_gpu, gtype, &value, (int)idx.size(), idx.data(), size.data(), H_old.rows(),
                                             H_old.data(), H_new.data(), res.data(), noise_iso_meas, chi2_mult_factor, do_chi2 ? 1 : 0,
                                             1.0, &accepted, dx.data());
    check(state, st);
    if (st != IGV_OK) return false;
    if (!accepted) {  // single filter: drop the decoupled placeholder again -> exactly the reference's "return false"
      check(state, igv_marginalize(state->_gpu, n0));
      std::printf("[StateManager]: Cannot add variable due to chi2 test failure!\n");
      return false;
    }
    var_new->set_cov_idx(n0);
    state->_err_variables.push_back(var_new);
    if (H_old.rows() > 1) boxPlus(state, dx);   // the residual EKF of StateManager.cpp:626-627
    return true;
  }
  // pull the mean mirror back into the Type objects (used after fused device-side updates)
  static void sync_mean_from_device(std::shared_ptr<State> state) {
    std::vector<double> x(igv_state_size(state->_gpu));
    check(state, igv_state_get(state->_gpu, x.data()));
    Mat3 R; Vec3d p, v, bg, ba;
    for (int i = 0; i < 9; ++i) R.m[i] = x[i];
    for (int i = 0; i < 3; ++i) { p[i] = x[9 + i]; v[i] = x[12 + i]; bg[i] = x[15 + i]; ba[i] = x[18 + i]; }
    state->_extended_pose->setValueLinearByMat(R); state->_extended_pose->setValueTrans1(p); state->_extended_pose->setValueTrans2(v);
    state->_bg->setValue(bg); state->_ba->setValue(ba);
    for (int i = 0; i < 9; ++i) R.m[i] = x[21 + i];
    for (int i = 0; i < 3; ++i) p[i] = x[30 + i];
    state->_camleft_imu_extrinsics->setValueLinearByMat(R); state->_camleft_imu_extrinsics->setValueTrans(p);
    for (auto& g : state->_gnss) g.second->setValue(x[33 + g.first]);
    int s = 0;
    for (auto& it : state->_sw_camleft_poses) {
      for (int i = 0; i < 9; ++i) R.m[i] = x[39 + 12 * s + i];
      for (int i = 0; i < 3; ++i) p[i] = x[39 + 12 * s + 9 + i];
      it.second->setValueLinearByMat(R); it.second->setValueTrans(p);
      ++s;
    }
  }
  static igv_batch* handle(const std::shared_ptr<State>& s) { return s->_gpu; }
};

// ---- UpdateBase (Update.h:36-97, Update.cpp:27-149) ------------------------------------------------------
class UpdateBase {
 public:
  UpdateBase(int max_dof = 150, double thres = 0.95) : _thres(thres) { setChiSquaredTable(max_dof, thres); }
  virtual ~UpdateBase() {}
  UpdateBase(const UpdateBase&) = delete;
  const std::map<int, double>& table() const { return _chi_squared_table; }
  void upload(std::shared_ptr<State> state) const {
    std::vector<double> t;
    for (auto& kv : _chi_squared_table) t.push_back(kv.second);
    igv_set_chi2_table(StateManager::handle(state), t.data(), (int)t.size());
  }
 protected:
  double _thres;
  std::map<int, double> _chi_squared_table;
  void setChiSquaredTable(int max_dof, double thres) { for (int i = 1; i <= max_dof; ++i) _chi_squared_table[i] = chi2_quantile(i, thres); }
 public:
  double whitenResidual(std::shared_ptr<State> state, const Vector& res, const Matrix& H, const std::vector<std::shared_ptr<Type>>& var_order, double noise) {
    std::vector<int> idx, size;
    for (auto& v : var_order) { idx.push_back(v->idx()); size.push_back(v->size()); }
    const double n2 = noise * noise;
    double gamma = 0.0;
    StateManager::check(state, igv_chi2_whiten(StateManager::handle(state), (int)idx.size(), idx.data(), size.data(), H.rows(), H.data(), H.rows(),
                                               res.data(), &n2, IGV_R_ISO, &gamma));
    return gamma;
  }
  bool testChiSquared(std::shared_ptr<State> state, const Vector& res, const Matrix& H, const std::vector<std::shared_ptr<Type>>& var_order, double noise, int dof) {
    const double prob = whitenResidual(state, res, H, var_order, noise);
    if (!_chi_squared_table.count(dof))
      for (int i = _chi_squared_table.rbegin()->first + 1; i <= dof; ++i) _chi_squared_table[i] = chi2_quantile(i, _thres);
    return prob < _chi_squared_table.at(dof);
  }
};

}  // namespace ingvio
