// CPU restatement (C++, single-threaded per sequence, FP64) of the InGVIO EKF hot path.
//
// TEST / MEASUREMENT INFRASTRUCTURE (oracle). This is the timed CPU baseline ("port") of bench.py and a
// second checker; it is never linked into or called from the product library. The reference itself
// cannot be compiled here (no Eigen / SuiteSparse / Boost / ROS in the image), so the
// steps are restated with a small dense linear-algebra layer, in the reference's order:
//   ImuPropagator::stateAndCovTransition           ImuPropagator.cpp:98-162
//   StateManager::propagateStateCov (full N x N temporaries, as the reference does)   StateManager.cpp:42-119
//   StateManager::augmentSlidingWindowPose / marginalize                             :253-296, :155-192
//   RemoveLostUpdate::calcResJacobianSingleFeatAll*Obs (dense zero-filled blocks)   RemoveLostUpdate.cpp:169-273,:407-523
//   UpdateBase::testChiSquared / whitenResidual                                     Update.cpp:36-124
//   stacking + QR compression (dense Householder in place of SuiteSparse SPQR)     RemoveLostUpdate.cpp:99-160
//   StateManager::ekfUpdate (block PH^T loop, explicit S^-1 by LU, full symmetrise)  StateManager.cpp:359-426
//   GnssUpdate::updateTrackedSys                                                   GnssUpdate.cpp:124-290
// Validated against the numpy oracle in tests/test_cpu_port.py.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;  // row-major
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

Mat matmul(const Mat& A, const Mat& B) {  // C = A B
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int k = 0; k < A.c; ++k) {
      const double aik = A(i, k);
      if (aik == 0.0) continue;
      const double* b = &B.a[(size_t)k * B.c];
      double* c = &C.a[(size_t)i * C.c];
      for (int j = 0; j < B.c; ++j) c[j] += aik * b[j];
    }
  return C;
}
Mat matmulT(const Mat& A, const Mat& B) {  // C = A B^T
  Mat C(A.r, B.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < B.r; ++j) {
      const double* a = &A.a[(size_t)i * A.c];
      const double* b = &B.a[(size_t)j * B.c];
      double s = 0.0;
      for (int k = 0; k < A.c; ++k) s += a[k] * b[k];
      C(i, j) = s;
    }
  return C;
}
Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; ++i) for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
  return T;
}

// In-place Householder QR of A (m x n); the same reflectors are applied to the columns of B (m x nb).
void householder_qr(Mat& A, Mat* B) {
  const int m = A.r, n = A.c;
  std::vector<double> v(m);
  for (int j = 0; j < std::min(m - 1, n); ++j) {
    double ss = 0.0;
    for (int i = j + 1; i < m; ++i) ss += A(i, j) * A(i, j);
    if (ss == 0.0) continue;
    const double alpha = A(j, j);
    const double beta = -std::copysign(std::sqrt(alpha * alpha + ss), alpha);
    const double tau = (beta - alpha) / beta, scale = 1.0 / (alpha - beta);
    v[j] = 1.0;
    for (int i = j + 1; i < m; ++i) v[i] = A(i, j) * scale;
    A(j, j) = beta;
    for (int i = j + 1; i < m; ++i) A(i, j) = 0.0;
    for (int k = j + 1; k < n; ++k) {
      double w = 0.0;
      for (int i = j; i < m; ++i) w += v[i] * A(i, k);
      w *= tau;
      for (int i = j; i < m; ++i) A(i, k) -= w * v[i];
    }
    if (B)
      for (int k = 0; k < B->c; ++k) {
        double w = 0.0;
        for (int i = j; i < m; ++i) w += v[i] * (*B)(i, k);
        w *= tau;
        for (int i = j; i < m; ++i) (*B)(i, k) -= w * v[i];
      }
  }
}

bool cholesky(Mat& S) {  // lower, in place
  const int n = S.r;
  for (int j = 0; j < n; ++j) {
    double d = S(j, j);
    for (int k = 0; k < j; ++k) d -= S(j, k) * S(j, k);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    S(j, j) = d;
    for (int i = j + 1; i < n; ++i) {
      double s = S(i, j);
      for (int k = 0; k < j; ++k) s -= S(i, k) * S(j, k);
      S(i, j) = s / d;
    }
  }
  return true;
}
double chi2_stat(Mat S, const std::vector<double>& r) {  // r^T S^-1 r (Update.cpp:55: ldlt().solve)
  if (!cholesky(S)) return NAN;
  const int n = S.r;
  std::vector<double> y(n);
  double g = 0.0;
  for (int i = 0; i < n; ++i) {
    double s = r[i];
    for (int k = 0; k < i; ++k) s -= S(i, k) * y[k];
    y[i] = s / S(i, i);
    g += y[i] * y[i];
  }
  return g;
}
Mat lu_inverse(Mat A) {  // Eigen MatrixXd::inverse() = PartialPivLU (StateManager.cpp:405)
  const int n = A.r;
  Mat I(n, n);
  for (int i = 0; i < n; ++i) I(i, i) = 1.0;
  for (int j = 0; j < n; ++j) {
    int p = j;
    for (int i = j + 1; i < n; ++i) if (std::fabs(A(i, j)) > std::fabs(A(p, j))) p = i;
    if (p != j) for (int k = 0; k < n; ++k) { std::swap(A(j, k), A(p, k)); std::swap(I(j, k), I(p, k)); }
    const double d = 1.0 / A(j, j);
    for (int i = j + 1; i < n; ++i) {
      const double f = A(i, j) * d;
      if (f == 0.0) continue;
      for (int k = j; k < n; ++k) A(i, k) -= f * A(j, k);
      for (int k = 0; k < n; ++k) I(i, k) -= f * I(j, k);
    }
  }
  for (int j = n - 1; j >= 0; --j) {
    const double d = 1.0 / A(j, j);
    for (int k = 0; k < n; ++k) I(j, k) *= d;
    for (int i = 0; i < j; ++i) {
      const double f = A(i, j);
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) I(i, k) -= f * I(j, k);
    }
  }
  return I;
}

// ---- 3x3 ------------------------------------------------------------------------------------------
struct M3 { double m[9]; };
M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
void mv(const M3& A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A.m[3 * i] * x[0] + A.m[3 * i + 1] * x[1] + A.m[3 * i + 2] * x[2]; }
void mtv(const M3& A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A.m[i] * x[0] + A.m[3 + i] * x[1] + A.m[6 + i] * x[2]; }
M3 skew(const double* v) { return M3{{0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}}; }
M3 scaled(const M3& A, double s) { M3 B = A; for (double& x : B.m) x *= s; return B; }
M3 gamma_func(const double* vec, int m) {  // AuxGammaFunc.cpp:46-113
  const double th = std::sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  M3 out{};
  if (std::fabs(th) < 1e-6) {
    const double f = (m == 3) ? 1.0 / 6.0 : (m == 2 ? 0.5 : 1.0);
    out.m[0] = out.m[4] = out.m[8] = f;
    return out;
  }
  const double n[3] = {vec[0] / th, vec[1] / th, vec[2] / th};
  const M3 nx = skew(n), nx2 = mul(nx, nx);
  const double s = std::sin(th), c = std::cos(th);
  double f0, f1, f2;
  if (m == 1) { f0 = 1; f1 = (1 - c) / th; f2 = (th - s) / th; }
  else if (m == 2) { f0 = 0.5; f1 = (th - s) / (th * th); f2 = (th * th + 2 * c - 2) / (2 * th * th); }
  else if (m == 3) { const double t3 = th * th * th; f0 = 1.0 / 6; f1 = (th * th + 2 * c - 2) / (2 * t3); f2 = (t3 - 6 * th + 6 * s) / (6 * t3); }
  else { f0 = 1; f1 = s; f2 = 1 - c; }
  for (int i = 0; i < 9; ++i) out.m[i] = f1 * nx.m[i] + f2 * nx2.m[i];
  out.m[0] += f0; out.m[4] += f0; out.m[8] += f0;
  return out;
}
M3 psi_func(const double* w, const double* a, double dt, int which) {  // AuxGammaFunc.cpp:115-225
  const double wn = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  M3 zero{};
  if (wn * std::fabs(dt) < (which == 1 ? 1e-8 : 1e-7)) return zero;
  const double nw[3] = {-w[0] * dt, -w[1] * dt, -w[2] * dt};
  const M3 W = skew(w), A = skew(a);
  const M3 M1 = scaled(mul(A, gamma_func(nw, which == 1 ? 2 : 3)), which == 1 ? dt * dt : dt * dt * dt);
  const M3 WA = mul(W, A), WAW = mul(WA, W), WAW2 = mul(WAW, W), W2A = mul(W, WA), W2AW = mul(W2A, W), W2AW2 = mul(W2AW, W);
  const double eta = wn, xi = eta * dt, xi2 = xi * xi, xi3 = xi2 * xi;
  const double s1 = std::sin(xi), c1 = std::cos(xi), s2 = std::sin(2 * xi), c2 = std::cos(2 * xi);
  const double e3 = eta * eta * eta, e4 = e3 * eta, e5 = e4 * eta, e6 = e5 * eta, e7 = e6 * eta;
  double k1, k2, k3, k4, k5, k6;
  if (which == 1) {
    k1 = (s1 - xi * c1) / e3; k2 = (c2 - 4 * c1 + 3) / (4 * e4); k3 = (4 * s1 + s2 - 4 * xi * c1 - 2 * xi) / (4 * e5);
    k4 = (xi2 - 2 * xi * s1 - 2 * c1 + 2) / (2 * e4); k5 = (6 * xi - 8 * s1 + s2) / (4 * e5); k6 = (2 * xi2 - 4 * xi * s1 - c2 + 1) / (4 * e6);
  } else {
    k1 = (xi * s1 + 2 * c1 - 2) / e4; k2 = (6 * xi - 8 * s1 + s2) / (8 * e5); k3 = (2 * xi2 + 8 * xi * s1 + 16 * c1 + c2 - 17) / (8 * e6);
    k4 = (xi3 + 6 * xi - 12 * s1 + 6 * xi * c1) / (6 * e5); k5 = (6 * xi2 + 16 * c1 - c2 - 15) / (8 * e6);
    k6 = (4 * xi3 + 6 * xi - 24 * s1 - 3 * s2 + 24 * xi * c1) / (24 * e7);
  }
  M3 T;
  for (int i = 0; i < 9; ++i) T.m[i] = k1 * WA.m[i] + k2 * WAW.m[i] + k3 * WAW2.m[i] + k4 * W2A.m[i] + k5 * W2AW.m[i] + k6 * W2AW2.m[i];
  return mul(M1, T);
}

struct Clone { M3 R; double p[3]; int idx; };

struct Filter {
  // parameters
  double ng, na, nbg, nba, ncb, ncbrw, g[3];
  M3 Rc; double pc[3];
  int stereo;
  std::vector<double> chi2;
  // state
  Mat P;
  M3 R; double p[3], v[3], bg[3], ba[3];
  M3 Rext; double pext[3];
  double gval[6]; int gidx[6];
  std::vector<Clone> clones;
  int n_accepted = 0;
  std::vector<double> gammas;

  int N() const { return P.r; }

  void retract(M3& Rr, double* p1, double* p2, const double* d) {
    const M3 G0 = gamma_func(d, 0), G1 = gamma_func(d, 1);
    Rr = mul(G0, Rr);
    double t[3], u[3];
    mv(G0, p1, t); mv(G1, d + 3, u);
    for (int i = 0; i < 3; ++i) p1[i] = t[i] + u[i];
    if (p2) { mv(G0, p2, t); mv(G1, d + 6, u); for (int i = 0; i < 3; ++i) p2[i] = t[i] + u[i]; }
  }
  void box_plus(const std::vector<double>& dx) {  // StateManager.cpp:244-251
    retract(R, p, v, &dx[0]);
    for (int i = 0; i < 3; ++i) { bg[i] += dx[9 + i]; ba[i] += dx[12 + i]; }
    retract(Rext, pext, nullptr, &dx[15]);
    for (int gq = 0; gq < 6; ++gq) if (gidx[gq] >= 0) gval[gq] += dx[gidx[gq]];
    for (auto& c : clones) retract(c.R, c.p, nullptr, &dx[c.idx]);
  }

  void imu_step(const double* wraw, const double* araw, double dt) {
    // ---- ImuPropagator::stateAndCovTransition (analytic) ----
    Mat Phi(15, 15), G(15, 12);
    for (int i = 0; i < 15; ++i) Phi(i, i) = 1.0;
    const M3 Rh = R;
    double ph[3], vh[3];
    for (int i = 0; i < 3; ++i) { ph[i] = p[i]; vh[i] = v[i]; }
    const M3 SpR = mul(skew(ph), Rh), SvR = mul(skew(vh), Rh);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
      G(r, c) = Rh.m[3 * r + c]; G(3 + r, c) = SpR.m[3 * r + c]; G(6 + r, c) = SvR.m[3 * r + c]; G(6 + r, 3 + c) = Rh.m[3 * r + c];
    }
    for (int r = 0; r < 3; ++r) { G(9 + r, 6 + r) = 1.0; G(12 + r, 9 + r) = 1.0; }
    const double w[3] = {wraw[0] - bg[0], wraw[1] - bg[1], wraw[2] - bg[2]};
    const double a[3] = {araw[0] - ba[0], araw[1] - ba[1], araw[2] - ba[2]};
    const double wd[3] = {w[0] * dt, w[1] * dt, w[2] * dt};
    const M3 G0 = gamma_func(wd, 0), RG1 = mul(Rh, gamma_func(wd, 1)), RG2 = mul(Rh, gamma_func(wd, 2));
    R = mul(Rh, G0);
    double t[3], vn[3], pn[3];
    mv(RG1, a, t);
    for (int i = 0; i < 3; ++i) vn[i] = vh[i] + g[i] * dt + t[i] * dt;
    mv(RG2, a, t);
    for (int i = 0; i < 3; ++i) pn[i] = ph[i] + vh[i] * dt + 0.5 * g[i] * dt * dt + t[i] * dt * dt;
    for (int i = 0; i < 3; ++i) { p[i] = pn[i]; v[i] = vn[i]; }
    if (gidx[4] >= 0) for (int i = 0; i < 4; ++i) if (gidx[i] >= 0) gval[i] += dt * gval[4];
    const M3 Sg = skew(g);
    const M3 A1v = mul(skew(vn), RG1), A2v = mul(Rh, psi_func(w, a, dt, 1));
    const M3 A1p = mul(skew(pn), RG1), A2p = mul(Rh, psi_func(w, a, dt, 2));
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
      Phi(3 + r, c) = 0.5 * Sg.m[3 * r + c] * dt * dt;
      Phi(6 + r, c) = Sg.m[3 * r + c] * dt;
      Phi(r, 9 + c) = -RG1.m[3 * r + c] * dt;
      Phi(6 + r, 12 + c) = -RG1.m[3 * r + c] * dt;
      Phi(3 + r, 12 + c) = -RG2.m[3 * r + c] * dt * dt;
      Phi(6 + r, 9 + c) = -A1v.m[3 * r + c] * dt + A2v.m[3 * r + c];
      Phi(3 + r, 9 + c) = -A1p.m[3 * r + c] * dt + A2p.m[3 * r + c];
    }
    for (int r = 0; r < 3; ++r) Phi(3 + r, 6 + r) = dt;
    propagate_cov(Phi, G, dt);
  }

  void propagate_cov(const Mat& Phi, const Mat& Gin, double dt) {  // StateManager.cpp:42-119
    const int n = N();
    Mat cov_tmp(n, n);
    Mat P11(15, 15);
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) P11(i, j) = P(i, j);
    const Mat A = matmulT(matmul(Phi, P11), Phi);
    Mat P21(n - 15, 15);
    for (int i = 15; i < n; ++i) for (int j = 0; j < 15; ++j) P21(i - 15, j) = P(i, j);
    Mat cov21 = matmulT(P21, Phi);
    Mat cov22(n - 15, n - 15);
    for (int i = 15; i < n; ++i) for (int j = 15; j < n; ++j) cov22(i - 15, j - 15) = P(i, j);
    if (gidx[4] >= 0) {
      Mat c21t = cov21, c22t = cov22;
      const int lc = gidx[4] - 15;
      for (int q = 0; q < 4; ++q) if (gidx[q] >= 0) {
        const int lr = gidx[q] - 15;
        for (int j = 0; j < 15; ++j) c21t(lr, j) += dt * cov21(lc, j);
        for (int j = 0; j < n - 15; ++j) c22t(lr, j) += dt * cov22(lc, j);
      }
      cov21 = c21t;
      cov22 = c22t;
      for (int q = 0; q < 4; ++q) if (gidx[q] >= 0) {
        const int lr = gidx[q] - 15;
        for (int i = 0; i < n - 15; ++i) c22t(i, lr) += dt * cov22(i, lc);
      }
      cov22 = c22t;
    }
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) cov_tmp(i, j) = A(i, j);
    for (int i = 15; i < n; ++i) for (int j = 0; j < 15; ++j) { cov_tmp(i, j) = cov21(i - 15, j); cov_tmp(j, i) = cov21(i - 15, j); }
    for (int i = 15; i < n; ++i) for (int j = 15; j < n; ++j) cov_tmp(i, j) = cov22(i - 15, j - 15);
    Mat Gt = Gin;
    const double sg[4] = {ng, na, nbg, nba};
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 12; ++j) Gt(i, j) *= sg[j / 3];
    const Mat PG = matmul(Phi, Gt);
    const Mat Q = matmulT(PG, PG);
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) cov_tmp(i, j) += dt * Q(i, j);
    for (int i = 0; i < 5; ++i) {
      if (gidx[i] < 0) continue;
      for (int j = 0; j < 5; ++j) {
        if (gidx[j] < 0) continue;
        double q;
        if (i != 4 && j != 4) q = dt * ncb * ncb + dt * dt * dt * ncbrw * ncbrw;
        else if (i == 4 && j == 4) q = dt * ncbrw * ncbrw;
        else q = dt * dt * ncbrw * ncbrw;
        cov_tmp(gidx[i], gidx[j]) += q;
      }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) P(i, j) = 0.5 * (cov_tmp(i, j) + cov_tmp(j, i));
  }

  void augment() {  // StateManager.cpp:253-296
    const int n = N();
    Clone c;
    c.R = mul(R, Rext);
    double t[3];
    mv(R, pext, t);
    for (int i = 0; i < 3; ++i) c.p[i] = t[i] + p[i];
    c.idx = n;
    clones.push_back(c);
    Mat J(6, 21);
    for (int i = 0; i < 6; ++i) J(i, i) = 1.0;
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) { J(r, 15 + cc) = R.m[3 * r + cc]; J(3 + r, 18 + cc) = R.m[3 * r + cc]; }
    Mat P21n(21, n);
    for (int i = 0; i < 21; ++i) for (int j = 0; j < n; ++j) P21n(i, j) = P(i, j);
    const Mat JP = matmul(J, P21n);
    Mat JP21(6, 21);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 21; ++j) JP21(i, j) = JP(i, j);
    const Mat C = matmulT(JP21, J);
    Mat Pn(n + 6, n + 6);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Pn(i, j) = P(i, j);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < n; ++j) { Pn(n + i, j) = JP(i, j); Pn(j, n + i) = JP(i, j); }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Pn(n + i, n + j) = C(i, j);
    P = Mat(n + 6, n + 6);
    for (int i = 0; i < n + 6; ++i) for (int j = 0; j < n + 6; ++j) P(i, j) = 0.5 * (Pn(i, j) + Pn(j, i));
  }

  void marginalize(int s, int sz) {  // StateManager.cpp:155-192
    const int n = N(), nn = n - sz;
    Mat Pn(nn, nn);
    for (int i = 0; i < nn; ++i) for (int j = 0; j < nn; ++j) Pn(i, j) = P(i + (i >= s ? sz : 0), j + (j >= s ? sz : 0));
    P = Pn;
    for (int q = 0; q < 6; ++q) if (gidx[q] > s) gidx[q] -= sz;
    for (auto& c : clones) if (c.idx > s) c.idx -= sz;
  }
  void marg_clone(int slot) {
    const int s = clones[slot].idx;
    clones.erase(clones.begin() + slot);
    marginalize(s, 6);
  }

  // StateManager::ekfUpdate with var_order given as (idx,size) blocks
  void ekf_update(const std::vector<std::pair<int, int>>& order, const Mat& H, const std::vector<double>& res, const Mat& Rn) {
    const int n = N(), r = H.r;
    std::vector<int> cols;
    for (auto& b : order) for (int k = 0; k < b.second; ++k) cols.push_back(b.first + k);
    Mat PHt(n, r);
    for (int i = 0; i < n; ++i)
      for (size_t c = 0; c < cols.size(); ++c) {
        const double pv = P(i, cols[c]);
        if (pv == 0.0) continue;
        for (int a = 0; a < r; ++a) PHt(i, a) += pv * H(a, (int)c);
      }
    Mat small((int)cols.size(), (int)cols.size());
    for (size_t i = 0; i < cols.size(); ++i) for (size_t j = 0; j < cols.size(); ++j) small((int)i, (int)j) = P(cols[i], cols[j]);
    Mat S = matmulT(matmul(H, small), H);
    for (int i = 0; i < r; ++i) for (int j = 0; j < r; ++j) S(i, j) += Rn(i, j);
    const Mat K = matmul(PHt, lu_inverse(S));
    const Mat KPHt = matmulT(K, PHt);
    Mat T(n, n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) T(i, j) = P(i, j) - KPHt(i, j);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) P(i, j) = 0.5 * (T(i, j) + T(j, i));
    std::vector<double> dx(n, 0.0);
    for (int i = 0; i < n; ++i) { double s = 0.0; for (int a = 0; a < r; ++a) s += K(i, a) * res[a]; dx[i] = s; }
    box_plus(dx);
  }

  // RemoveLost-style all-observation update over the window; `keep cols` compression rule.
  void msckf_update(int F, const double* pfw, const int* anchor, const double* obs, const unsigned char* mask,
                    const int* dof, int obs_slots, double noise, int max_valid) {
    const int rho = stereo ? 4 : 2, ncl = (int)clones.size(), maxc = 6 * ncl;
    std::vector<Mat> blocks;
    std::vector<std::vector<double>> rblocks;
    gammas.assign(F, NAN);
    int valid = 0, rows_total = 0;
    for (int f = 0; f < F; ++f) {
      const double* pf = pfw + 3 * f;
      int nobs = 0;
      for (int s = 0; s < ncl; ++s) nobs += mask[f * obs_slots + s] ? 1 : 0;
      const int M = rho * nobs;
      if (M - 3 < 1) continue;
      Mat Hx(M, maxc), Hf(M, 3);
      std::vector<double> r0(M);
      int row = 0;
      const int anc = anchor[f];
      for (int s = 0; s < ncl; ++s) {
        if (!mask[f * obs_slots + s]) continue;
        const Clone& c = clones[s];
        const double d[3] = {pf[0] - c.p[0], pf[1] - c.p[1], pf[2] - c.p[2]};
        double pcm[3];
        mtv(c.R, d, pcm);
        double Hp[2][3] = {{1 / pcm[2], 0, -pcm[0] / (pcm[2] * pcm[2])}, {0, 1 / pcm[2], -pcm[1] / (pcm[2] * pcm[2])}};
        // H_pf2x blocks: rot (R^T skew(pf)) on the observing clone, minus it on the anchor, -R^T on trans
        M3 Rt;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt.m[3 * i + j] = c.R.m[3 * j + i];
        const M3 RtS = mul(Rt, skew(pf));
        const int nr = rho / 2;
        double Hpr[2][3];
        double prm[3] = {0, 0, 0};
        if (stereo) {
          mv(Rc, pcm, prm);
          for (int i = 0; i < 3; ++i) prm[i] += pc[i];
          const double hp[2][3] = {{1 / prm[2], 0, -prm[0] / (prm[2] * prm[2])}, {0, 1 / prm[2], -prm[1] / (prm[2] * prm[2])}};
          for (int t = 0; t < 2; ++t) for (int j = 0; j < 3; ++j) Hpr[t][j] = hp[t][0] * Rc.m[j] + hp[t][1] * Rc.m[3 + j] + hp[t][2] * Rc.m[6 + j];
        }
        for (int cam = 0; cam < nr; ++cam) {
          const double (*HP)[3] = cam == 0 ? Hp : Hpr;
          for (int t = 0; t < 2; ++t) {
            const int rr = row + 2 * cam + t;
            for (int j = 0; j < 3; ++j) {
              double a_rot = 0, a_tr = 0;
              for (int k = 0; k < 3; ++k) { a_rot += HP[t][k] * RtS.m[3 * k + j]; a_tr += HP[t][k] * Rt.m[3 * k + j]; }
              if (s != anc) { Hx(rr, 6 * s + j) = a_rot; Hx(rr, 6 * anc + j) = -a_rot; }
              Hx(rr, 6 * s + 3 + j) = -a_tr;
              Hf(rr, j) = a_tr;
            }
          }
        }
        const double* z = obs + ((size_t)f * obs_slots + s) * rho;
        r0[row] = z[0] - pcm[0] / pcm[2];
        r0[row + 1] = z[1] - pcm[1] / pcm[2];
        if (stereo) { r0[row + 2] = z[2] - prm[0] / prm[2]; r0[row + 3] = z[3] - prm[1] / prm[2]; }
        row += rho;
      }
      // left null space of Hf: Q^T from Householder QR of Hf applied to [Hx | r]
      Mat B(M, maxc + 1);
      for (int i = 0; i < M; ++i) { for (int j = 0; j < maxc; ++j) B(i, j) = Hx(i, j); B(i, maxc) = r0[i]; }
      householder_qr(Hf, &B);
      const int q = M - 3;
      Mat Hb(q, maxc);
      std::vector<double> rb(q);
      for (int i = 0; i < q; ++i) { for (int j = 0; j < maxc; ++j) Hb(i, j) = B(i + 3, j); rb[i] = B(i + 3, maxc); }
      // chi^2 gate against the marginal covariance of the window
      Mat small(maxc, maxc);
      for (int i = 0; i < maxc; ++i) for (int j = 0; j < maxc; ++j) small(i, j) = P(clones[i / 6].idx + i % 6, clones[j / 6].idx + j % 6);
      Mat S = matmulT(matmul(Hb, small), Hb);
      for (int i = 0; i < q; ++i) S(i, i) += noise * noise;
      const double gam = chi2_stat(S, rb);
      gammas[f] = gam;
      const int d = dof[f];
      if (!(d >= 1 && d <= (int)chi2.size() && gam < chi2[d - 1])) continue;
      blocks.push_back(Hb);
      rblocks.push_back(rb);
      rows_total += q;
      if (++valid >= max_valid && max_valid > 0) break;
    }
    n_accepted = valid;
    if (rows_total == 0) return;
    Mat HL(rows_total, maxc), rL(rows_total, 1);
    int r0i = 0;
    for (size_t k = 0; k < blocks.size(); ++k) {
      for (int i = 0; i < blocks[k].r; ++i) { for (int j = 0; j < maxc; ++j) HL(r0i + i, j) = blocks[k](i, j); rL(r0i + i, 0) = rblocks[k][i]; }
      r0i += blocks[k].r;
    }
    int keep = rows_total;
    if (rows_total > maxc) { householder_qr(HL, &rL); keep = maxc; }
    Mat Ht(keep, maxc), Rn(keep, keep);
    std::vector<double> rt(keep);
    for (int i = 0; i < keep; ++i) { for (int j = 0; j < maxc; ++j) Ht(i, j) = HL(i, j); rt[i] = rL(i, 0); Rn(i, i) = noise * noise; }
    std::vector<std::pair<int, int>> order;
    for (auto& c : clones) order.push_back({c.idx, 6});
    ekf_update(order, Ht, rt, Rn);
  }

  void gnss_update(int S, const double* unit, const double* res_pos, const double* res_vel, const double* sig_psr,
                   const double* sig_dopp, const int* sys, const double* Renu, int adjust_yof, int strong_reject) {
    if (gidx[5] < 0 || gidx[4] < 0) return;
    bool any = false;
    for (int i = 0; i < 4; ++i) any = any || gidx[i] >= 0;
    if (!any || S <= 0) return;
    const double yo = gval[5], cy = std::cos(yo), sy = std::sin(yo);
    M3 Re;
    for (int i = 0; i < 9; ++i) Re.m[i] = Renu[i];
    const M3 Rw = mul(Re, M3{{cy, -sy, 0, sy, cy, 0, 0, 0, 1}}), dRw = mul(Re, M3{{-sy, -cy, 0, cy, -sy, 0, 0, 0, 0}});
    std::vector<std::pair<int, int>> order = {{0, 9}, {gidx[5], 1}};
    int col[6] = {-1, -1, -1, -1, -1, -1}, ncol = 10;
    Mat H(2 * S, 16), Rn(2 * S, 2 * S);
    std::vector<double> res(2 * S);
    int row = 0;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) { col[4] = ncol++; order.push_back({gidx[4], 1}); }
      for (int i = 0; i < S; ++i) {
        const int gq = sys[i];
        if (gq < 0 || gq > 3 || gidx[gq] < 0) continue;
        const double* u = unit + 3 * i;
        double uR[3], udR[3];
        mtv(Rw, u, uR);
        mtv(dRw, u, udR);
        const double* x = pass == 0 ? p : v;
        H(row, 0) = uR[1] * x[2] - uR[2] * x[1];
        H(row, 1) = uR[2] * x[0] - uR[0] * x[2];
        H(row, 2) = uR[0] * x[1] - uR[1] * x[0];
        const int o = pass == 0 ? 3 : 6;
        for (int k = 0; k < 3; ++k) H(row, o + k) = -uR[k];
        if (adjust_yof) H(row, 9) = -(udR[0] * x[0] + udR[1] * x[1] + udR[2] * x[2]);
        if (pass == 0) {
          if (col[gq] < 0) { col[gq] = ncol++; order.push_back({gidx[gq], 1}); }
          H(row, col[gq]) = 1.0;
          res[row] = -res_pos[i];
          Rn(row, row) = sig_psr[i] * sig_psr[i];
        } else {
          H(row, col[4]) = 1.0;
          res[row] = -res_vel[i];
          Rn(row, row) = sig_dopp[i] * sig_dopp[i];
        }
        ++row;
      }
    }
    Mat Ht(row, ncol), Rt(row, row);
    std::vector<double> rt(row);
    for (int i = 0; i < row; ++i) { for (int j = 0; j < ncol; ++j) Ht(i, j) = H(i, j); rt[i] = res[i]; Rt(i, i) = Rn(i, i); }
    if (row == 0) return;
    if (row <= 14 && strong_reject) {
      std::vector<int> cols;
      for (auto& b : order) for (int k = 0; k < b.second; ++k) cols.push_back(b.first + k);
      Mat small(ncol, ncol);
      for (int i = 0; i < ncol; ++i) for (int j = 0; j < ncol; ++j) small(i, j) = P(cols[i], cols[j]);
      Mat Sx = matmulT(matmul(Ht, small), Ht);
      for (int i = 0; i < row; ++i) Sx(i, i) += Rt(i, i);
      const double gam = chi2_stat(Sx, rt);
      if (!(row <= (int)chi2.size() && gam < chi2[row - 1])) return;
    }
    ekf_update(order, Ht, rt, Rt);
  }
};

}  // namespace

extern "C" {

void* orc_create(const double* noise6, const double* gravity3, const double* Rc9, const double* pc3, int stereo,
                 const double* chi2, int chi2_n) {
  Filter* f = new Filter();
  f->ng = noise6[0]; f->na = noise6[1]; f->nbg = noise6[2]; f->nba = noise6[3]; f->ncb = noise6[4]; f->ncbrw = noise6[5];
  for (int i = 0; i < 3; ++i) { f->g[i] = gravity3[i]; f->pc[i] = pc3[i]; }
  for (int i = 0; i < 9; ++i) f->Rc.m[i] = Rc9[i];
  f->stereo = stereo;
  f->chi2.assign(chi2, chi2 + chi2_n);
  for (int i = 0; i < 6; ++i) { f->gidx[i] = -1; f->gval[i] = 0.0; }
  return f;
}
void orc_destroy(void* h) { delete static_cast<Filter*>(h); }

void orc_init(void* h, const double* R9, const double* p, const double* v, const double* bg, const double* ba,
              const double* Rext9, const double* pext, const double* diag21) {
  Filter* f = static_cast<Filter*>(h);
  for (int i = 0; i < 9; ++i) { f->R.m[i] = R9[i]; f->Rext.m[i] = Rext9[i]; }
  for (int i = 0; i < 3; ++i) { f->p[i] = p[i]; f->v[i] = v[i]; f->bg[i] = bg[i]; f->ba[i] = ba[i]; f->pext[i] = pext[i]; }
  f->P = Mat(21, 21);
  for (int i = 0; i < 21; ++i) f->P(i, i) = diag21[i];
  f->clones.clear();
  for (int i = 0; i < 6; ++i) { f->gidx[i] = -1; f->gval[i] = 0.0; }
}
void orc_add_gnss(void* h, int gtype, double value, double cov) {  // StateManager.cpp:194-231
  Filter* f = static_cast<Filter*>(h);
  const int n = f->N();
  Mat Pn(n + 1, n + 1);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Pn(i, j) = f->P(i, j);
  Pn(n, n) = cov;
  f->P = Pn;
  f->gidx[gtype] = n;
  f->gval[gtype] = value;
}
int orc_dim(void* h) { return static_cast<Filter*>(h)->N(); }
int orc_num_clones(void* h) { return (int)static_cast<Filter*>(h)->clones.size(); }
int orc_n_accepted(void* h) { return static_cast<Filter*>(h)->n_accepted; }
void orc_get_cov(void* h, double* dst) {
  Filter* f = static_cast<Filter*>(h);
  std::memcpy(dst, f->P.a.data(), sizeof(double) * f->P.a.size());
}
void orc_get_gammas(void* h, double* dst, int F) {
  Filter* f = static_cast<Filter*>(h);
  for (int i = 0; i < F && i < (int)f->gammas.size(); ++i) dst[i] = f->gammas[i];
}
// packed mean in the layout of include/ingvio_b200.h (39 + 12 per clone)
void orc_get_state(void* h, double* x) {
  Filter* f = static_cast<Filter*>(h);
  for (int i = 0; i < 9; ++i) { x[i] = f->R.m[i]; x[21 + i] = f->Rext.m[i]; }
  for (int i = 0; i < 3; ++i) { x[9 + i] = f->p[i]; x[12 + i] = f->v[i]; x[15 + i] = f->bg[i]; x[18 + i] = f->ba[i]; x[30 + i] = f->pext[i]; }
  for (int i = 0; i < 6; ++i) x[33 + i] = f->gidx[i] >= 0 ? f->gval[i] : 0.0;
  for (size_t s = 0; s < f->clones.size(); ++s) {
    for (int i = 0; i < 9; ++i) x[39 + 12 * s + i] = f->clones[s].R.m[i];
    for (int i = 0; i < 3; ++i) x[39 + 12 * s + 9 + i] = f->clones[s].p[i];
  }
}

// One frame cycle in the order of IngvioFilter::callbackMonoFrame (IngvioFilter.cpp:143-231).
void orc_step(void* h, int K, const double* gyro, const double* accel, const double* dt, int F, const double* pf,
              const int* anchor, const double* obs, const unsigned char* mask, const int* dof, int obs_slots,
              double noise, int max_valid, int n_marg, const int* marg_slots_desc, int S, const double* unit,
              const double* res_pos, const double* res_vel, const double* sig_psr, const double* sig_dopp,
              const int* sys, const double* Renu, int adjust_yof, int strong_reject) {
  Filter* f = static_cast<Filter*>(h);
  for (int k = 0; k < K; ++k) if (dt[k] >= 1e-6) f->imu_step(gyro + 3 * k, accel + 3 * k, dt[k]);
  f->augment();
  if (F > 0) f->msckf_update(F, pf, anchor, obs, mask, dof, obs_slots, noise, max_valid);
  for (int i = 0; i < n_marg; ++i) f->marg_clone(marg_slots_desc[i]);
  if (S > 0) f->gnss_update(S, unit, res_pos, res_vel, sig_psr, sig_dopp, sys, Renu, adjust_yof, strong_reject);
}

}  // extern "C"
