// Per-feature inverse-depth Levenberg-Marquardt triangulation with Huber weights, clone poses read from the device
// mean mirror.  Default kernel k_triangulate_grp: a GROUP of 16 (or 32) lanes per track, one lane per view -- every lane
// keeps its view's relative pose in registers for the whole solve, residuals / Jacobians / costs are evaluated once per
// lane and the 3 x 3 normal equations, the cost and the argmax of findLongestTrans are group reductions by shuffles.
// k_triangulate (one THREAD per track, relative poses recomputed on the fly) is the first version, kept as the
// cross-check (IGV_TRI_CFG=1) and for the CPU execution model of tests/emul.
//
// Reference: Triangulator::triangulateMonoObs / triangulateStereoObs (Triangulator.cpp:173-359) with its
// helpers findLongestTrans (:30-66), calcRelaSwPose (:68-87), initDepth (:89-105), calcUnitCost (:107-124),
// calcResJacobian (:138-171), and the anchor-depth check of
// FeatureInfoManager::triangulateFeatureInfo{Mono,Stereo} (MapServerManager.cpp:275-341).
// The control flow (loop counters, damping schedule, acceptance tests, depth gates) is restated 1:1;
// the 3x3 `ldlt().solve` becomes an explicit LDL^T. Relative poses are recomputed on the fly instead of
// being materialised per track (no per-thread arrays).
#include "igv_device.cuh"

using namespace igv;

namespace {

struct TriArgs {
  const double* X; int xsize; int n_clones;
  int F, obs_slots, rho;
  const double* obs; const unsigned char* mask; const int* anchor;
  igv_tri_params prm;
  double Rc[9], pc[3];
  double* pf_out; unsigned char* ok_out;
  int B;
};

struct View { double R[9]; double p[3]; double m[2]; };

__device__ __forceinline__ int nth_set(unsigned long long bits, int n) {
  for (int i = 0; i < n; ++i) bits &= bits - 1;
  return __ffsll((long long)bits) - 1;
}

// k-th mono view of the track (stereo: left camera = even k, right camera = odd k with pose T_left * T_cl2cr^-1)
__device__ __forceinline__ void load_view(const TriArgs& a, const double* Xb, const double* ob, unsigned long long bits, int k,
                                          View& v) {
  const int cam = (a.rho == 4) ? (k & 1) : 0;
  const int slot = nth_set(bits, (a.rho == 4) ? (k >> 1) : k);
  const double* c = Xb + IGV_X_CORE + 12 * slot;
  if (cam == 0) {
    for (int i = 0; i < 9; ++i) v.R[i] = c[i];
    for (int i = 0; i < 3; ++i) v.p[i] = c[9 + i];
    v.m[0] = ob[slot * a.rho]; v.m[1] = ob[slot * a.rho + 1];
  } else {
    // R_r = R Rc^T ; p_r = p - R_r pc   (Triangulator.cpp:353)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) v.R[3 * i + j] = c[3 * i] * a.Rc[3 * j] + c[3 * i + 1] * a.Rc[3 * j + 1] + c[3 * i + 2] * a.Rc[3 * j + 2];
    double t[3];
    mat3_vec(v.R, a.pc, t);
    for (int i = 0; i < 3; ++i) v.p[i] = c[9 + i] - t[i];
    v.m[0] = ob[slot * a.rho + 2]; v.m[1] = ob[slot * a.rho + 3];
  }
}

// T_rel = T_v^-1 T_last  (identity when v is the last view, Triangulator.cpp:80-84)
__device__ __forceinline__ void rel_pose(const View& v, const View& last, bool is_last, double* R, double* t) {
  if (is_last) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    t[0] = t[1] = t[2] = 0.0;
    return;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = v.R[i] * last.R[j] + v.R[3 + i] * last.R[3 + j] + v.R[6 + i] * last.R[6 + j];
  const double d[3] = {last.p[0] - v.p[0], last.p[1] - v.p[1], last.p[2] - v.p[2]};
  mat3T_vec(v.R, d, t);
}

__device__ double total_cost(const TriArgs& a, const double* Xb, const double* ob, unsigned long long bits, int nv,
                             const View& last, const double* sol) {
  const double z = 1.0 / sol[2];
  const double pf0[3] = {sol[0] * z, sol[1] * z, z};
  double tot = 0.0;
  for (int k = 0; k < nv; ++k) {
    View v;
    if (k == nv - 1) v = last; else load_view(a, Xb, ob, bits, k, v);
    double R[9], t[3], pf[3];
    rel_pose(v, last, k == nv - 1, R, t);
    mat3_vec(R, pf0, pf);
    for (int i = 0; i < 3; ++i) pf[i] += t[i];
    const double d0 = v.m[0] - pf[0] / pf[2], d1 = v.m[1] - pf[1] / pf[2];
    tot += d0 * d0 + d1 * d1;
  }
  return tot;
}

__global__ void __launch_bounds__(128) k_triangulate(TriArgs a) {
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)a.B * a.F) return;
  const int b = (int)(gid / a.F);
  const double* Xb = a.X + (size_t)b * a.xsize;
  const double* ob = a.obs + gid * a.obs_slots * a.rho;
  const unsigned char* mk = a.mask + gid * a.obs_slots;
  double* pf_out = a.pf_out + gid * 3;
  pf_out[0] = pf_out[1] = pf_out[2] = 0.0;
  a.ok_out[gid] = 0;
  unsigned long long bits = 0ull;
  for (int s = 0; s < a.n_clones && s < 64; ++s) if (mk[s]) bits |= (1ull << s);
  const int nobs = __popcll(bits);
  const int nv = (a.rho == 4) ? 2 * nobs : nobs;
  if (nv <= 4) return;  // Triangulator.cpp:183
  const igv_tri_params& P = a.prm;
  View last;
  load_view(a, Xb, ob, bits, nv - 1, last);
  // findLongestTrans (:30-66)
  int max_k = nv - 1;
  double max_len = -INFINITY;
  {
    double u[3] = {last.m[0], last.m[1], 1.0};
    const double un = sqrt(u[0] * u[0] + u[1] * u[1] + 1.0);
    for (int i = 0; i < 3; ++i) u[i] /= un;
    double uw[3];
    mat3_vec(last.R, u, uw);
    for (int k = 0; k < nv - 1; ++k) {
      View v;
      load_view(a, Xb, ob, bits, k, v);
      const double d[3] = {v.p[0] - last.p[0], v.p[1] - last.p[1], v.p[2] - last.p[2]};
      const double dot = uw[0] * d[0] + uw[1] * d[1] + uw[2] * d[2];
      const double t[3] = {d[0] - uw[0] * dot, d[1] - uw[1] * dot, d[2] - uw[2] * dot};
      const double len = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
      if (len > max_len) { max_len = len; max_k = k; }
    }
  }
  if (max_len < P.trans_thres) return;  // :192
  // initial solution (:201-203)
  double sol[3];
  {
    View v;
    load_view(a, Xb, ob, bits, max_k, v);
    double R[9], t[3];
    rel_pose(v, last, false, R, t);
    const double m1[3] = {last.m[0], last.m[1], 1.0};
    double tm[3];
    mat3_vec(R, m1, tm);
    const double A0 = tm[0] - v.m[0] * tm[2], A1 = tm[1] - v.m[1] * tm[2];
    const double b0 = v.m[0] * t[2] - t[0], b1 = v.m[1] * t[2] - t[1];
    const double depth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
    sol[0] = last.m[0]; sol[1] = last.m[1]; sol[2] = 1.0 / depth;
  }
  double total = total_cost(a, Xb, ob, bits, nv, last, sol);
  double lambda = P.init_damping;
  int inner = 0, outer = 0;
  bool reduced = false;
  double delta_norm = INFINITY;
  do {
    double A[6] = {0, 0, 0, 0, 0, 0};  // symmetric: 00 01 02 11 12 22
    double bb[3] = {0, 0, 0};
    for (int k = 0; k < nv; ++k) {
      View v;
      if (k == nv - 1) v = last; else load_view(a, Xb, ob, bits, k, v);
      double R[9], t[3];
      rel_pose(v, last, k == nv - 1, R, t);
      // calcResJacobian (:138-171)
      const double tp[3] = {R[0] * sol[0] + R[1] * sol[1] + R[2] + t[0] * sol[2], R[3] * sol[0] + R[4] * sol[1] + R[5] + t[1] * sol[2],
                            R[6] * sol[0] + R[7] * sol[1] + R[8] + t[2] * sol[2]};
      const double iz = 1.0 / tp[2];
      const double res[2] = {tp[0] * iz - v.m[0], tp[1] * iz - v.m[1]};
      const double w02 = -tp[0] / (tp[2] * tp[2]), w12 = -tp[1] / (tp[2] * tp[2]);
      // U = [R(:,0) R(:,1) t];  J = W U
      const double U[9] = {R[0], R[1], t[0], R[3], R[4], t[1], R[6], R[7], t[2]};
      double J[6];
      for (int j = 0; j < 3; ++j) { J[j] = iz * U[j] + w02 * U[6 + j]; J[3 + j] = iz * U[3 + j] + w12 * U[6 + j]; }
      const double e = sqrt(res[0] * res[0] + res[1] * res[1]);
      double w2 = 1.0;
      if (!(e <= P.huber_epsilon)) { const double w = sqrt(2.0 * P.huber_epsilon / e); w2 = w * w; }
      A[0] += w2 * (J[0] * J[0] + J[3] * J[3]); A[1] += w2 * (J[0] * J[1] + J[3] * J[4]); A[2] += w2 * (J[0] * J[2] + J[3] * J[5]);
      A[3] += w2 * (J[1] * J[1] + J[4] * J[4]); A[4] += w2 * (J[1] * J[2] + J[4] * J[5]); A[5] += w2 * (J[2] * J[2] + J[5] * J[5]);
      for (int j = 0; j < 3; ++j) bb[j] -= w2 * (J[j] * res[0] + J[3 + j] * res[1]);
    }
    do {
      // (A + lambda I) delta = b by LDL^T
      const double a00 = A[0] + lambda, a11 = A[3] + lambda, a22 = A[5] + lambda;
      const double l10 = A[1] / a00, l20 = A[2] / a00;
      const double d1 = a11 - l10 * A[1];
      const double l21 = (A[4] - l20 * A[1]) / d1;
      const double d2 = a22 - l20 * A[2] - l21 * l21 * d1;
      const double y0 = bb[0], y1 = bb[1] - l10 * y0, y2 = bb[2] - l20 * y0 - l21 * y1;
      double delta[3];
      delta[2] = y2 / d2;
      delta[1] = y1 / d1 - l21 * delta[2];
      delta[0] = y0 / a00 - l10 * delta[1] - l20 * delta[2];
      const double nsol[3] = {sol[0] + delta[0], sol[1] + delta[1], sol[2] + delta[2]};
      delta_norm = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
      const double ntotal = total_cost(a, Xb, ob, bits, nv, last, nsol);
      if (ntotal < total) {
        total = ntotal;
        sol[0] = nsol[0]; sol[1] = nsol[1]; sol[2] = nsol[2];
        reduced = true;
        lambda = (lambda / 10.0 > 1e-10) ? lambda / 10.0 : 1e-10;
      } else {
        reduced = false;
        lambda = (lambda * 10 < 1e12) ? lambda * 10 : 1e12;
      }
    } while (inner++ < P.inner_loop_max_iter && !reduced);
    inner = 0;
  } while (outer++ < P.outer_loop_max_iter && delta_norm > P.conv_precision);
  const double z = 1.0 / sol[2];
  const double pl[3] = {sol[0] * z, sol[1] * z, z};
  if ((outer >= P.outer_loop_max_iter && inner >= P.inner_loop_max_iter) || delta_norm > P.conv_precision) return;  // :281
  for (int k = 0; k < nv; ++k) {  // :284-289
    View v;
    if (k == nv - 1) v = last; else load_view(a, Xb, ob, bits, k, v);
    double R[9], t[3], q[3];
    rel_pose(v, last, k == nv - 1, R, t);
    mat3_vec(R, pl, q);
    if (q[2] + t[2] <= P.min_depth) return;
  }
  if (pl[2] < P.min_depth || pl[2] > P.max_depth) return;  // :306
  double pf[3];
  mat3_vec(last.R, pl, pf);
  for (int i = 0; i < 3; ++i) pf[i] += last.p[i];
  if (isnan(pf[0]) || isnan(pf[1]) || isnan(pf[2])) return;  // :311
  bool ok = true;
  if (a.anchor) {  // MapServerManager.cpp:289-291: the landmark must be in front of its anchor camera
    const int an = a.anchor[gid];
    if (an >= 0 && an < a.n_clones) {
      const double* c = Xb + IGV_X_CORE + 12 * an;
      const double d[3] = {pf[0] - c[9], pf[1] - c[10], pf[2] - c[11]};
      const double bz = c[2] * d[0] + c[5] * d[1] + c[8] * d[2];
      if (bz <= 0.0) ok = false;
    }
  }
  pf_out[0] = pf[0]; pf_out[1] = pf[1]; pf_out[2] = pf[2];
  a.ok_out[gid] = ok ? 1 : 0;
}

#ifndef IGV_EMULATE
template <int GS>
__device__ __forceinline__ double grp_sum(double v, unsigned gmask) {
#pragma unroll
  for (int off = GS / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(gmask, v, off, GS);
  return v;
}

// One group of GS lanes per track; lane gl owns views gl, gl + GS, ... (the first one cached in registers).
template <int GS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_triangulate_grp(TriArgs a) {
  constexpr int GPB = 128 / GS;                      // groups per block
  const int gl = threadIdx.x % GS;
  const long gid = (long)blockIdx.x * GPB + threadIdx.x / GS;
  if (gid >= (long)a.B * a.F) return;                // group-uniform
  const unsigned gmask = (GS == 32) ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) / 16)));
  const int b = (int)(gid / a.F);
  const double* Xb = a.X + (size_t)b * a.xsize;
  const double* ob = a.obs + gid * a.obs_slots * a.rho;
  const unsigned char* mk = a.mask + gid * a.obs_slots;
  double* pf_out = a.pf_out + gid * 3;
  if (gl == 0) { pf_out[0] = pf_out[1] = pf_out[2] = 0.0; a.ok_out[gid] = 0; }
  unsigned long long bits = 0ull;
  for (int s = 0; s < a.n_clones && s < 64; ++s) if (mk[s]) bits |= (1ull << s);
  const int nobs = __popcll(bits);
  const int nv = (a.rho == 4) ? 2 * nobs : nobs;
  if (nv <= 4) return;  // Triangulator.cpp:183
  const igv_tri_params& P = a.prm;
  View last;
  load_view(a, Xb, ob, bits, nv - 1, last);
  // this lane's first view and its pose relative to the last view, kept for the whole solve
  const bool own = gl < nv;
  View v0;
  double R0[9], t0[3];
  if (own) {
    if (gl == nv - 1) v0 = last; else load_view(a, Xb, ob, bits, gl, v0);
    rel_pose(v0, last, gl == nv - 1, R0, t0);
  }
  auto view_at = [&](int k, double* R, double* t, double* m) {   // k = gl: the cached one; beyond GS views: recomputed
    if (k == gl) {
      for (int i = 0; i < 9; ++i) R[i] = R0[i];
      for (int i = 0; i < 3; ++i) t[i] = t0[i];
      m[0] = v0.m[0]; m[1] = v0.m[1];
    } else {
      View v;
      if (k == nv - 1) v = last; else load_view(a, Xb, ob, bits, k, v);
      rel_pose(v, last, k == nv - 1, R, t);
      m[0] = v.m[0]; m[1] = v.m[1];
    }
  };
  // findLongestTrans (:30-66): first view with the largest translation orthogonal to the last view's bearing
  int max_k = nv - 1;
  double max_len = -INFINITY;
  {
    double u[3] = {last.m[0], last.m[1], 1.0};
    const double un = sqrt(u[0] * u[0] + u[1] * u[1] + 1.0);
    for (int i = 0; i < 3; ++i) u[i] /= un;
    double uw[3];
    mat3_vec(last.R, u, uw);
    for (int k = gl; k < nv - 1; k += GS) {
      View v;
      if (k == gl) v = v0; else load_view(a, Xb, ob, bits, k, v);
      const double d[3] = {v.p[0] - last.p[0], v.p[1] - last.p[1], v.p[2] - last.p[2]};
      const double dot = uw[0] * d[0] + uw[1] * d[1] + uw[2] * d[2];
      const double t[3] = {d[0] - uw[0] * dot, d[1] - uw[1] * dot, d[2] - uw[2] * dot};
      const double len = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
      if (len > max_len) { max_len = len; max_k = k; }
    }
#pragma unroll
    for (int off = GS / 2; off >= 1; off >>= 1) {   // (len, k): larger len wins, ties go to the smaller k like the serial loop
      const double ol = __shfl_xor_sync(gmask, max_len, off, GS);
      const int ok = __shfl_xor_sync(gmask, max_k, off, GS);
      if (ol > max_len || (ol == max_len && ok < max_k)) { max_len = ol; max_k = ok; }
    }
  }
  if (max_len < P.trans_thres) return;  // :192
  double sol[3];
  {   // initial solution (:201-203), evaluated by every lane (uniform)
    View v;
    load_view(a, Xb, ob, bits, max_k, v);
    double R[9], t[3];
    rel_pose(v, last, false, R, t);
    const double m1[3] = {last.m[0], last.m[1], 1.0};
    double tm[3];
    mat3_vec(R, m1, tm);
    const double A0 = tm[0] - v.m[0] * tm[2], A1 = tm[1] - v.m[1] * tm[2];
    const double b0 = v.m[0] * t[2] - t[0], b1 = v.m[1] * t[2] - t[1];
    const double depth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
    sol[0] = last.m[0]; sol[1] = last.m[1]; sol[2] = 1.0 / depth;
  }
  auto cost = [&](const double* s_) {
    const double z = 1.0 / s_[2];
    const double pf0[3] = {s_[0] * z, s_[1] * z, z};
    double tot = 0.0;
    for (int k = gl; k < nv; k += GS) {
      double R[9], t[3], m[2], pf[3];
      view_at(k, R, t, m);
      mat3_vec(R, pf0, pf);
      for (int i = 0; i < 3; ++i) pf[i] += t[i];
      const double d0 = m[0] - pf[0] / pf[2], d1 = m[1] - pf[1] / pf[2];
      tot += d0 * d0 + d1 * d1;
    }
    return grp_sum<GS>(tot, gmask);
  };
  double total = cost(sol);
  double lambda = P.init_damping;
  int inner = 0, outer = 0;
  bool reduced = false;
  double delta_norm = INFINITY;
  do {
    double A[6] = {0, 0, 0, 0, 0, 0};  // symmetric: 00 01 02 11 12 22
    double bb[3] = {0, 0, 0};
    for (int k = gl; k < nv; k += GS) {
      double R[9], t[3], m[2];
      view_at(k, R, t, m);
      // calcResJacobian (:138-171)
      const double tp[3] = {R[0] * sol[0] + R[1] * sol[1] + R[2] + t[0] * sol[2], R[3] * sol[0] + R[4] * sol[1] + R[5] + t[1] * sol[2],
                            R[6] * sol[0] + R[7] * sol[1] + R[8] + t[2] * sol[2]};
      const double iz = 1.0 / tp[2];
      const double res[2] = {tp[0] * iz - m[0], tp[1] * iz - m[1]};
      const double w02 = -tp[0] / (tp[2] * tp[2]), w12 = -tp[1] / (tp[2] * tp[2]);
      const double U[9] = {R[0], R[1], t[0], R[3], R[4], t[1], R[6], R[7], t[2]};
      double J[6];
      for (int j = 0; j < 3; ++j) { J[j] = iz * U[j] + w02 * U[6 + j]; J[3 + j] = iz * U[3 + j] + w12 * U[6 + j]; }
      const double e = sqrt(res[0] * res[0] + res[1] * res[1]);
      double w2 = 1.0;
      if (!(e <= P.huber_epsilon)) { const double w = sqrt(2.0 * P.huber_epsilon / e); w2 = w * w; }
      A[0] += w2 * (J[0] * J[0] + J[3] * J[3]); A[1] += w2 * (J[0] * J[1] + J[3] * J[4]); A[2] += w2 * (J[0] * J[2] + J[3] * J[5]);
      A[3] += w2 * (J[1] * J[1] + J[4] * J[4]); A[4] += w2 * (J[1] * J[2] + J[4] * J[5]); A[5] += w2 * (J[2] * J[2] + J[5] * J[5]);
      for (int j = 0; j < 3; ++j) bb[j] -= w2 * (J[j] * res[0] + J[3 + j] * res[1]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) A[j] = grp_sum<GS>(A[j], gmask);
#pragma unroll
    for (int j = 0; j < 3; ++j) bb[j] = grp_sum<GS>(bb[j], gmask);
    do {
      // (A + lambda I) delta = b by LDL^T
      const double a00 = A[0] + lambda, a11 = A[3] + lambda, a22 = A[5] + lambda;
      const double l10 = A[1] / a00, l20 = A[2] / a00;
      const double d1 = a11 - l10 * A[1];
      const double l21 = (A[4] - l20 * A[1]) / d1;
      const double d2 = a22 - l20 * A[2] - l21 * l21 * d1;
      const double y0 = bb[0], y1 = bb[1] - l10 * y0, y2 = bb[2] - l20 * y0 - l21 * y1;
      double delta[3];
      delta[2] = y2 / d2;
      delta[1] = y1 / d1 - l21 * delta[2];
      delta[0] = y0 / a00 - l10 * delta[1] - l20 * delta[2];
      const double nsol[3] = {sol[0] + delta[0], sol[1] + delta[1], sol[2] + delta[2]};
      delta_norm = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
      const double ntotal = cost(nsol);
      if (ntotal < total) {
        total = ntotal;
        sol[0] = nsol[0]; sol[1] = nsol[1]; sol[2] = nsol[2];
        reduced = true;
        lambda = (lambda / 10.0 > 1e-10) ? lambda / 10.0 : 1e-10;
      } else {
        reduced = false;
        lambda = (lambda * 10 < 1e12) ? lambda * 10 : 1e12;
      }
    } while (inner++ < P.inner_loop_max_iter && !reduced);
    inner = 0;
  } while (outer++ < P.outer_loop_max_iter && delta_norm > P.conv_precision);
  const double z = 1.0 / sol[2];
  const double pl[3] = {sol[0] * z, sol[1] * z, z};
  if ((outer >= P.outer_loop_max_iter && inner >= P.inner_loop_max_iter) || delta_norm > P.conv_precision) return;  // :281
  bool behind = false;
  for (int k = gl; k < nv; k += GS) {  // :284-289
    double R[9], t[3], m[2], q[3];
    view_at(k, R, t, m);
    mat3_vec(R, pl, q);
    if (q[2] + t[2] <= P.min_depth) behind = true;
  }
  if (__ballot_sync(gmask, behind) & gmask) return;
  if (pl[2] < P.min_depth || pl[2] > P.max_depth) return;  // :306
  double pf[3];
  mat3_vec(last.R, pl, pf);
  for (int i = 0; i < 3; ++i) pf[i] += last.p[i];
  if (isnan(pf[0]) || isnan(pf[1]) || isnan(pf[2])) return;  // :311
  bool ok = true;
  if (a.anchor) {  // MapServerManager.cpp:289-291: the landmark must be in front of its anchor camera
    const int an = a.anchor[gid];
    if (an >= 0 && an < a.n_clones) {
      const double* c = Xb + IGV_X_CORE + 12 * an;
      const double d[3] = {pf[0] - c[9], pf[1] - c[10], pf[2] - c[11]};
      const double bz = c[2] * d[0] + c[5] * d[1] + c[8] * d[2];
      if (bz <= 0.0) ok = false;
    }
  }
  if (gl == 0) {
    pf_out[0] = pf[0]; pf_out[1] = pf[1]; pf_out[2] = pf[2];
    a.ok_out[gid] = ok ? 1 : 0;
  }
}
#endif  // !IGV_EMULATE

}  // namespace

#if !defined(IGV_EMULATE) || defined(IGV_EMULATE_LAUNCHERS)   // tests/emul: kernels only, or (full model) launchers too
void igv_launch_triangulate(igv_batch* h, int F, int obs_slots, const double* obs, const unsigned char* mask,
                            const int* anchor, const igv_tri_params& prm, double* pf_out, unsigned char* ok_out) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  TriArgs a;
  a.X = h->Xc(); a.xsize = h->xsize; a.n_clones = h->layout().n_clones;
  a.F = F; a.obs_slots = obs_slots; a.rho = h->rho;
  a.obs = obs; a.mask = mask; a.anchor = anchor; a.prm = prm;
  for (int i = 0; i < 9; ++i) a.Rc[i] = h->params.Rc[i];
  for (int i = 0; i < 3; ++i) a.pc[i] = h->params.pc[i];
  a.pf_out = pf_out; a.ok_out = ok_out; a.B = h->B;
  const long n = (long)h->B * F;
#ifndef IGV_EMULATE
  if (h->knobs.tri_cfg != 1) {
    // a group of lanes per track: 16 when every track fits (views <= 16: mono windows up to 16 clones), else 32
    const int max_views = (h->rho == 4 ? 2 : 1) * a.n_clones;
    // MINB resident blocks per SM: the natural register count (168) leaves 3; IGV_TRI_MINB caps the registers (A/B runs)
    const unsigned g16 = (unsigned)((n + 7) / 8), g32 = (unsigned)((n + 3) / 4);
    const int mb = h->knobs.tri_minb;
    if (max_views <= 16) {
      if (mb == 3) k_triangulate_grp<16, 3><<<g16, 128, 0, h->stream>>>(a);
      else if (mb == 5) k_triangulate_grp<16, 5><<<g16, 128, 0, h->stream>>>(a);
      else if (mb == 6) k_triangulate_grp<16, 6><<<g16, 128, 0, h->stream>>>(a);
      else k_triangulate_grp<16, 4><<<g16, 128, 0, h->stream>>>(a);
    } else {
      if (mb == 3) k_triangulate_grp<32, 3><<<g32, 128, 0, h->stream>>>(a);
      else if (mb == 5) k_triangulate_grp<32, 5><<<g32, 128, 0, h->stream>>>(a);
      else if (mb == 6) k_triangulate_grp<32, 6><<<g32, 128, 0, h->stream>>>(a);
      else k_triangulate_grp<32, 4><<<g32, 128, 0, h->stream>>>(a);
    }
    h->launches++;
    return;
  }
#endif
  k_triangulate<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a);
  h->launches++;
}
#endif  // IGV_EMULATE
