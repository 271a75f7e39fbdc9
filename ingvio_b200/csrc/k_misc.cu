// Rare operations: delayed initialisation of a new scalar variable and linear variable replacement.
//
// Reference: StateManager::addVariableDelayed (StateManager.cpp:547-630) = Givens QR of H_new applied
// to [H_old, res] (:580-592) + chi^2 on the remainder (:604-621) + addVariableDelayedInvertible
// (:462-541) + ekfUpdate on the remaining rows (:626-627);  StateManager::replaceVarLinear (:632-693).
// The Givens sweep is replaced by one Householder reflector (H_new is a single column for every
// call site in scope, GnssUpdate.cpp:389-390,448-449): same orthogonal split up to row signs, same
// posterior.
#include "igv_device.cuh"

using namespace igv;

namespace {

// Householder on H_new (rows x 1) applied to H_old (rows x n, col-major ld=rows) and res. Output in
// the workspace: Hw (rows x n), rw (rows), rho (first entry of Q^T H_new).
__global__ void k_delayed_prep(int rows, int n, const double* Hold, const double* Hnew, const double* res, double* Hw,
                               double* rw, double* rho_out) {
  const int b = blockIdx.x;
  const double* Ho = Hold + (size_t)b * rows * n;
  const double* hn = Hnew + (size_t)b * rows;
  const double* rr = res + (size_t)b * rows;
  double* Hb = Hw + (size_t)b * rows * n;
  double* rb = rw + (size_t)b * rows;
  __shared__ double v[128];
  __shared__ double s_tau, s_beta;
  if (threadIdx.x == 0) {
    double ss = 0.0;
    for (int i = 1; i < rows; ++i) ss += hn[i] * hn[i];
    const double alpha = hn[0];
    double beta = alpha, tau = 0.0, scale = 0.0;
    if (ss > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    v[0] = 1.0;
    for (int i = 1; i < rows; ++i) v[i] = hn[i] * scale;
    s_tau = tau; s_beta = beta;
    rho_out[b] = beta;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= n; c += blockDim.x) {
    const double* src = (c < n) ? Ho + (size_t)c * rows : rr;
    double* dst = (c < n) ? Hb + (size_t)c * rows : rb;
    double w = 0.0;
    for (int i = 0; i < rows; ++i) w += v[i] * src[i];
    w *= s_tau;
    for (int i = 0; i < rows; ++i) dst[i] = src[i] - w * v[i];
  }
}

// Decide + augment (StateManager.cpp:617-623 + :462-541) for a 1-dim new variable appended at N.
__global__ void k_delayed_augment(double* P, int ld, int N, IgvBlocks blk, int rows, const double* Hw, const double* rw,
                                  const double* rho, const double* gamma, double noise2_all, double thr_all, int do_chi2,
                                  double prior_cov, int* accepted, double* X, int xsize, int gslot,
                                  const double* noise2_dev, const int* rows_dev, const double* chi2_095, double chi2_mult) {
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  const double* Hb = Hw + (size_t)b * rows * blk.n;  // row 0 = Hxinit
  __shared__ int cols[6 * IGV_MAX_BLOCKS];
  __shared__ double sPH[512];
  __shared__ int s_acc;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int q = 0; q < blk.n_blocks; ++q) for (int k = 0; k < blk.size[q]; ++k) cols[off++] = blk.idx[q] + k;
    // per-sequence row count (rows beyond it are zero padding): dof = res.rows() of THIS sequence; fewer than two rows:
    // "H_new rows should be larger than H_new cols" -> not added (StateManager.cpp:574-578)
    const int rows_b = rows_dev ? rows_dev[b] : rows;
    const double thr = rows_dev ? ((rows_b >= 1 && rows_b <= 128) ? chi2_mult * chi2_095[rows_b - 1] : 0.0) : thr_all;
    const bool rej = rows_b < 2 || (do_chi2 && rows > 1 && !(gamma[b] <= thr));   // reject if chi2 > mult*quantile
    s_acc = rej ? 0 : 1;
    accepted[b] = s_acc;
  }
  __syncthreads();
  const int n = blk.n;
  const double noise2 = noise2_dev ? noise2_dev[b] : noise2_all;
  const double ir = s_acc ? 1.0 / rho[b] : 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc = fma(Pb[i + (size_t)cols[c] * ld], Hb[(size_t)c * rows], acc);
    sPH[i] = acc;   // (P H^T)[i]
  }
  __syncthreads();
  if (s_acc) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double val = -sPH[i] * ir;
      Pb[i + (size_t)N * ld] = val;
      Pb[N + (size_t)i * ld] = val;
    }
    if (threadIdx.x == 0) {
      double s = noise2;
      for (int c = 0; c < n; ++c) s = fma(Hb[(size_t)c * rows], sPH[cols[c]], s);
      Pb[N + (size_t)N * ld] = s * ir * ir;
    }
  } else {
    for (int i = threadIdx.x; i < N; i += blockDim.x) { Pb[i + (size_t)N * ld] = 0.0; Pb[N + (size_t)i * ld] = 0.0; }
    if (threadIdx.x == 0) Pb[N + (size_t)N * ld] = prior_cov;
  }
  (void)rw; (void)X; (void)xsize; (void)gslot;
}

__global__ void k_replace_var_linear(double* P, int ld, int N, int t0, int ts, IgvBlocks blk, const double* H) {
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  const double* Hb = H + (size_t)b * ts * blk.n;  // ts x n col-major
  extern __shared__ double sm[];
  double* PH = sm;                 // N x ts col-major
  double* HPH = PH + (size_t)N * ts;  // ts x ts
  __shared__ int cols[6 * IGV_MAX_BLOCKS];
  if (threadIdx.x == 0) {
    int off = 0;
    for (int q = 0; q < blk.n_blocks; ++q) for (int k = 0; k < blk.size[q]; ++k) cols[off++] = blk.idx[q] + k;
  }
  __syncthreads();
  const int n = blk.n;
  for (int t = threadIdx.x; t < N * ts; t += blockDim.x) {
    const int i = t % N, a_ = t / N;
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc = fma(Pb[i + (size_t)cols[c] * ld], Hb[a_ + (size_t)c * ts], acc);
    PH[t] = acc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ts * ts; t += blockDim.x) {
    const int a_ = t % ts, c_ = t / ts;
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc = fma(Hb[a_ + (size_t)c * ts], PH[cols[c] + (size_t)c_ * N], acc);
    HPH[t] = acc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < N * ts; t += blockDim.x) {
    const int i = t % N, a_ = t / N;
    Pb[i + (size_t)(t0 + a_) * ld] = PH[t];
    Pb[(t0 + a_) + (size_t)i * ld] = PH[t];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ts * ts; t += blockDim.x) {
    const int a_ = t % ts, c_ = t / ts;
    Pb[(t0 + a_) + (size_t)(t0 + c_) * ld] = HPH[t];
  }
}

}  // namespace

void igv_launch_delayed_init(igv_batch* h, const IgvBlocks& blk, int rows, const double* Hold, const double* Hnew,
                             const double* res, double noise_iso, const double* noise2_dev, const int* rows_dev,
                             double chi2_mult, int do_chi2, double prior_cov, int* accepted_dev) {
  igv_commit_copies(h);
  // workspace: Hw (rows x n) | rw (rows) | rho   -- carved from Dws (B x (128*18+2) doubles, n <= 16)
  double* Hw = h->Dws;
  double* rw = Hw + (size_t)h->B * rows * blk.n;
  double* rho = rw + (size_t)h->B * rows;
  double* gam = h->gam_ws;
  k_delayed_prep<<<h->B, 64, 0, h->stream>>>(rows, blk.n, Hold, Hnew, res, Hw, rw, rho);
  h->launches++;
  if (rows > 1) {  // chi^2 of the remaining rows against the prior (StateManager.cpp:604-611)
    IgvEkfLaunch e{};
    e.blk = blk; e.rows = rows - 1;
    e.H = Hw + 1; e.strideH = (long)rows * blk.n; e.h_ld = rows; e.h_rowmajor = 0;
    e.res = rw + 1; e.strideRes = rows; e.res_inc = 1;
    e.R = noise2_dev; e.strideR = noise2_dev ? 1 : 0; e.r_kind = IGV_R_ISO; e.r_iso_value = noise_iso * noise_iso;
    e.gamma_only = 1; e.gamma_out = gam; e.apply_boxplus = 0;
    igv_launch_ekf(h, e);
  }
  // the reference hard-codes the 0.95 quantile at dof = res.rows() here, whatever UpdateBase::_thres the updaters
  // were built with (StateManager.cpp:613-617)
  const double thr = (rows >= 1) ? chi2_mult * igv_chi2_quantile(0.95, rows) : INFINITY;
  k_delayed_augment<<<h->B, 128, 0, h->stream>>>(h->Pc(), h->ld, h->N, blk, rows, Hw, rw, rho, gam,
                                                 noise_iso * noise_iso, thr, do_chi2, prior_cov, accepted_dev,
                                                 h->Xc(), h->xsize, -1, noise2_dev, rows_dev, h->chi2_095, chi2_mult);
  h->launches++;
}

void igv_launch_replace_var_linear(igv_batch* h, int tidx, int tsize, const IgvBlocks& blk, const double* H) {
  igv_commit_copies(h);
  const size_t smem = sizeof(double) * ((size_t)h->N * tsize + tsize * tsize);
  k_replace_var_linear<<<h->B, 128, smem, h->stream>>>(h->Pc(), h->ld, h->N, tidx, tsize, blk, H);
  h->launches++;
}
