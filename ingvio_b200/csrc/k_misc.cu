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

// k Householder reflectors (k = size of the new variable, 1 or 3) on H_new (rows x k, col-major, ld = rows), applied to
// H_old (rows x n, col-major, ld = rows) and res. Outputs in the workspace: Hw (rows x n), rw (rows) and Rf, the k x k
// upper triangle of Q^T H_new (row-major, stride 3).
__global__ void k_delayed_prep(int rows, int n, int k, const double* Hold, const double* Hnew, const double* res, double* Hw,
                               double* rw, double* Rf_out) {
  const int b = blockIdx.x;
  const double* Ho = Hold + (size_t)b * rows * n;
  const double* hn = Hnew + (size_t)b * rows * k;
  const double* rr = res + (size_t)b * rows;
  double* Hb = Hw + (size_t)b * rows * n;
  double* rb = rw + (size_t)b * rows;
  __shared__ double hf[3 * 128];   // working copy of H_new
  __shared__ double v[128];
  __shared__ double s_tau;
  for (int t = threadIdx.x; t < rows * k; t += blockDim.x) hf[t] = hn[t];
  for (int t = threadIdx.x; t < rows * n; t += blockDim.x) Hb[t] = Ho[t];
  for (int t = threadIdx.x; t < rows; t += blockDim.x) rb[t] = rr[t];
  if (threadIdx.x < 9) Rf_out[(size_t)b * 9 + threadIdx.x] = 0.0;
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (threadIdx.x == 0) {
      double* col = hf + j * rows;
      double ss = 0.0;
      for (int i = j + 1; i < rows; ++i) ss += col[i] * col[i];
      const double alpha = col[j];
      double beta = alpha, tau = 0.0, scale = 0.0;
      if (ss > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + ss), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      for (int i = 0; i < j; ++i) v[i] = 0.0;
      v[j] = 1.0;
      for (int i = j + 1; i < rows; ++i) v[i] = col[i] * scale;
      s_tau = tau;
      Rf_out[(size_t)b * 9 + 3 * j + j] = beta;
      for (int c = j + 1; c < k; ++c) {          // the remaining columns of H_new
        double* cc = hf + c * rows;
        double w = 0.0;
        for (int i = j; i < rows; ++i) w += v[i] * cc[i];
        w *= tau;
        for (int i = j; i < rows; ++i) cc[i] -= w * v[i];
        Rf_out[(size_t)b * 9 + 3 * j + c] = cc[j];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= n; c += blockDim.x) {
      double* x = (c < n) ? Hb + (size_t)c * rows : rb;
      double w = 0.0;
      for (int i = j; i < rows; ++i) w += v[i] * x[i];
      w *= s_tau;
      for (int i = j; i < rows; ++i) x[i] -= w * v[i];
    }
    __syncthreads();
  }
}

// Decide + augment (StateManager.cpp:617-623 + :462-541) for a k-dim new variable appended at N:
//   P[:, new] = -(P H_x^T) H_f^-T,  P[new, new] = H_f^-1 (H_x P_s H_x^T + sigma^2 I) H_f^-T   with H_f = Rf (upper triangular).
__global__ void k_delayed_augment(double* P, int ld, int N, IgvBlocks blk, int rows, int k, const double* Hw, const double* Rf,
                                  const double* gamma, double noise2_all, double thr_all, int do_chi2,
                                  double prior_cov, int* accepted, const double* noise2_dev, const int* rows_dev,
                                  const double* chi2_095, double chi2_mult) {
  extern __shared__ double sPH[];                       // N x k, column a at sPH + a * N
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  const double* Hb = Hw + (size_t)b * rows * blk.n;    // rows 0..k-1 = Hxinit
  __shared__ int cols[6 * IGV_MAX_BLOCKS];
  __shared__ int s_acc;
  __shared__ double s_Rinv[9], s_S[9];
  if (threadIdx.x == 0) {
    int off = 0;
    for (int q = 0; q < blk.n_blocks; ++q) for (int c = 0; c < blk.size[q]; ++c) cols[off++] = blk.idx[q] + c;
    // per-sequence row count (rows beyond it are zero padding): dof = res.rows() of THIS sequence; no more rows than
    // columns: "H_new rows should be larger than H_new cols" -> not added (StateManager.cpp:574-578)
    const int rows_b = rows_dev ? rows_dev[b] : rows;
    const double thr = rows_dev ? ((rows_b >= 1 && rows_b <= 128) ? chi2_mult * chi2_095[rows_b - 1] : 0.0) : thr_all;
    bool rej = rows_b <= k || (do_chi2 && rows > k && !(gamma[b] <= thr));   // reject if chi2 > mult * quantile
    const double* R = Rf + (size_t)b * 9;
    for (int a = 0; a < k; ++a) if (!(fabs(R[3 * a + a]) > 0.0)) rej = true;   // H_f not invertible
    if (!rej) {   // inverse of the upper triangle by back substitution
      for (int e = 0; e < 9; ++e) s_Rinv[e] = 0.0;
      for (int c = 0; c < k; ++c) {
        s_Rinv[3 * c + c] = 1.0 / R[3 * c + c];
        for (int a = c - 1; a >= 0; --a) {
          double acc = 0.0;
          for (int d = a + 1; d <= c; ++d) acc += R[3 * a + d] * s_Rinv[3 * d + c];
          s_Rinv[3 * a + c] = -acc / R[3 * a + a];
        }
      }
    }
    s_acc = rej ? 0 : 1;
    accepted[b] = s_acc;
  }
  __syncthreads();
  const int n = blk.n;
  const double noise2 = noise2_dev ? noise2_dev[b] : noise2_all;
  for (int t = threadIdx.x; t < N * k; t += blockDim.x) {
    const int i = t % N, a = t / N;
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc = fma(Pb[i + (size_t)cols[c] * ld], Hb[a + (size_t)c * rows], acc);
    sPH[t] = acc;   // (P H_x^T)[i][a]
  }
  __syncthreads();
  if (s_acc) {
    if (threadIdx.x < k * k) {
      const int a = threadIdx.x / k, d = threadIdx.x % k;
      double acc = (a == d) ? noise2 : 0.0;
      for (int c = 0; c < n; ++c) acc = fma(Hb[a + (size_t)c * rows], sPH[cols[c] + d * N], acc);
      s_S[3 * a + d] = acc;
    }
    for (int t = threadIdx.x; t < N * k; t += blockDim.x) {
      const int i = t % N, a = t / N;
      double val = 0.0;
      for (int c = 0; c < k; ++c) val = fma(sPH[i + c * N], s_Rinv[3 * a + c], val);   // (P H^T R^-T)[i][a]
      Pb[i + (size_t)(N + a) * ld] = -val;
      Pb[(N + a) + (size_t)i * ld] = -val;
    }
    __syncthreads();
    if (threadIdx.x < k * k) {
      const int a = threadIdx.x / k, d = threadIdx.x % k;
      double acc = 0.0;
      for (int c = 0; c < k; ++c)
        for (int e = 0; e < k; ++e) acc += s_Rinv[3 * a + c] * 0.5 * (s_S[3 * c + e] + s_S[3 * e + c]) * s_Rinv[3 * d + e];
      Pb[(N + a) + (size_t)(N + d) * ld] = acc;
    }
  } else {
    for (int t = threadIdx.x; t < N * k; t += blockDim.x) {
      const int i = t % N, a = t / N;
      Pb[i + (size_t)(N + a) * ld] = 0.0;
      Pb[(N + a) + (size_t)i * ld] = 0.0;
    }
    if (threadIdx.x < k * k) {
      const int a = threadIdx.x / k, d = threadIdx.x % k;
      Pb[(N + a) + (size_t)(N + d) * ld] = (a == d) ? prior_cov : 0.0;
    }
  }
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

void igv_launch_delayed_init(igv_batch* h, const IgvBlocks& blk, int rows, int k, const double* Hold, const double* Hnew,
                             const double* res, double noise_iso, const double* noise2_dev, const int* rows_dev,
                             double chi2_mult, int do_chi2, double prior_cov, int* accepted_dev) {
  igv_commit_copies(h);
  // workspace: Hw (rows x n) | rw (rows) | Rf (9)  -- carved from Dws (B x dws_stride doubles)
  double* Hw = h->Dws;
  double* rw = Hw + (size_t)h->B * rows * blk.n;
  double* Rf = rw + (size_t)h->B * rows;
  double* gam = h->gam_ws;
  k_delayed_prep<<<h->B, 64, 0, h->stream>>>(rows, blk.n, k, Hold, Hnew, res, Hw, rw, Rf);
  h->launches++;
  if (rows > k) {  // chi^2 of the remaining rows against the prior (StateManager.cpp:604-611)
    IgvEkfLaunch e{};
    e.blk = blk; e.rows = rows - k;
    e.H = Hw + k; e.strideH = (long)rows * blk.n; e.h_ld = rows; e.h_rowmajor = 0;
    e.res = rw + k; e.strideRes = rows; e.res_inc = 1;
    e.R = noise2_dev; e.strideR = noise2_dev ? 1 : 0; e.r_kind = IGV_R_ISO; e.r_iso_value = noise_iso * noise_iso;
    e.gamma_only = 1; e.gamma_out = gam; e.apply_boxplus = 0;
    igv_launch_ekf(h, e);
  }
  // the reference hard-codes the 0.95 quantile at dof = res.rows() here, whatever UpdateBase::_thres the updaters
  // were built with (StateManager.cpp:613-617)
  const double thr = (rows >= 1) ? chi2_mult * igv_chi2_quantile(0.95, rows) : INFINITY;
  const size_t smem = sizeof(double) * (size_t)h->N * k;
  k_delayed_augment<<<h->B, 128, smem, h->stream>>>(h->Pc(), h->ld, h->N, blk, rows, k, Hw, Rf, gam,
                                                    noise_iso * noise_iso, thr, do_chi2, prior_cov, accepted_dev,
                                                    noise2_dev, rows_dev, h->chi2_095, chi2_mult);
  h->launches++;
}

void igv_launch_replace_var_linear(igv_batch* h, int tidx, int tsize, const IgvBlocks& blk, const double* H) {
  igv_commit_copies(h);
  const size_t smem = sizeof(double) * ((size_t)h->N * tsize + tsize * tsize);
  k_replace_var_linear<<<h->B, 128, smem, h->stream>>>(h->Pc(), h->ld, h->N, tidx, tsize, blk, H);
  h->launches++;
}
