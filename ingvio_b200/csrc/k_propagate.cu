// IMU propagation: mean + analytic (Phi, G) + covariance strip update, all K steps of a frame in
// one launch with the strip resident in shared memory.
//
// Reference: ImuPropagator::stateAndCovTransition (ImuPropagator.cpp:98-162, analytic branch),
// the loop body of propagateUntil (:260-271) and StateManager::propagateStateCov
// (StateManager.cpp:42-119).  The reference forms Phi_bar P Phi_bar^T on the full N x N matrix with
// four N x N temporaries per IMU sample; only the rows/cols of the 15 IMU states and of the clock
// biases change, plus the clock-drift row through Q, so this kernel keeps that (<= 20) x N strip
// in SMEM across all steps and touches HBM once per frame: 2*nS*N*8 bytes read+write.
#include "igv_device.cuh"

using namespace igv;

namespace {

constexpr int kMaxStrip = 20;  // 15 IMU + 4 clock biases + clock drift
constexpr int kPreStride = 408; // doubles per (sequence, IMU sample) transition record: 225 + 180, padded to 16 bytes

struct PropArgs {
  double* P; int ld; int N;
  double* X; int xsize;
  int n_steps;
  const double* gyro; const double* accel; const double* dt;  // IMU mode (Phi == nullptr)
  const double* Phi; const double* G;                         // covariance-only mode
  const double* pre;                                          // IMU mode: per (b, step) 225 Phi + 180 G*sigma, row-major, stride kPreStride
  int stage_pre;                                              // fetch the sequence's records with one bulk TMA copy
  IgvDevParams prm;
  int idx_cb[4]; int idx_fs; int enable_gnss;
};

// One IMU sample in three pieces (ImuPropagator::stateAndCovTransition, ImuPropagator.cpp:98-162, analytic branch):
//   imu_gammas : Gamma_0,1,2(w dt) -- depends on the sample only                       (any lane, any order)
//   imu_mean   : R,p,v <- ... (:126-137) and the clock biases (:139-148)               (sequential over samples)
//   imu_phi_g  : Phi (15x15, :150-161) and G diag(sigma) (15x12, :112-117, StateManager.cpp:92-96) from the state
//                BEFORE the sample, the sample and its Gammas                          (any lane, any order)
// Phi / G go to a zero-initialised workspace whose sparsity pattern never changes, so only non-zeros are written.
__device__ __forceinline__ void imu_gammas(const double* wraw, const double* bg, double dt, double* G012) {
  const double wd[3] = {(wraw[0] - bg[0]) * dt, (wraw[1] - bg[1]) * dt, (wraw[2] - bg[2]) * dt};
  gamma_func(wd, 0, G012); gamma_func(wd, 1, G012 + 9); gamma_func(wd, 2, G012 + 18);
}

__device__ __forceinline__ void imu_mean(double* X, const double* araw, double dt, const IgvDevParams& prm,
                                         const int* idx_cb, int idx_fs, const double* G012) {
  double* R = X; double* p = X + 9; double* v = X + 12;
  const double* ba = X + 18;
  const double a[3] = {araw[0] - ba[0], araw[1] - ba[1], araw[2] - ba[2]};
  double Rn[9], RG1[9], RG2[9], t1[3], t2[3];
  mat3_mul(R, G012, Rn); mat3_mul(R, G012 + 9, RG1); mat3_mul(R, G012 + 18, RG2);
  mat3_vec(RG1, a, t1);
  mat3_vec(RG2, a, t2);
  const double* g = prm.g;
  for (int i = 0; i < 3; ++i) {
    const double vh = v[i];
    v[i] = vh + g[i] * dt + t1[i] * dt;
    p[i] = p[i] + vh * dt + 0.5 * g[i] * dt * dt + t2[i] * dt * dt;
  }
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  if (idx_fs >= 0)  // ImuPropagator.cpp:139-148
    for (int i = 0; i < 4; ++i) if (idx_cb[i] >= 0) X[33 + i] += dt * X[33 + 4];
}

// pre: R (9), p (3), v (3) before the sample; post: p, v after it (X + 9 of the next state)
__device__ void imu_phi_g(const double* pre, const double* pn, const double* vn, const double* bg, const double* ba,
                          const double* wraw, const double* araw, double dt, const IgvDevParams& prm,
                          const double* G012, double* sPhi, double* sG) {
  const double* Rh = pre; const double* ph = pre + 9; const double* vh = pre + 12;
  for (int i = 0; i < 15; ++i) sPhi[16 * i] = 1.0;
  double S[9], T[9];
  skew3(ph, S); mat3_mul(S, Rh, T);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    sG[r * 12 + c] = Rh[3 * r + c] * prm.noise_g;
    sG[(3 + r) * 12 + c] = T[3 * r + c] * prm.noise_g;
  }
  skew3(vh, S); mat3_mul(S, Rh, T);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    sG[(6 + r) * 12 + c] = T[3 * r + c] * prm.noise_g;
    sG[(6 + r) * 12 + 3 + c] = Rh[3 * r + c] * prm.noise_a;
  }
  for (int r = 0; r < 3; ++r) { sG[(9 + r) * 12 + 6 + r] = prm.noise_bg; sG[(12 + r) * 12 + 9 + r] = prm.noise_ba; }
  const double w[3] = {wraw[0] - bg[0], wraw[1] - bg[1], wraw[2] - bg[2]};
  const double a[3] = {araw[0] - ba[0], araw[1] - ba[1], araw[2] - ba[2]};
  double RG1[9], RG2[9];
  mat3_mul(Rh, G012 + 9, RG1); mat3_mul(Rh, G012 + 18, RG2);
  const double* g = prm.g;
  // Phi blocks (ImuPropagator.cpp:150-161)
  double Sg[9];
  skew3(g, Sg);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    sPhi[(3 + r) * 15 + c] = 0.5 * Sg[3 * r + c] * dt * dt;
    sPhi[(6 + r) * 15 + c] = Sg[3 * r + c] * dt;
    sPhi[r * 15 + 9 + c] = -RG1[3 * r + c] * dt;
    sPhi[(6 + r) * 15 + 12 + c] = -RG1[3 * r + c] * dt;
    sPhi[(3 + r) * 15 + 12 + c] = -RG2[3 * r + c] * dt * dt;
  }
  for (int r = 0; r < 3; ++r) sPhi[(3 + r) * 15 + 6 + r] = dt;
  double Ps1[9], Ps2[9], A1[9], A2[9];
  psi_func(w, a, dt, 1, Ps1); psi_func(w, a, dt, 2, Ps2);
  skew3(vn, S); mat3_mul(S, RG1, A1); mat3_mul(Rh, Ps1, A2);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) sPhi[(6 + r) * 15 + 9 + c] = -A1[3 * r + c] * dt + A2[3 * r + c];
  skew3(pn, S); mat3_mul(S, RG1, A1); mat3_mul(Rh, Ps2, A2);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) sPhi[(3 + r) * 15 + 9 + c] = -A1[3 * r + c] * dt + A2[3 * r + c];
}

// One WARP per sequence. The K samples of a frame are sequential only through the mean: lanes first evaluate the
// Gamma functions of their sample, lane 0 then runs the short mean recursion (keeping every intermediate state),
// and the lanes finally build Phi_k and G_k sigma of their sample for the strip kernel.
constexpr int kImuWarps = 2, kImuChunk = 32;
__global__ void __launch_bounds__(kImuWarps * 32) k_imu_mean(double* X, int xsize, int B, int n_steps, const double* gyro,
                                                             const double* accel, const double* dt, IgvDevParams prm,
                                                             int idx_cb0, int idx_cb1, int idx_cb2, int idx_cb3, int idx_fs,
                                                             double* pre) {
  __shared__ double s_gam[kImuWarps][kImuChunk][27];
  __shared__ double s_st[kImuWarps][kImuChunk + 1][15];   // R, p, v before sample k (entry kImuChunk: after the last)
  __shared__ double s_x[kImuWarps][IGV_X_CORE];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kImuWarps + wib;
  if (b >= B) return;
  const int idx_cb[4] = {idx_cb0, idx_cb1, idx_cb2, idx_cb3};
  double* Xb = X + (size_t)b * xsize;
  double* Xl = s_x[wib];
  for (int i = lane; i < IGV_X_CORE; i += 32) Xl[i] = Xb[i];
  __syncwarp();
  for (int base = 0; base < n_steps; base += kImuChunk) {
    const int cnt = min(kImuChunk, n_steps - base);
    const size_t o = (size_t)b * n_steps + base + lane;
    const double d = (lane < cnt) ? dt[o] : 0.0;
    const bool live = d >= 1e-6;   // ImuPropagator.cpp:262
    if (live) imu_gammas(gyro + o * 3, Xl + 15, d, s_gam[wib][lane]);
    __syncwarp();
    if (lane == 0) {
      for (int k = 0; k < cnt; ++k) {
        for (int i = 0; i < 15; ++i) s_st[wib][k][i] = Xl[i];
        const size_t ok = (size_t)b * n_steps + base + k;
        const double dk = dt[ok];
        if (dk >= 1e-6) imu_mean(Xl, accel + ok * 3, dk, prm, idx_cb, idx_fs, s_gam[wib][k]);
      }
      for (int i = 0; i < 15; ++i) s_st[wib][cnt][i] = Xl[i];
    }
    __syncwarp();
    if (live)
      imu_phi_g(s_st[wib][lane], s_st[wib][lane + 1] + 9, s_st[wib][lane + 1] + 12, Xl + 15, Xl + 18, gyro + o * 3, accel + o * 3,
                d, prm, s_gam[wib][lane], pre + o * kPreStride, pre + o * kPreStride + 225);
    __syncwarp();
  }
  for (int i = lane; i < 15; i += 32) Xb[i] = Xl[i];                  // R, p, v
  if (lane < 4) Xb[33 + lane] = Xl[33 + lane];                       // clock biases (ImuPropagator.cpp:139-148)
}

__global__ void __launch_bounds__(128) k_propagate(PropArgs a) {
  extern __shared__ double sm[];   // (dynamic shared memory starts 16-byte aligned: the bulk copy below relies on it)
  constexpr int S = kMaxStrip;  // fixed stride of every small matrix: index math folds to shifts/multiplies
  const int b = blockIdx.x, N = a.N, ld = a.ld, tid = threadIdx.x;
  double* Pb = a.P + (size_t)b * ld * ld;
  __shared__ int cidx[S];
  __shared__ int s_nC, s_nS, s_has_fs;
  __shared__ __align__(16) double sPhi[225], sG[180], sT[S * S], sQ[S * S], sM[15 * 12];
  __shared__ __align__(16) double sBk[S * S], sX[S * S], sY[S * S], sTt[S * S], sTn[S * S];
  __shared__ double s_dt;
  // IMU mode: the transition records of ALL samples of this sequence (n_steps x kPreStride doubles, one contiguous block
  // written by k_imu_mean) are fetched by ONE bulk TMA copy issued here and landing behind the strip in shared memory
  // while the strip itself is loaded; the per-sample loop then reads them in place (no per-sample global round trip).
  __shared__ __align__(8) unsigned long long s_bar;
  double* stage = sm + (size_t)S * N + (((size_t)S * N) & 1);   // 16-byte aligned
  const bool staged = a.stage_pre != 0;
#ifndef IGV_EMULATE
  if (staged && tid == 0) mbar_init(&s_bar, 1);
#endif
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < 15; ++i) cidx[n++] = i;
    s_has_fs = 0;
    if (a.enable_gnss)
      for (int i = 0; i < 4; ++i) if (a.idx_cb[i] >= 0) cidx[n++] = a.idx_cb[i];
    s_nC = n;
    if (a.enable_gnss && a.idx_fs >= 0) { cidx[n++] = a.idx_fs; s_has_fs = 1; }
    s_nS = n;
    for (int i = n; i < S; ++i) cidx[i] = 0;
  }
  __syncthreads();
  if (staged) {
    const double* src = a.pre + (size_t)b * a.n_steps * kPreStride;
    const unsigned bytes = (unsigned)(a.n_steps * kPreStride * sizeof(double));
#ifndef IGV_EMULATE
    if (tid == 0) {
      mbar_expect_tx(&s_bar, bytes);
      bulk_g2s(stage, src, bytes, &s_bar);
    }
#else
    for (int t = tid; t < a.n_steps * kPreStride; t += blockDim.x) stage[t] = src[t];
#endif
  }
  const int nC = s_nC, nS = s_nS;
  const bool has_fs = s_has_fs != 0;
  double* W = sm;  // S x N row-major (rows >= nS are zero): W[r*N + j] = P[cidx[r], j]
  for (int j = tid; j < N; j += blockDim.x) {   // all S loads of a column in flight before the first store
    double col[S];
#pragma unroll
    for (int r = 0; r < S; ++r) col[r] = (r < nS) ? Pb[j + (size_t)cidx[r] * ld] : 0.0;  // P symmetric: row == column (coalesced)
#pragma unroll
    for (int r = 0; r < S; ++r) W[r * N + j] = col[r];
  }
  __syncthreads();
  // Per IMU sample the reference updates P11 <- sym(T P11 T^T + Q) and P21 <- P21 T^T (StateManager.cpp:51-118).
  // The strip x strip block Bk = P11 follows that recursion step by step (including the per-step symmetrisation);
  // the other columns of the strip only ever see the left factor, w <- T_k w, so the product T_K ... T_1 is
  // accumulated (S x S) and applied to them ONCE after the last sample: per-step cost O(S^3), not O(S^2 N).
  for (int t = tid; t < S * S; t += blockDim.x) {
    const int r = t / S, c = t % S;
    sBk[t] = (r < nS && c < nS) ? W[r * N + cidx[c]] : 0.0;
    sTt[t] = (r == c) ? 1.0 : 0.0;
  }
#ifndef IGV_EMULATE
  if (staged) mbar_wait(&s_bar, 0);   // the bulk copy has landed (it overlapped the strip load above)
#endif
  __syncthreads();
  for (int step = 0; step < a.n_steps; ++step) {
    const double* cPhi = sPhi;     // this sample's Phi (15 x 15) and G diag(sigma) (15 x 12), row-major
    const double* cG = sG;
    // load this step's Phi (15x15) and G*diag(sigma) (15x12), row-major, all threads
    if (a.Phi) {
      const double* Ph = a.Phi + (size_t)b * 225;
      const double* Gg = a.G + (size_t)b * 180;
      const double sg[4] = {a.prm.noise_g, a.prm.noise_a, a.prm.noise_bg, a.prm.noise_ba};
      for (int t = tid; t < 225; t += blockDim.x) sPhi[t] = Ph[(t / 15) + 15 * (t % 15)];
      for (int t = tid; t < 180; t += blockDim.x) sG[t] = Gg[(t / 12) + 15 * (t % 12)] * sg[(t % 12) / 3];
      if (tid == 0) s_dt = a.dt[b];
    } else if (staged) {
      cPhi = stage + (size_t)step * kPreStride;
      cG = cPhi + 225;
      if (tid == 0) s_dt = a.dt[(size_t)b * a.n_steps + step];
    } else {
      const double* pr = a.pre + ((size_t)b * a.n_steps + step) * kPreStride;
      for (int t = tid; t < 225; t += blockDim.x) sPhi[t] = pr[t];
      for (int t = tid; t < 180; t += blockDim.x) sG[t] = pr[225 + t];
      if (tid == 0) s_dt = a.dt[(size_t)b * a.n_steps + step];
    }
    __syncthreads();
    const double dt = s_dt;
    if (!a.Phi && dt < 1e-6) { __syncthreads(); continue; }  // uniform (ImuPropagator.cpp:262)
    // small transition T (S x S, zero padded): [Phi 0; 0 I] + dt on (cb, fs); rows >= nC are identity rows
    for (int t = tid; t < S * S; t += blockDim.x) {
      const int r = t / S, c = t % S;
      double val = (r == c && r < nS) ? 1.0 : 0.0;
      if (r < 15 && c < 15) val = cPhi[r * 15 + c];
      else if (r >= 15 && r < nC && c == nS - 1 && has_fs) val = dt;
      sT[t] = val;
    }
    __syncthreads();
    // the S x S products run on DMMA (cta_gemm_mma, 8 x 8 tiles dealt to the 4 warps): the scalar version spent
    // two shared-memory loads per FMA and was bound by shared-memory bandwidth
    auto none = [](int, int) { return false; };
    // M = Phi * Gs (15 x 12)
    cta_gemm_mma<1, 1>(15, 12, 15, [&](int i, int k) { return cPhi[i * 15 + k]; }, [&](int k, int j) { return cG[k * 12 + j]; },
                       [&](int i, int j, double v) { sM[i * 12 + j] = v; }, none);
    // X = Bk T^T and the accumulated transition Tn = T Tt
    cta_gemm_mma<1, 1>(S, S, S, [&](int i, int k) { return sBk[i * S + k]; }, [&](int k, int j) { return sT[j * S + k]; },
                       [&](int i, int j, double v) { sX[i * S + j] = v; }, none);
    cta_gemm_mma<1, 1>(S, S, S, [&](int i, int k) { return sT[i * S + k]; }, [&](int k, int j) { return sTt[k * S + j]; },
                       [&](int i, int j, double v) { sTn[i * S + j] = v; }, none);
    __syncthreads();
    // Q block: dt * M M^T on the IMU part (StateManager.cpp:97) + clock terms (:99-116)
    cta_gemm_mma<1, 1>(15, 15, 12, [&](int i, int k) { return sM[i * 12 + k]; }, [&](int k, int j) { return sM[j * 12 + k]; },
                       [&](int i, int j, double v) { sQ[i * S + j] = v * dt; }, none);
    for (int t = tid; t < S * S; t += blockDim.x) {
      const int r = t / S, c = t % S;
      if (r < 15 && c < 15) continue;
      double q = 0.0;
      if (a.enable_gnss && r >= 15 && c >= 15 && r < nS && c < nS) {
        const bool rf = has_fs && (r == nS - 1), cf = has_fs && (c == nS - 1);
        const double rw2 = a.prm.noise_cb_rw * a.prm.noise_cb_rw;
        if (!rf && !cf) q = dt * a.prm.noise_cb * a.prm.noise_cb + dt * dt * dt * rw2;
        else if (rf && cf) q = dt * rw2;
        else q = dt * dt * rw2;
      }
      sQ[t] = q;
    }
    // Y = T X
    cta_gemm_mma<1, 1>(S, S, S, [&](int i, int k) { return sT[i * S + k]; }, [&](int k, int j) { return sX[k * S + j]; },
                       [&](int i, int j, double v) { sY[i * S + j] = v; }, none);
    for (int t = tid; t < S * S; t += blockDim.x) sTt[t] = sTn[t];
    __syncthreads();
    for (int t = tid; t < S * S; t += blockDim.x) {   // + Q, symmetrise (StateManager.cpp:118)
      const int r = t / S, c = t % S;
      sBk[t] = (r < nS && c < nS) ? 0.5 * ((sY[r * S + c] + sQ[r * S + c]) + (sY[c * S + r] + sQ[c * S + r])) : 0.0;
    }
    __syncthreads();
  }
  // apply the accumulated transition to the columns outside the strip (one thread per column, the column in
  // registers), then drop the strip x strip block in
  for (int j = tid; j < N; j += blockDim.x) {
    bool in_strip = false;
    for (int c = 0; c < nS; ++c) in_strip |= (cidx[c] == j);
    if (in_strip) continue;
    double col[S];
#pragma unroll
    for (int k = 0; k < S; ++k) col[k] = W[k * N + j];
#pragma unroll 4
    for (int c = 0; c < S; ++c) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int k = 0; k < S; k += 2) {
        const double2 tv = *reinterpret_cast<const double2*>(&sTt[c * S + k]);
        acc0 = fma(tv.x, col[k], acc0);
        acc1 = fma(tv.y, col[k + 1], acc1);
      }
      if (c < nS) W[c * N + j] = acc0 + acc1;
    }
  }
  for (int t = tid; t < S * S; t += blockDim.x) {
    const int r = t / S, c = t % S;
    if (r < nS && c < nS) W[r * N + cidx[c]] = sBk[t];
  }
  __syncthreads();
  // write back rows and mirrored columns
  for (int r = 0; r < nS; ++r)
    for (int j = tid; j < N; j += blockDim.x) {
      const double val = W[r * N + j];
      Pb[j + (size_t)cidx[r] * ld] = val;
      Pb[cidx[r] + (size_t)j * ld] = val;
    }
}

}  // namespace

void igv_launch_propagate(igv_batch* h, int n_steps, const double* gyro, const double* accel, const double* dt,
                          const double* Phi, const double* G) {
  IgvProfScope prof_scope_(h, IGV_K_PROPAGATE);
  PropArgs a;
  a.P = h->Pc(); a.ld = h->ld; a.N = h->N;
  a.X = h->Xc(); a.xsize = h->xsize;
  a.n_steps = n_steps;
  a.gyro = gyro; a.accel = accel; a.dt = dt; a.Phi = Phi; a.G = G; a.pre = nullptr; a.stage_pre = 0;
  a.prm = h->params;
  IgvLayout L = h->layout();
  for (int i = 0; i < 4; ++i) a.idx_cb[i] = L.idx_gnss[i];
  a.idx_fs = L.idx_gnss[IGV_GNSS_FS];
  a.enable_gnss = 1;
  if (!Phi) {
    const size_t need = (size_t)h->B * n_steps * kPreStride;
    if (need > h->pre_cap) {
      if (h->pre_ws) { cudaStreamSynchronize(h->stream); cudaFree(h->pre_ws); }
      cudaMalloc(reinterpret_cast<void**>(&h->pre_ws), sizeof(double) * need);
      cudaMemsetAsync(h->pre_ws, 0, sizeof(double) * need, h->stream);
      h->pre_cap = need;
    }
    a.pre = h->pre_ws;
    k_imu_mean<<<(h->B + kImuWarps - 1) / kImuWarps, kImuWarps * 32, 0, h->stream>>>(h->Xc(), h->xsize, h->B, n_steps, gyro, accel, dt, h->params,
                                                      a.idx_cb[0], a.idx_cb[1], a.idx_cb[2], a.idx_cb[3], a.idx_fs,
                                                      h->pre_ws);
    h->launches++;
  }
  size_t smem = sizeof(double) * kMaxStrip * h->N;
  // IMU mode: stage the whole sequence's transition records behind the strip with one bulk TMA copy when they fit
  const size_t stage_bytes = sizeof(double) * (size_t)n_steps * kPreStride;
  a.stage_pre = (!Phi && h->knobs.prop_tma != 0 && smem + 16 + stage_bytes <= 160 * 1024) ? 1 : 0;
  if (a.stage_pre) smem += 16 + stage_bytes;
  IGV_SMEM_OPTIN((k_propagate), 168 * 1024);   // strip: 20 x 512 x 8 B at the largest max_dim igv_create accepts (+ staging)
  k_propagate<<<h->B, 128, smem, h->stream>>>(a);
  h->launches++;
}
