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

struct PropArgs {
  double* P; int ld; int N;
  double* X; int xsize;
  int n_steps;
  const double* gyro; const double* accel; const double* dt;  // IMU mode (Phi == nullptr)
  const double* Phi; const double* G;                         // covariance-only mode
  const double* pre;                                          // IMU mode: per (b, step) 225 Phi + 180 G*sigma, row-major
  IgvDevParams prm;
  int idx_cb[4]; int idx_fs; int enable_gnss;
};

// thread 0: mean propagation and Phi (15x15, row-major in sPhi), Gs = G*diag(sigma) (15x12 row-major)
__device__ void imu_step_mean(double* X, const double* wraw, const double* araw, double dt, const IgvDevParams& prm,
                              const int* idx_cb, int idx_fs, double* sPhi, double* sG, bool zero_fill = true) {
  double* R = X; double* p = X + 9; double* v = X + 12;
  const double* bg = X + 15; const double* ba = X + 18;
  if (zero_fill) {  // the (b, step) workspace is zeroed once at allocation: the sparsity pattern never changes
    for (int i = 0; i < 225; ++i) sPhi[i] = 0.0;
    for (int i = 0; i < 180; ++i) sG[i] = 0.0;
  }
  for (int i = 0; i < 15; ++i) sPhi[16 * i] = 1.0;
  double Rh[9], ph[3], vh[3], S[9], T[9];
  for (int i = 0; i < 9; ++i) Rh[i] = R[i];
  for (int i = 0; i < 3; ++i) { ph[i] = p[i]; vh[i] = v[i]; }
  // G (ImuPropagator.cpp:112-117), scaled by the noise sigmas (StateManager.cpp:92-96)
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
  const double wd[3] = {w[0] * dt, w[1] * dt, w[2] * dt};
  double G0[9], G1[9], G2[9], RG1[9], RG2[9], Rn[9], t[3];
  gamma_func(wd, 0, G0); gamma_func(wd, 1, G1); gamma_func(wd, 2, G2);
  mat3_mul(Rh, G0, Rn); mat3_mul(Rh, G1, RG1); mat3_mul(Rh, G2, RG2);
  const double* g = prm.g;
  double vn[3], pn[3];
  mat3_vec(RG1, a, t);
  for (int i = 0; i < 3; ++i) vn[i] = vh[i] + g[i] * dt + t[i] * dt;
  mat3_vec(RG2, a, t);
  for (int i = 0; i < 3; ++i) pn[i] = ph[i] + vh[i] * dt + 0.5 * g[i] * dt * dt + t[i] * dt * dt;
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  for (int i = 0; i < 3; ++i) { p[i] = pn[i]; v[i] = vn[i]; }
  if (idx_fs >= 0)  // ImuPropagator.cpp:139-148
    for (int i = 0; i < 4; ++i) if (idx_cb[i] >= 0) X[33 + i] += dt * X[33 + 4];
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

// One THREAD per sequence: the K mean-propagation steps are sequential within a sequence but independent
// across sequences; Phi_k and G_k*sigma go to a workspace that the strip kernel consumes.
__global__ void __launch_bounds__(64) k_imu_mean(double* X, int xsize, int B, int n_steps, const double* gyro,
                                                   const double* accel, const double* dt, IgvDevParams prm,
                                                   int idx_cb0, int idx_cb1, int idx_cb2, int idx_cb3, int idx_fs,
                                                   double* pre) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int idx_cb[4] = {idx_cb0, idx_cb1, idx_cb2, idx_cb3};
  double* Xb = X + (size_t)b * xsize;
  double Xl[IGV_X_CORE];  // R, p, v, bg, ba, (extrinsics untouched), GNSS scalars: keep on chip across the K steps
  for (int i = 0; i < IGV_X_CORE; ++i) Xl[i] = Xb[i];
  for (int step = 0; step < n_steps; ++step) {
    const size_t o = (size_t)b * n_steps + step;
    const double d = dt[o];
    if (d >= 1e-6)  // ImuPropagator.cpp:262
      imu_step_mean(Xl, gyro + o * 3, accel + o * 3, d, prm, idx_cb, idx_fs, pre + o * 405, pre + o * 405 + 225, false);
  }
  for (int i = 0; i < 15; ++i) Xb[i] = Xl[i];          // R, p, v
  for (int i = 33; i < 37; ++i) Xb[i] = Xl[i];         // clock biases (ImuPropagator.cpp:139-148)
}

__global__ void __launch_bounds__(128) k_propagate(PropArgs a) {
  extern __shared__ double sm[];
  constexpr int S = kMaxStrip;  // fixed stride of every small matrix: index math folds to shifts/multiplies
  const int b = blockIdx.x, N = a.N, ld = a.ld, tid = threadIdx.x;
  double* Pb = a.P + (size_t)b * ld * ld;
  __shared__ int cidx[S];
  __shared__ int s_nC, s_nS, s_has_fs;
  __shared__ __align__(16) double sPhi[225], sG[180], sT[S * S], sQ[S * S], sM[15 * 12];
  __shared__ __align__(16) double sBk[S * S], sX[S * S], sY[S * S], sTt[S * S], sTn[S * S];
  __shared__ double s_dt;
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
  const int nC = s_nC, nS = s_nS;
  const bool has_fs = s_has_fs != 0;
  double* W = sm;  // S x N row-major (rows >= nS are zero): W[r*N + j] = P[cidx[r], j]
  for (int r = 0; r < S; ++r)
    for (int j = tid; j < N; j += blockDim.x)
      W[r * N + j] = (r < nS) ? Pb[j + (size_t)cidx[r] * ld] : 0.0;  // P symmetric: row == column (coalesced)
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
  __syncthreads();
  for (int step = 0; step < a.n_steps; ++step) {
    // load this step's Phi (15x15) and G*diag(sigma) (15x12), row-major, all threads
    if (a.Phi) {
      const double* Ph = a.Phi + (size_t)b * 225;
      const double* Gg = a.G + (size_t)b * 180;
      const double sg[4] = {a.prm.noise_g, a.prm.noise_a, a.prm.noise_bg, a.prm.noise_ba};
      for (int t = tid; t < 225; t += blockDim.x) sPhi[t] = Ph[(t / 15) + 15 * (t % 15)];
      for (int t = tid; t < 180; t += blockDim.x) sG[t] = Gg[(t / 12) + 15 * (t % 12)] * sg[(t % 12) / 3];
      if (tid == 0) s_dt = a.dt[b];
    } else {
      const double* pr = a.pre + ((size_t)b * a.n_steps + step) * 405;
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
      if (r < 15 && c < 15) val = sPhi[r * 15 + c];
      else if (r >= 15 && r < nC && c == nS - 1 && has_fs) val = dt;
      sT[t] = val;
    }
    __syncthreads();
    // the S x S products run on DMMA (cta_gemm_mma, 8 x 8 tiles dealt to the 4 warps): the scalar version spent
    // two shared-memory loads per FMA and was bound by shared-memory bandwidth
    auto none = [](int, int) { return false; };
    // M = Phi * Gs (15 x 12)
    cta_gemm_mma<1, 1>(15, 12, 15, [&](int i, int k) { return sPhi[i * 15 + k]; }, [&](int k, int j) { return sG[k * 12 + j]; },
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
  a.gyro = gyro; a.accel = accel; a.dt = dt; a.Phi = Phi; a.G = G; a.pre = nullptr;
  a.prm = h->params;
  IgvLayout L = h->layout();
  for (int i = 0; i < 4; ++i) a.idx_cb[i] = L.idx_gnss[i];
  a.idx_fs = L.idx_gnss[IGV_GNSS_FS];
  a.enable_gnss = 1;
  if (!Phi) {
    const size_t need = (size_t)h->B * n_steps * 405;
    if (need > h->pre_cap) {
      if (h->pre_ws) { cudaStreamSynchronize(h->stream); cudaFree(h->pre_ws); }
      cudaMalloc(reinterpret_cast<void**>(&h->pre_ws), sizeof(double) * need);
      cudaMemsetAsync(h->pre_ws, 0, sizeof(double) * need, h->stream);
      h->pre_cap = need;
    }
    a.pre = h->pre_ws;
    k_imu_mean<<<(h->B + 63) / 64, 64, 0, h->stream>>>(h->Xc(), h->xsize, h->B, n_steps, gyro, accel, dt, h->params,
                                                      a.idx_cb[0], a.idx_cb[1], a.idx_cb[2], a.idx_cb[3], a.idx_fs,
                                                      h->pre_ws);
    h->launches++;
  }
  const size_t smem = sizeof(double) * kMaxStrip * h->N;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_propagate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set = true;
  }
  k_propagate<<<h->B, 128, smem, h->stream>>>(a);
  h->launches++;
}
