// Device-side helpers: 3x3 / so(3) math, CTA-cooperative FP64 GEMM / Cholesky / TRSM, box-plus.
// sm_100a only. FP64 has no tcgen05 kind, so dense contractions here run on the DFMA pipe
// (B200: 64 DFMA/clk/SM); see DESIGN.md "why not tcgen05".
#pragma once
#include <math.h>

#include "igv_internal.h"

namespace igv {

// ---------------------------------------------------------------------------------------------
// 3x3 row-major helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_vec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
__device__ __forceinline__ void mat3T_vec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
__device__ __forceinline__ void skew3(const double* v, double* S) {
  S[0] = 0.0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0.0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0.0;
}

// Gamma_m(phi), m = 0..3.  Reference: AuxGammaFunc.cpp:46-113 (small-angle cut-off 1e-6).
__device__ inline void gamma_func(const double* vec, int m, double* out) {
  const double theta = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  if (fabs(theta) < 1e-6) {
    const double f = (m == 3) ? (1.0 / 6.0) : ((m == 2) ? 0.5 : 1.0);
    for (int i = 0; i < 9; ++i) out[i] = 0.0;
    out[0] = out[4] = out[8] = f;
    return;
  }
  const double n[3] = {vec[0] / theta, vec[1] / theta, vec[2] / theta};
  double nx[9], nx2[9];
  skew3(n, nx);
  mat3_mul(nx, nx, nx2);
  double s, c;
  sincos(theta, &s, &c);
  double f0, f1, f2;
  if (m == 1) {
    f0 = 1.0; f1 = (1.0 - c) / theta; f2 = (theta - s) / theta;
  } else if (m == 2) {
    const double t2 = theta * theta;
    f0 = 0.5; f1 = (theta - s) / t2; f2 = (t2 + 2.0 * c - 2.0) / (2.0 * t2);
  } else if (m == 3) {
    const double t2 = theta * theta, t3 = t2 * theta;
    f0 = 1.0 / 6.0; f1 = (t2 + 2.0 * c - 2.0) / (2.0 * t3); f2 = (t3 - 6.0 * theta + 6.0 * s) / (6.0 * t3);
  } else {
    f0 = 1.0; f1 = s; f2 = 1.0 - c;
  }
  for (int i = 0; i < 9; ++i) out[i] = f1 * nx[i] + f2 * nx2[i];
  out[0] += f0; out[4] += f0; out[8] += f0;
}

// Psi1 / Psi2 (AuxGammaFunc.cpp:115-166, :168-225), reproduced as the reference evaluates them.
__device__ inline void psi_func(const double* w, const double* a, double dt, int which, double* out) {
  const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (wn * fabs(dt) < ((which == 1) ? 1e-8 : 1e-7)) {
    for (int i = 0; i < 9; ++i) out[i] = 0.0;
    return;
  }
  double W[9], A[9], Gm[9], M1[9], nw[3] = {-w[0] * dt, -w[1] * dt, -w[2] * dt};
  skew3(w, W);
  skew3(a, A);
  gamma_func(nw, (which == 1) ? 2 : 3, Gm);
  mat3_mul(A, Gm, M1);
  const double sc = (which == 1) ? dt * dt : dt * dt * dt;
  for (int i = 0; i < 9; ++i) M1[i] *= sc;
  double WA[9], WAW[9], WAW2[9], W2A[9], W2AW[9], W2AW2[9];
  mat3_mul(W, A, WA);
  mat3_mul(WA, W, WAW);
  mat3_mul(WAW, W, WAW2);
  mat3_mul(W, WA, W2A);
  mat3_mul(W2A, W, W2AW);
  mat3_mul(W2AW, W, W2AW2);
  const double eta = wn, xi = eta * dt, xi2 = xi * xi, xi3 = xi * xi2;
  double s1, c1, s2, c2;
  sincos(xi, &s1, &c1);
  sincos(2.0 * xi, &s2, &c2);
  const double e3 = eta * eta * eta, e4 = eta * e3, e5 = eta * e4, e6 = eta * e5, e7 = eta * e6;
  double k1, k2, k3, k4, k5, k6;
  if (which == 1) {
    k1 = (s1 - xi * c1) / e3;
    k2 = (c2 - 4 * c1 + 3) / (4 * e4);
    k3 = (4 * s1 + s2 - 4 * xi * c1 - 2 * xi) / (4 * e5);
    k4 = (xi2 - 2 * xi * s1 - 2 * c1 + 2) / (2 * e4);
    k5 = (6 * xi - 8 * s1 + s2) / (4 * e5);
    k6 = (2 * xi2 - 4 * xi * s1 - c2 + 1) / (4 * e6);
  } else {
    k1 = (xi * s1 + 2 * c1 - 2) / e4;
    k2 = (6 * xi - 8 * s1 + s2) / (8 * e5);
    k3 = (2 * xi2 + 8 * xi * s1 + 16 * c1 + c2 - 17) / (8 * e6);
    k4 = (xi3 + 6 * xi - 12 * s1 + 6 * xi * c1) / (6 * e5);
    k5 = (6 * xi2 + 16 * c1 - c2 - 15) / (8 * e6);
    k6 = (4 * xi3 + 6 * xi - 24 * s1 - 3 * s2 + 24 * xi * c1) / (24 * e7);
  }
  double T[9];
  for (int i = 0; i < 9; ++i)
    T[i] = k1 * WA[i] + k2 * WAW[i] + k3 * WAW2[i] + k4 * W2A[i] + k5 * W2AW[i] + k6 * W2AW2[i];
  mat3_mul(M1, T, out);
}

// ---------------------------------------------------------------------------------------------
// Bulk asynchronous copy global -> shared through the TMA unit (cp.async.bulk, SASS UBLKCP), completion on an mbarrier.
// One elected thread issues the copy; the data lands while the CTA does other work; every thread waits on the
// barrier's phase before the first read.  Addresses and sizes are multiples of 16 bytes.
// ---------------------------------------------------------------------------------------------
#ifndef IGV_EMULATE
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// box-plus on the packed mean (PoseState.cpp:79-88,174-186, VecState.cpp:27-47)
// ---------------------------------------------------------------------------------------------
__device__ inline void retract_pose(double* R, double* p1, double* p2, const double* dth, const double* d1,
                                    const double* d2) {
  double G0[9], G1[9], Rn[9], t[3], u[3];
  gamma_func(dth, 0, G0);
  gamma_func(dth, 1, G1);
  mat3_mul(G0, R, Rn);
  for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  mat3_vec(G0, p1, t);
  mat3_vec(G1, d1, u);
  for (int i = 0; i < 3; ++i) p1[i] = t[i] + u[i];
  if (p2) {
    mat3_vec(G0, p2, t);
    mat3_vec(G1, d2, u);
    for (int i = 0; i < 3; ++i) p2[i] = t[i] + u[i];
  }
}

// One thread per variable applies x <- x [+] dx.  Call with at least (4 + 6 + n_clones) threads
// active, or loop: `for (v = tid; v < nvar; v += nthreads)`.
__device__ inline void boxplus_var(double* X, const double* dx, const IgvLayout& L, int v) {
  if (v == 0) {
    retract_pose(X, X + 9, X + 12, dx, dx + 3, dx + 6);
  } else if (v == 1) {
    for (int i = 0; i < 3; ++i) X[15 + i] += dx[9 + i];
  } else if (v == 2) {
    for (int i = 0; i < 3; ++i) X[18 + i] += dx[12 + i];
  } else if (v == 3) {
    retract_pose(X + 21, X + 30, nullptr, dx + 15, dx + 18, nullptr);
  } else if (v < 10) {
    const int g = v - 4;
    if (L.idx_gnss[g] >= 0) X[33 + g] += dx[L.idx_gnss[g]];
  } else if (v < 10 + L.n_clones) {
    const int s = v - 10;
    double* c = X + IGV_X_CORE + 12 * s;
    const double* d = dx + L.idx_clone[s];
    retract_pose(c, c + 9, nullptr, d, d + 3, nullptr);
  } else {
    // AnchoredLandmark::update (AnchoredLandmark.cpp:227-243): p_f <- Gamma0(dtheta_anchor) p_f + Gamma1(dtheta_anchor) dp
    const int l = v - 10 - L.n_clones;
    if (l < L.n_lm) {
      double* pf = X + L.lm_off + 3 * l;
      const double* dp = dx + L.idx_lm[l];
      const int a = L.lm_anchor[l];
      if (a >= 0 && a < L.n_clones) {
        const double* dth = dx + L.idx_clone[a];
        double G0[9], G1[9], t[3], u[3];
        gamma_func(dth, 0, G0);
        gamma_func(dth, 1, G1);
        mat3_vec(G0, pf, t);
        mat3_vec(G1, dp, u);
        for (int i = 0; i < 3; ++i) pf[i] = t[i] + u[i];
      } else {
        for (int i = 0; i < 3; ++i) pf[i] += dp[i];
      }
    }
  }
}
__device__ inline void boxplus_all(double* X, const double* dx, const IgvLayout& L) {
  for (int v = threadIdx.x; v < 10 + L.n_clones + L.n_lm; v += blockDim.x) boxplus_var(X, dx, L, v);
}

// ---------------------------------------------------------------------------------------------
// CTA-cooperative FP64 GEMM with functor operands:  C(i,j) (op)= sum_k A(i,k) * B(k,j)
// Each thread owns a TM x TN register tile whose rows/cols are INTERLEAVED over the tile grid, so
// that consecutive threads touch consecutive i (coalesced / conflict-free for i-contiguous storage).
// ---------------------------------------------------------------------------------------------
template <int TM, int TN, class FA, class FB, class FC>
__device__ __forceinline__ void cta_gemm(int M, int N, int K, FA a, FB b, FC c) {
  const int gm = (M + TM - 1) / TM, gn = (N + TN - 1) / TN;
  for (int t = threadIdx.x; t < gm * gn; t += blockDim.x) {
    const int ti = t % gm, tj = t / gm;
    int ri[TM], cj[TN];
#pragma unroll
    for (int u = 0; u < TM; ++u) ri[u] = min(ti + u * gm, M - 1);
#pragma unroll
    for (int v = 0; v < TN; ++v) cj[v] = min(tj + v * gn, N - 1);
    double acc[TM][TN];
#pragma unroll
    for (int u = 0; u < TM; ++u)
#pragma unroll
      for (int v = 0; v < TN; ++v) acc[u][v] = 0.0;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      double av[TM], bv[TN];
#pragma unroll
      for (int u = 0; u < TM; ++u) av[u] = a(ri[u], k);
#pragma unroll
      for (int v = 0; v < TN; ++v) bv[v] = b(k, cj[v]);
#pragma unroll
      for (int u = 0; u < TM; ++u)
#pragma unroll
        for (int v = 0; v < TN; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < TM; ++u)
#pragma unroll
      for (int v = 0; v < TN; ++v)
        if (ti + u * gm < M && tj + v * gn < N) c(ti + u * gm, tj + v * gn, acc[u][v]);
  }
}

// ---------------------------------------------------------------------------------------------
// The same contraction on the FP64 tensor pipe: C(i,j) (op)= sum_k A(i,k) * B(k,j) with DMMA.8x8x4
// (mma.sync.m8n8k4.f64). Every warp owns whole (8 TU) x (8 TV) blocks of C, dealt round-robin; per k-step of
// 4 it fetches TU + TV operand values per lane through the functors and issues TU * TV DMMAs (256 FMAs each),
// against TM + TN fetches per TM * TN FMAs of the register-tiled version above.
// Fragment ownership: lane l supplies A(i0 + l/4, k0 + l%4) and B(k0 + l%4, j0 + l/4) and receives
// C(i0 + l/4, j0 + 2 (l%4) + {0,1}).  `skip(i0, j0)` drops whole blocks (e.g. above the diagonal).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// `kbegin(i0, j0)` (a multiple of 4) lets a caller with a triangular operand skip the leading zero part of the
// contraction; `cload(i, j)` is issued BEFORE the k loop and its value handed to `c(i, j, acc, old)`, so a
// read-modify-write of C in global memory overlaps the contraction instead of trailing it.
template <int TU, int TV, class FA, class FB, class FC, class FS, class FK, class FL>
__device__ __forceinline__ void cta_gemm_mma_ex(int M, int N, int K, FA a, FB b, FC c, FS skip, FK kbegin, FL cload) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int fk = lane & 3, fc = lane >> 2;
  constexpr int BM = 8 * TU, BN = 8 * TV;
  const int gm = (M + BM - 1) / BM, gn = (N + BN - 1) / BN;
  for (int t = warp; t < gm * gn; t += nw) {
    const int i0 = (t % gm) * BM, j0 = (t / gm) * BN;
    if (skip(i0, j0)) continue;
    int ia[TU], jb[TV];
#pragma unroll
    for (int u = 0; u < TU; ++u) ia[u] = min(i0 + 8 * u + fc, M - 1);
#pragma unroll
    for (int v = 0; v < TV; ++v) jb[v] = min(j0 + 8 * v + fc, N - 1);
    double acc[TU][TV][2], old[TU][TV][2];
#pragma unroll
    for (int u = 0; u < TU; ++u)
#pragma unroll
      for (int v = 0; v < TV; ++v) {
        acc[u][v][0] = acc[u][v][1] = 0.0;
        const int row = min(i0 + 8 * u + fc, M - 1), col = j0 + 8 * v + 2 * fk;
        old[u][v][0] = cload(row, min(col, N - 1));
        old[u][v][1] = cload(row, min(col + 1, N - 1));
      }
    // four k-steps per trip: all of their operand fetches are issued before the first DMMA, so operands that live in
    // global memory (P, the compressed Jacobian) cost one round trip per 16 contraction indices instead of one per 4
    for (int k0 = kbegin(i0, j0); k0 < K; k0 += 16) {
      double fa[4][TU], fb[4][TV];
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int k = k0 + 4 * s4 + fk;
        const bool vk = k < K;
        const int kk = vk ? k : K - 1;
#pragma unroll
        for (int u = 0; u < TU; ++u) { const double x = a(ia[u], kk); fa[s4][u] = vk ? x : 0.0; }
#pragma unroll
        for (int v = 0; v < TV; ++v) { const double x = b(kk, jb[v]); fb[s4][v] = vk ? x : 0.0; }
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        if (k0 + 4 * s4 < K) {
#pragma unroll
          for (int u = 0; u < TU; ++u)
#pragma unroll
            for (int v = 0; v < TV; ++v) mma884(acc[u][v][0], acc[u][v][1], fa[s4][u], fb[s4][v]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < TU; ++u)
#pragma unroll
      for (int v = 0; v < TV; ++v) {
        const int row = i0 + 8 * u + fc, col = j0 + 8 * v + 2 * fk;
        if (row < M) {
          if (col < N) c(row, col, acc[u][v][0], old[u][v][0]);
          if (col + 1 < N) c(row, col + 1, acc[u][v][1], old[u][v][1]);
        }
      }
  }
}

template <int TU, int TV, class FA, class FB, class FC, class FS>
__device__ __forceinline__ void cta_gemm_mma(int M, int N, int K, FA a, FB b, FC c, FS skip) {
  cta_gemm_mma_ex<TU, TV>(M, N, K, a, b, [&](int i, int j, double v, double) { c(i, j, v); }, skip,
                          [](int, int) { return 0; }, [](int, int) { return 0.0; });
}

// In-place lower Cholesky of the n x n matrix S (column-major, leading dim lds), CTA-cooperative,
// right-looking. Returns false (uniformly) if a non-positive pivot appears. `s_ok` is a shared int.
__device__ inline bool cta_cholesky(double* S, int n, int lds, int* s_ok) {
  if (threadIdx.x == 0) *s_ok = 1;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double d = S[j + (long)j * lds];
      if (!(d > 0.0)) *s_ok = 0;
      S[j + (long)j * lds] = sqrt(d);
    }
    __syncthreads();
    if (!*s_ok) return false;
    const double dj = S[j + (long)j * lds];
    for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) S[i + (long)j * lds] /= dj;
    __syncthreads();
    // trailing update of the lower triangle: S[i,c] -= L[i,j] L[c,j], j < c <= i < n
    const int m = n - j - 1;
    for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
      const int i = j + 1 + t % m, c = j + 1 + t / m;
      if (c <= i) S[i + (long)c * lds] -= S[i + (long)j * lds] * S[c + (long)j * lds];
    }
    __syncthreads();
  }
  return true;
}

// Z <- L^{-1} Z for a rows x ncol column-major Z (leading dim ldz): one thread per column.
__device__ inline void cta_trsm_lower(const double* L, int ldl, double* Z, int rows, int ncol, int ldz) {
  for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
    double* z = Z + (long)c * ldz;
    for (int i = 0; i < rows; ++i) {
      double acc = z[i];
      for (int k = 0; k < i; ++k) acc = fma(-L[i + (long)k * ldl], z[k], acc);
      z[i] = acc / L[i + (long)i * ldl];
    }
  }
}

// Blocked right-looking Cholesky fused with the forward substitution of the right-hand sides:
//   S = L L^T (lower, in place; S is r x r column-major, leading dim lds) and
//   Z[:, c0:c0+nc] <- L^-1 Z[:, c0:c0+nc]   (Z is r x * column-major, leading dim ldz).
// Per block of NB columns: (a) warp 0 factors the NB x NB diagonal block, (b) one thread per row of the
// panel / per right-hand side solves against it, (c) the trailing update of [S | Z] is a K = NB
// register-tiled GEMM over all threads.  3 barriers per block; the right-hand sides ride along, so
// there is no separate triangular solve.  Returns false (uniformly) on a non-positive pivot; with `dref` (the
// original diagonal) the factorisation is the semi-definite one described at (a) and never fails; bit 1 of *s_ok then
// reports a kept pivot below kWeakPivot of its column's original diagonal, i.e. within two decades of the threshold
// under which the column would have been treated as dependent (that row of the factor carries about five digits).
constexpr double kWeakPivot = 1e-11;
#ifndef IGV_EKF_NB
#define IGV_EKF_NB 8    // panel width of the blocked Cholesky (measured at c2: 16 is slower, EKF 0.50 -> 0.63 ms, Gram factor 0.137 -> 0.211 ms)
#endif
__device__ __forceinline__ double rsqrt_nobranch(double x);
template <int NB>
__device__ inline bool cta_chol_solve_fused(double* S, int r, int lds, double* Z, int ldz, int c0, int nc, int* s_ok,
                                            const double* dref = nullptr, double tol = 0.0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_rdiag[NB];   // 1 / L[k][k] of the current diagonal block
  if (tid == 0) *s_ok = 1;
  __syncthreads();
  for (int jb = 0; jb < r; jb += NB) {
    const int nb = min(NB, r - jb);
    double* D = S + jb + (long)jb * lds;   // diagonal block, element (i,k) at D[i + k*r]
    // (a) unblocked Cholesky of the diagonal block by warp 0: lane = row, the row's lower part in REGISTERS, column
    // k's entries exchanged by shuffles -- the other warps wait for this serial section, so it has no shared-memory
    // round trips or warp barriers on its critical path
    if (warp == 0) {
      double arow[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) arow[c] = (lane < nb && c <= lane) ? D[lane + (long)c * lds] : 0.0;
      bool ok = true, weak = false;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        if (k >= nb) break;
        const double d = __shfl_sync(0xffffffffu, arow[k], k);
        // semi-definite mode (dref given): a pivot at rounding level relative to the original diagonal marks a
        // dependent column -- its column of L and its entry of the solved right-hand sides become zero
        const bool dead = dref && !(d > tol * dref[jb + k] && d > 1e-280);
        if (!dref && !(d > 0.0)) { ok = false; break; }
        if (dref && !dead && d < kWeakPivot * dref[jb + k]) weak = true;
        const double inv = dead ? 0.0 : rsqrt_nobranch(d);   // d > 0 and normal here; no slow-path call on the serial chain
        const double l = arow[k] * inv;          // L[lane][k] (lane == k: sqrt(d))
        arow[k] = l;
        if (lane == k) s_rdiag[k] = inv;
#pragma unroll
        for (int c = k + 1; c < NB; ++c) {
          const double lc = __shfl_sync(0xffffffffu, l, c);
          if (c <= lane) arow[c] = fma(-l, lc, arow[c]);
        }
      }
      if (!ok && lane == 0) *s_ok = 0;
      if (ok && weak && lane == 0) *s_ok |= 2;
      if (lane < nb) {
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c <= lane && c < nb) D[lane + (long)c * lds] = arow[c];
      }
    }
    __syncthreads();
    if (!*s_ok) break;
    const int i0 = jb + nb, m = r - i0;  // trailing rows
    // (b) panel rows: L21[i,:] = A21[i,:] L11^-T ; right-hand sides: Y1 = L11^-1 Z1
    for (int t = tid; t < m + nc; t += blockDim.x) {
      double x[NB];
      if (t < m) {
        double* row = S + (i0 + t) + (long)jb * lds;   // element k at row[k*r]
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < nb) {
            double acc = row[(long)k * lds];
#pragma unroll
            for (int p = 0; p < NB; ++p) if (p < k) acc = fma(-x[p], D[k + (long)p * lds], acc);
            x[k] = acc * s_rdiag[k];
          }
#pragma unroll
        for (int k = 0; k < NB; ++k) if (k < nb) row[(long)k * lds] = x[k];
      } else {
        double* col = Z + jb + (long)(c0 + t - m) * ldz;   // element k at col[k]
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < nb) {
            double acc = col[k];
#pragma unroll
            for (int p = 0; p < NB; ++p) if (p < k) acc = fma(-D[k + (long)p * lds], x[p], acc);
            x[k] = acc * s_rdiag[k];
          }
#pragma unroll
        for (int k = 0; k < NB; ++k) if (k < nb) col[k] = x[k];
      }
    }
    __syncthreads();
    // (c) trailing update of rows [i0, r) x columns [S: i0..r-1 (lower part) | Z: c0..c0+nc-1], K = nb, on DMMA
    if (m > 0) {
      const int ncol = m + nc;
      cta_gemm_mma<2, 2>(
          m, ncol, nb, [&](int il, int k) { return S[(i0 + il) + (long)(jb + k) * lds]; },
          [&](int k, int cl) {
            return (cl < m) ? S[(i0 + cl) + (long)(jb + k) * lds] : Z[jb + k + (long)(c0 + cl - m) * ldz];
          },
          [&](int il, int cl, double v) {
            const int i = i0 + il;
            if (cl < m) {
              if (cl <= il) S[i + (long)(i0 + cl) * lds] -= v;
            } else {
              Z[i + (long)(c0 + cl - m) * ldz] -= v;
            }
          },
          [&](int bi, int bj) { return bj + 16 <= m && bj > bi + 15; });
    }
    __syncthreads();
  }
  __syncthreads();
  return *s_ok != 0;
}

// Branch-free reciprocal / reciprocal square root for normal positive arguments: MUFU seed (~2^-21) +
// Newton steps. Unlike 1.0 / x, rsqrt(x) and __drcp_rn(x) there is no slow-path call, so the compiler can
// schedule the chain underneath independent work in the same basic block.
__device__ __forceinline__ double rsqrt_nobranch(double x) {   // x normal, > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // ~2^-21 seed
  const double e = fma(-x, y * y, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);                    // cubic step -> ~1 ulp
}
__device__ __forceinline__ double rcp_nobranch(double x) {     // x normal, > 0
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// as rcp_nobranch without the final rounding-correction step (relative error ~2^-52 .. 2^-51)
__device__ __forceinline__ double rcp_short(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace igv
