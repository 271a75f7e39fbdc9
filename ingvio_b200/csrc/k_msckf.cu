// MSCKF per-feature kernel: residual + Jacobian over the clone poses, left null-space projection of
// the landmark, chi^2 gate.  One WARP per feature; lanes own columns of the projected block, so the
// write of the (M-3) x (n+1) block to HBM is coalesced; clone poses and the window's marginal
// covariance P_s are staged in shared memory once per CTA.
//
// Reference: RemoveLostUpdate::calcResJacobianSingleFeatAll{Mono,Stereo}Obs
// (RemoveLostUpdate.cpp:169-273, :407-523), KeyframeUpdate::calcResJacobianSingleFeatSelected*
// (KeyframeUpdate.cpp:160-249, :330-436), SwMargUpdate (SwMargUpdate.cpp:499-700), the chi^2 gate
// UpdateBase::testChiSquared / whitenResidual (Update.cpp:36-56, :104-124).
//
// Maths. Per observation k at clone c_k with p_c = R^T (p_f - p):  A_k = Hproj(p_c) R^T (rho x 3),
// B_k = A_k [p_f]x.  Row block of H_x: [B_k | -A_k] on clone c_k and -B_k on the anchor's rotation
// columns (zero rotation block when c_k is the anchor); H_f rows = A_k; residual z - pi(p_c).
// The reference takes the left null space of H_f from a full-U JacobiSVD; any orthonormal basis gives
// the same gate value and the same posterior, so this kernel uses 3 Householder reflectors
// Q^T = (I - V T V^T)^T and evaluates, with the sparse rows of H_x,
//   H_proj[:, j] = (Q^T H_x[:, j])[3:],   r_proj = (Q^T r)[3:],
//   S = (Q^T (H_x P_s H_x^T) Q)[3:,3:] + sigma^2 I,   gamma = r_proj^T S^-1 r_proj.
// The forward substitution L y = r_proj rides along the Cholesky as one extra row of S.
// "SELECTED" mode reproduces the reference's overwrite of the anchor's six columns
// (KeyframeUpdate.cpp:523, SwMargUpdate.cpp:127): the anchor's own observation loses its -A_k block.
//
// Algorithmic bytes per feature: reads 24 + 8*rho*nobs (+ 96*SW poses and 8 n^2 of P_s per CTA),
// writes 8*(M-3)*(n+1).
#include <cstdint>
#include <cstdlib>

#include "igv_device.cuh"

using namespace igv;

namespace {

constexpr int kWarps = 8;
#ifndef IGV_STEREO_THREADS
#define IGV_STEREO_THREADS 384   // unfused stereo: up to 12 warps per CTA (168 registers)
#endif

struct FeatArgs {
  const double* P; int ld;
  const double* X; int xsize; IgvLayout L;
  int mode, F, obs_slots;
  const double* pf; const int* anchor; const double* obs; const unsigned char* mask; const int* dof;
  const unsigned char* feat_ok;
  double noise2;
  IgvDevParams prm;
  const double* chi2; int chi2_n;
  double* Hs; int qmax; int ldo;     // out block [f][qmax][ldo], ldo = 6*n_clones + 1
  size_t hs_seq_stride; int F_alloc; // per-sequence strides of Hs and of f_rows / f_gamma
  int* f_rows; double* f_gamma;
  int Mmax;                          // rho * n_clones
  int ldm;                           // (unused: the per-warp S matrix is packed, see SP below)
  int per_warp;                      // doubles of per-warp scratch
  int ssz;                           // doubles of the per-warp S region (>= (Mmax+1)(Mmax+2)/2; FUSE: also holds Z rows + D blocks)
  // FUSE: the track's contribution to the Gram matrix of the stack is accumulated in the kernel (see (6'))
  int nt, ldz;                       // 8-column tiles of the stack's n+1 columns; row stride of the Z rows
  double* G; long g_seq_stride; int n1p;   // out: upper triangle of [H r]^T [H r], row stride n1p
  int* n_acc; int max_valid;
  int const_sizes;                   // host: the launch matches the compile-time layout of the NCL > 0 instances
  int hs_f32;                        // IGV_PREC_FP32_STACK: the projected block is stored as float (same element strides)
};

// per-warp scratch layout (doubles): A[M][3] B[M][3] V[M][3] Am[M][3] E[M][3] r[M] qr[M] S (packed lower triangle of
// (M+1) rows: half the shared memory of full storage, which is what bounds the resident warps of stereo / wide windows) maps
__host__ __device__ inline int feat_ssz(int Mmax) { return (Mmax + 1) * (Mmax + 2) / 2; }
__host__ __device__ inline int feat_per_warp(int Mmax, int ssz) {
  return Mmax * 15 + 2 * Mmax + ssz + 2 * Mmax + 8;
}

// QT: largest projected block (rows) whose gate runs in registers, see (5a). FUSE: one CTA of up to 16 warps per
// sequence; instead of writing the projected blocks to HBM the kernel accumulates their Gram matrix, see (6').
// NCL > 0: the window size is a compile-time constant (the shipped / benchmarked windows), so every shared-memory array
// base and stride below folds into immediates instead of integer multiply-adds per access; NCL == 0: any window.
template <int RHO, bool PS_SMEM, int QT, bool FUSE, int NCL>
__global__ void __launch_bounds__(FUSE ? 512 : (RHO == 4 ? IGV_STEREO_THREADS : kWarps * 32), (FUSE || RHO == 4) ? 1 : 2) k_msckf_features(FeatArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int ncl = NCL > 0 ? NCL : a.L.n_clones, n = 6 * ncl, Mmax = NCL > 0 ? RHO * NCL : a.Mmax;
  // S is symmetric: only the lower triangle is stored, row r at r (r + 1) / 2; (r, c) with r >= c
  auto SP = [](int r, int c) { return ((r * (r + 1)) >> 1) + c; };
  const int c_nt = NCL > 0 ? (6 * NCL + 1 + 7) / 8 : a.nt, c_ldz = 8 * c_nt + 4;
  const int c_ssz = NCL > 0 ? (FUSE ? max(feat_ssz(Mmax), 3 * c_ldz + NCL * 27) : feat_ssz(Mmax)) : a.ssz;
  const int c_per_warp = NCL > 0 ? feat_per_warp(Mmax, c_ssz) : a.per_warp;
  double* sPose = sm;                                   // 12 per clone
  double* sPs = sm + 12 * IGV_MAX_CLONES;               // [n][n] clone block of P (symmetric), if staged
  double* ws = sPs + (PS_SMEM ? n * n : 0) + (size_t)warp * c_per_warp;
  double* sA = ws;                 // [M][3] rows of A_k (== H_f)
  double* sB = sA + Mmax * 3;      // [M][3] rows of B_k = A_k [pf]x
  double* sV = sB + Mmax * 3;      // [M][3] Householder vectors
  double* sAm = sV + Mmax * 3;     // [M][3] W T
  double* sE = sAm + Mmax * 3;     // [M][3] W T - V D
  double* sr = sE + Mmax * 3;      // [M] residual
  double* sqr = sr + Mmax;         // [M] Q^T r
  double* sS = sqr + Mmax;         // packed lower triangle of (M+1) rows: H_x P_s H_x^T, then S and its Cholesky factor (+ rhs row M)
  int* k2slot = reinterpret_cast<int*>(sS + c_ssz);             // [ncl] obs k -> slot
  int* slot2k = k2slot + ncl;                                    // [ncl] slot -> obs k or -1
  const double* Xb = a.X + (size_t)b * a.xsize;
  for (int t = threadIdx.x; t < 12 * ncl; t += blockDim.x) sPose[t] = Xb[IGV_X_CORE + t];
  const double* Pb = a.P + (size_t)b * a.ld * a.ld;
  if (PS_SMEM) {
    // getMarginalCov of the window (StateManager.cpp:128-153), once per CTA; i is the fast index so the
    // column-major global reads are contiguous, and P's symmetry makes the transposed store exact.
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t % n, j = t / n;
      sPs[t] = Pb[(a.L.idx_clone[i / 6] + i % 6) + (size_t)(a.L.idx_clone[j / 6] + j % 6) * a.ld];
    }
  }
  // FUSE: CTA-level accumulators behind the per-warp regions
  double* wsbase = sPs + (PS_SMEM ? n * n : 0);                  // per-warp regions start here
  const int zoff = Mmax * 15 + 2 * Mmax;                         // offset of sS inside a per-warp region
  const int ntt = c_nt * (c_nt + 1) / 2;
  double* S_acc = wsbase + (size_t)nwarps * c_per_warp;          // [(ncl+1) anchors][ncl][27]
  double2* Gt = reinterpret_cast<double2*>(                                                     // [ntt][32], 16-byte aligned
      (reinterpret_cast<uintptr_t>(S_acc + (size_t)(ncl + 1) * ncl * 27) + 15) & ~static_cast<uintptr_t>(15));
  double* zero_row = reinterpret_cast<double*>(Gt + (size_t)ntt * 32);   // [ldz] zeros (masked rows of the Z fold)
  int* flag_s = reinterpret_cast<int*>(zero_row + c_ldz);        // [16]
  int* tile_cc = flag_s + 16;                                    // [ntt] tile -> (column tile i) | (column tile j) << 8
  if (FUSE) {
    for (int t = threadIdx.x; t < (ncl + 1) * ncl * 27; t += blockDim.x) S_acc[t] = 0.0;
    for (int t = threadIdx.x; t < ntt * 32; t += blockDim.x) Gt[t] = make_double2(0.0, 0.0);
    for (int t = threadIdx.x; t < c_ldz; t += blockDim.x) zero_row[t] = 0.0;
    for (int t = threadIdx.x; t < ntt; t += blockDim.x) {
      int ci = 0, rem = t;
      while (rem >= c_nt - ci) { rem -= c_nt - ci; ++ci; }
      tile_cc[t] = ci | ((ci + rem) << 8);
    }
  }
  __syncthreads();
  const bool drop = (a.mode == IGV_VIS_SELECTED);
  int acc_count = 0;   // FUSE: accepted tracks so far (for the max_valid cap), identical in every thread

  for (int fbase = blockIdx.x * nwarps; fbase < a.F; fbase += gridDim.x * nwarps) {
    const int f = fbase + warp;
    int flag = 0;      // FUSE: 0 = no contribution, else 1 + anchor slot (ncl + 1: anchor outside the window)
    {
      // pull the NEXT track's inputs towards L1/L2 while this one is processed (they are read by dependent global
      // loads at the top of the body; no registers are held across the body)
      const int fn = f + gridDim.x * nwarps;
      if (fn < a.F) {
        const size_t bn = (size_t)b * a.F + fn;
        const char* po = reinterpret_cast<const char*>(a.obs + bn * a.obs_slots * RHO);
        const int obytes = a.obs_slots * RHO * 8;
        if (lane * 128 < obytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(po + lane * 128));
        if (lane == 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.mask + bn * a.obs_slots));
        if (lane == 9) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pf + bn * 3));
        if (lane == 10) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.anchor + bn));
        if (lane == 11) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.dof + bn));
      }
    }
    if (f < a.F) do {
    const size_t bf = (size_t)b * a.F + f;          // index into the caller's arrays
    const size_t bo = (size_t)b * a.F_alloc + f;    // index into f_rows / f_gamma
    if (a.feat_ok && !a.feat_ok[bf]) {              // track rejected upstream (e.g. triangulation failed)
      if (lane == 0) { a.f_rows[bo] = 0; a.f_gamma[bo] = nan(""); }
      break;
    }
    const double pf[3] = {a.pf[bf * 3], a.pf[bf * 3 + 1], a.pf[bf * 3 + 2]};
    const int anc = a.anchor[bf];
    const unsigned char* mk = a.mask + bf * a.obs_slots;
    const double* ob = a.obs + bf * a.obs_slots * RHO;
    // ---- (1) per-observation quantities; lanes over slots; NaN rows dropped like the reference ----
    int nobs = 0;
    for (int base = 0; base < ncl; base += 32) {
      const int s = base + lane;
      bool valid = (s < ncl) && mk[s];
      double Ak[3 * RHO], rk[RHO];
      if (valid) {
        const double* R = sPose + 12 * s;
        const double d[3] = {pf[0] - R[9], pf[1] - R[10], pf[2] - R[11]};
        double pc[3];
        mat3T_vec(R, d, pc);
        const double iz = 1.0 / pc[2];
        const double h02 = -pc[0] / (pc[2] * pc[2]), h12 = -pc[1] / (pc[2] * pc[2]);
        if (isnan(iz) || isnan(h02) || isnan(h12)) valid = false;  // H_proj.hasNaN()
        // A = Hproj * R^T : row0 = iz*R[:,0]^T + h02*R[:,2]^T ; (R^T)[i][j] = R[j][i]
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Ak[j] = iz * R[3 * j + 0] + h02 * R[3 * j + 2];
          Ak[3 + j] = iz * R[3 * j + 1] + h12 * R[3 * j + 2];
        }
        rk[0] = ob[s * RHO + 0] - pc[0] * iz;
        rk[1] = ob[s * RHO + 1] - pc[1] * iz;
        if (RHO == 4) {
          double pr[3];
          mat3_vec(a.prm.Rc, pc, pr);
#pragma unroll
          for (int i = 0; i < 3; ++i) pr[i] += a.prm.pc[i];
          const double izr = 1.0 / pr[2];
          const double g02 = -pr[0] / (pr[2] * pr[2]), g12 = -pr[1] / (pr[2] * pr[2]);
          double Hr[6];  // Hproj_r * Rc (2x3), then * R^T
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            Hr[j] = izr * a.prm.Rc[j] + g02 * a.prm.Rc[6 + j];
            Hr[3 + j] = izr * a.prm.Rc[3 + j] + g12 * a.prm.Rc[6 + j];
          }
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              Ak[(RHO - 2 + t) * 3 + j] = Hr[3 * t] * R[3 * j] + Hr[3 * t + 1] * R[3 * j + 1] + Hr[3 * t + 2] * R[3 * j + 2];
          rk[RHO - 2] = ob[s * RHO + RHO - 2] - pr[0] * izr;
          rk[RHO - 1] = ob[s * RHO + RHO - 1] - pr[1] * izr;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      const int k = nobs + __popc(bal & ((1u << lane) - 1u));
      if (s < ncl) slot2k[s] = valid ? k : -1;
      if (valid) {
        k2slot[k] = s;
#pragma unroll
        for (int t = 0; t < RHO; ++t) {
          const int row = k * RHO + t;
          const double* ar = Ak + 3 * t;
          // B = A * skew(pf)
          const double b0 = ar[1] * pf[2] - ar[2] * pf[1], b1 = ar[2] * pf[0] - ar[0] * pf[2], b2 = ar[0] * pf[1] - ar[1] * pf[0];
          sA[row * 3] = ar[0]; sA[row * 3 + 1] = ar[1]; sA[row * 3 + 2] = ar[2];
          sV[row * 3] = ar[0]; sV[row * 3 + 1] = ar[1]; sV[row * 3 + 2] = ar[2];
          sB[row * 3] = b0; sB[row * 3 + 1] = b1; sB[row * 3 + 2] = b2;
          sr[row] = rk[t];
        }
      }
      nobs += __popc(bal);
    }
    __syncwarp();
    const int M = nobs * RHO, q = M - 3;
    if (q < 1) {
      if (lane == 0) { a.f_rows[bo] = 0; a.f_gamma[bo] = nan(""); }
      __syncwarp();
      break;
    }
    const int kanc = (anc >= 0 && anc < ncl) ? slot2k[anc] : -1;  // observation taken at the anchor clone
    const int ca = (anc >= 0 && anc < ncl) ? anc : -1;
    // ---- (2) Householder QR of H_f (M x 3) in sV; T for the compact WY form ----------------------
    double tau[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double ss = 0.0;
      for (int i = j + 1 + lane; i < M; i += 32) ss = fma(sV[i * 3 + j], sV[i * 3 + j], ss);
      ss = warp_sum(ss);
      const double alpha = sV[j * 3 + j];
      double tj = 0.0, scale = 0.0;
      if (ss > 0.0) {
        const double beta = -copysign(sqrt(alpha * alpha + ss), alpha);
        tj = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      tau[j] = tj;
      __syncwarp();
      for (int i = j + 1 + lane; i < M; i += 32) sV[i * 3 + j] *= scale;
      if (lane == 0) sV[j * 3 + j] = 1.0;
      if (lane < j) sV[lane * 3 + j] = 0.0;
      __syncwarp();
#pragma unroll
      for (int c = j + 1; c < 3; ++c) {  // apply to the remaining columns of H_f
        double w = 0.0;
        for (int i = j + lane; i < M; i += 32) w = fma(sV[i * 3 + j], sV[i * 3 + c], w);
        w = warp_sum(w) * tj;
        __syncwarp();
        for (int i = j + lane; i < M; i += 32) sV[i * 3 + c] = fma(-w, sV[i * 3 + j], sV[i * 3 + c]);
        __syncwarp();
      }
    }
    // T (3x3 upper): T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j,0:j] V[:,0:j]^T v_j
    double g01 = 0.0, g02 = 0.0, g12 = 0.0;
    for (int i = lane; i < M; i += 32) {
      const double v0 = sV[i * 3], v1 = sV[i * 3 + 1], v2 = sV[i * 3 + 2];
      g01 = fma(v0, v1, g01); g02 = fma(v0, v2, g02); g12 = fma(v1, v2, g12);
    }
    g01 = warp_sum(g01); g02 = warp_sum(g02); g12 = warp_sum(g12);
    const double T00 = tau[0], T11 = tau[1], T22 = tau[2];
    const double T01 = -tau[1] * T00 * g01;
    const double T02 = -tau[2] * (T00 * g02 + T01 * g12);
    const double T12 = -tau[2] * (T11 * g12);
    // Q^T x = x - V T^T (V^T x):  z = T^T y, z0 = T00 y0, z1 = T01 y0 + T11 y1, z2 = T02 y0 + T12 y1 + T22 y2
    // ---- (3) M0 = H_x P_s H_x^T into sS (full storage) ------------------------------------------------
    // Row block of observation k: U_k (rho x 6) on its clone c_k and W_k = -B_k (rho x 3) on the anchor's
    // rotation columns, so  M0[k1,k2] = U1 P[c1,c2] U2^T + (U1 P[c1,a] + W1 P[a,a]) W2^T + W1 (P[a,c2] U2^T).
    // The two anchor factors depend on ONE observation each: (3a) computes ga_k = U_k P[c_k,a] + W_k P[a,a]
    // (rho x 3, into sAm) and pa_k = P[a,c_k] U_k^T (3 x rho, into sE) with one lane per observation, (3b) then
    // needs a single 6 x 6 block of P_s per observation pair.
    const int rs = PS_SMEM ? n : 1;               // strides of the P blocks: the SMEM copy is [n][n]; the global
    const long cs = PS_SMEM ? 1 : a.ld;           // matrix is column-major (symmetric)
    auto pblk = [&](int cr, int cc) -> const double* {   // block (clone cr, clone cc), element (i,j) at [i*rs + j*cs]
      return PS_SMEM ? sPs + (6 * cr) * n + 6 * cc : Pb + a.L.idx_clone[cr] + (long)a.L.idx_clone[cc] * a.ld;
    };
    for (int row = lane; row < M; row += 32) {   // one lane per ROW of the stack (observation row / RHO)
      const int k = row / RHO;
      const int c = k2slot[k];
      const bool ak = (k == kanc);
      double ga[3] = {0.0, 0.0, 0.0}, pa[3] = {0.0, 0.0, 0.0};
      if (ca >= 0) {
        double u[6], w[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double bb = sB[row * 3 + j], aa = sA[row * 3 + j];
          u[j] = ak ? 0.0 : bb;
          u[3 + j] = (ak && drop) ? 0.0 : -aa;
          w[j] = ak ? 0.0 : -bb;
        }
        const double* p1a = pblk(c, ca);
        const double* pa2 = pblk(ca, c);
        const double* paa = pblk(ca, ca);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            ga[j] = fma(u[i], p1a[i * rs + j * cs], ga[j]);
            pa[j] = fma(pa2[j * rs + i * cs], u[i], pa[j]);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) ga[j] = fma(w[i], paa[i * rs + j * cs], ga[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) { sAm[row * 3 + j] = ga[j]; sE[row * 3 + j] = pa[j]; }
    }
    __syncwarp();
    {
      const int npair = nobs * (nobs + 1) / 2;
      for (int pr = lane; pr < npair; pr += 32) {
        int k1 = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
        while (k1 * (k1 + 1) / 2 > pr) --k1;
        while ((k1 + 1) * (k1 + 2) / 2 <= pr) ++k1;
        const int k2 = pr - k1 * (k1 + 1) / 2;  // k1 >= k2
        const double* p12 = pblk(k2slot[k1], k2slot[k2]);
        const bool a1 = (k1 == kanc), a2 = (k2 == kanc);
        // rows of observation k2: u2 (6, own clone), w2 (3, anchor rotation), pa2 (3: column of P[a,c2] U2^T)
        double u2[RHO][6], w2[RHO][3], pa2[RHO][3];
#pragma unroll
        for (int t2 = 0; t2 < RHO; ++t2) {
          const int r2 = k2 * RHO + t2;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double bb = sB[r2 * 3 + j], aa = sA[r2 * 3 + j];
            u2[t2][j] = a2 ? 0.0 : bb;
            u2[t2][3 + j] = (a2 && drop) ? 0.0 : -aa;
            w2[t2][j] = (a2 || ca < 0) ? 0.0 : -bb;
            pa2[t2][j] = sE[r2 * 3 + j];
          }
        }
#pragma unroll
        for (int t1 = 0; t1 < RHO; t1 += 2) {
          const int r1 = k1 * RHO + t1;
          double u[2][6], w[2][3], ga[2][3];
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double bb = sB[(r1 + t) * 3 + j], aa = sA[(r1 + t) * 3 + j];
              u[t][j] = a1 ? 0.0 : bb;
              u[t][3 + j] = (a1 && drop) ? 0.0 : -aa;
              w[t][j] = (a1 || ca < 0) ? 0.0 : -bb;
              ga[t][j] = sAm[(r1 + t) * 3 + j];
            }
          double g[2][6];   // rows r1, r1+1 of U1 P[c1,c2]
#pragma unroll
          for (int j = 0; j < 6; ++j) g[0][j] = g[1][j] = 0.0;
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const double p = p12[i * rs + j * cs];
              g[0][j] = fma(u[0][i], p, g[0][j]);
              g[1][j] = fma(u[1][i], p, g[1][j]);
            }
#pragma unroll
          for (int t2 = 0; t2 < RHO; ++t2) {
            const int r2 = k2 * RHO + t2;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) { acc0 = fma(g[0][j], u2[t2][j], acc0); acc1 = fma(g[1][j], u2[t2][j], acc1); }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              acc0 = fma(w[0][j], pa2[t2][j], acc0); acc0 = fma(ga[0][j], w2[t2][j], acc0);
              acc1 = fma(w[1][j], pa2[t2][j], acc1); acc1 = fma(ga[1][j], w2[t2][j], acc1);
            }
            if (r2 <= r1) sS[SP(r1, r2)] = acc0;
            if (r2 <= r1 + 1) sS[SP(r1 + 1, r2)] = acc1;
          }
        }
      }
    }
    __syncwarp();
    // ---- (4) two-sided projection  S' = Q^T M0 Q  (only rows/cols >= 3 are needed afterwards) -----
    // With W = M0 V (M x 3), C = V^T W, D = T^T C T, Am = W T, E = Am - V D:
    //   (Q^T M0 Q)[i][j] = M0[i][j] - sum_p ( V[i][p] Am[j][p] + E[i][p] V[j][p] )
    {
      double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = lane; i < M; i += 32) {
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        for (int k = 0; k < M; ++k) {
          const double m = sS[k <= i ? SP(i, k) : SP(k, i)];   // row i of the symmetric matrix (column part: consecutive lanes, consecutive words)
          w0 = fma(m, sV[k * 3 + 0], w0);
          w1 = fma(m, sV[k * 3 + 1], w1);
          w2 = fma(m, sV[k * 3 + 2], w2);
        }
        sAm[i * 3] = w0; sAm[i * 3 + 1] = w1; sAm[i * 3 + 2] = w2;   // W for now
        const double v[3] = {sV[i * 3], sV[i * 3 + 1], sV[i * 3 + 2]};
        const double ww[3] = {w0, w1, w2};
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int qq = 0; qq < 3; ++qq) C[p * 3 + qq] = fma(v[p], ww[qq], C[p * 3 + qq]);
      }
#pragma unroll
      for (int e = 0; e < 9; ++e) C[e] = warp_sum(C[e]);
      const double Tm[9] = {T00, T01, T02, 0.0, T11, T12, 0.0, 0.0, T22};
      double Et[9], D[9];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {  // Et = T^T C
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) acc += Tm[k * 3 + p] * C[k * 3 + qq];
          Et[p * 3 + qq] = acc;
        }
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {  // D = Et T
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) acc += Et[p * 3 + k] * Tm[k * 3 + qq];
          D[p * 3 + qq] = acc;
        }
      __syncwarp();
      for (int i = lane; i < M; i += 32) {
        const double w[3] = {sAm[i * 3], sAm[i * 3 + 1], sAm[i * 3 + 2]};
        const double v[3] = {sV[i * 3], sV[i * 3 + 1], sV[i * 3 + 2]};
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
          const double am = w[0] * Tm[qq] + w[1] * Tm[3 + qq] + w[2] * Tm[6 + qq];
          const double vd = v[0] * D[qq] + v[1] * D[3 + qq] + v[2] * D[6 + qq];
          sAm[i * 3 + qq] = am;
          sE[i * 3 + qq] = am - vd;
        }
      }
      // r_proj = (Q^T r)[3:]
      double y0 = 0.0, y1 = 0.0, y2 = 0.0;
      for (int i = lane; i < M; i += 32) {
        const double ri = sr[i];
        y0 = fma(sV[i * 3 + 0], ri, y0); y1 = fma(sV[i * 3 + 1], ri, y1); y2 = fma(sV[i * 3 + 2], ri, y2);
      }
      y0 = warp_sum(y0); y1 = warp_sum(y1); y2 = warp_sum(y2);
      const double z0 = T00 * y0, z1 = T01 * y0 + T11 * y1, z2 = T02 * y0 + T12 * y1 + T22 * y2;
      __syncwarp();
      for (int i = lane; i < M; i += 32) {
        const double qv = sr[i] - (sV[i * 3 + 0] * z0 + sV[i * 3 + 1] * z1 + sV[i * 3 + 2] * z2);
        sqr[i] = qv;
        if (i >= 3) sS[SP(M, i)] = qv;          // extra row M of S: the right-hand side of L y = r_proj
      }
      __syncwarp();
    }
    bool pd = true;
    double gamma = nan("");
    if (q <= QT && q < 32) {
      // ---- (5a) small blocks: S' stays in REGISTERS, lane = row (lane q holds the right-hand side), and
      // the gate value comes from a square-root-free LDL^T elimination: with S = L D L^T and L z = r_proj,
      // gamma = sum_j z_j^2 / d_j. Positive definiteness (all d_j > 0) is the same test the Cholesky makes.
      // Column j of the trailing update needs a_cj from lane c: one shuffle per (j, c) pair instead of the
      // shared-memory round trips and barriers of the in-place version below (30% of this kernel's time).
      const int row = lane;
      const bool on = row < q;
      const int ii = on ? 3 + row : 3;
      const double v0 = sV[ii * 3], v1 = sV[ii * 3 + 1], v2 = sV[ii * 3 + 2];
      const double e0 = sE[ii * 3], e1 = sE[ii * 3 + 1], e2 = sE[ii * 3 + 2];
      double ar[QT];
#pragma unroll
      for (int c = 0; c < QT; ++c) {
        const int j = min(c + 3, M - 1);
        // lower triangle only (c <= row): the elimination below never reads a lane's entries right of its diagonal
        double val = sS[SP(ii, min(j, ii))] - (v0 * sAm[j * 3] + v1 * sAm[j * 3 + 1] + v2 * sAm[j * 3 + 2] +
                                               e0 * sV[j * 3] + e1 * sV[j * 3 + 1] + e2 * sV[j * 3 + 2]);
        if (c == row) val += a.noise2;
        if (row == q) val = sqr[j];
        ar[c] = (c < q && (c <= row || row == q)) ? val : 0.0;
      }
      double gacc = 0.0;
#pragma unroll
      for (int j = 0; j < QT; ++j) {
        if (j >= q) break;
        const double d = __shfl_sync(0xffffffffu, ar[j], j);
        if (!(d > 0.0)) { pd = false; break; }
        const double inv = rcp_nobranch(d);
        const double t = ar[j];           // a_{row, j}
        const double l = t * inv;
        if (row == q) gacc = fma(t, l, gacc);
#pragma unroll
        for (int c = j + 1; c < QT; ++c) {
          const double tc = __shfl_sync(0xffffffffu, t, c);   // a_{c, j}
          ar[c] = fma(-l, tc, ar[c]);
        }
      }
      if (pd) gamma = __shfl_sync(0xffffffffu, gacc, q);
    } else {
      // trailing block, lower triangle: lane = row i, uniform loop over j (broadcast reads of row j data)
      for (int i0 = 3; i0 < M; i0 += 32) {
        const int i = i0 + lane;
        const bool on = i < M;
        const int ii = on ? i : 3;
        const double v0 = sV[ii * 3], v1 = sV[ii * 3 + 1], v2 = sV[ii * 3 + 2];
        const double e0 = sE[ii * 3], e1 = sE[ii * 3 + 1], e2 = sE[ii * 3 + 2];
        const int jmax = min(M, i0 + 32);
        for (int j = 3; j < jmax; ++j) {
          const int at = SP(ii, min(j, ii));
          const double val = sS[at] - (v0 * sAm[j * 3] + v1 * sAm[j * 3 + 1] + v2 * sAm[j * 3 + 2] +
                                       e0 * sV[j * 3] + e1 * sV[j * 3 + 1] + e2 * sV[j * 3 + 2]);
          if (on && j <= i) sS[at] = (j == i) ? val + a.noise2 : val;
        }
      }
      __syncwarp();
      // ---- (5b) Cholesky of S = sS[3:M,3:M] (lower) with the rhs as row M: L y = r_proj rides along ----
      // Blocked right-looking factorisation inside ONE warp, 8 columns at a time (the 41-row stereo and the wide-window
      // gates live here): (a) the 8 x 8 diagonal block is factored in registers, lane = row, columns exchanged by
      // shuffles; (b) the panel rows below it (the right-hand side is the last one) are solved against it, one row per
      // lane; (c) the trailing block is updated with DMMA.8x8x4 tiles. A column-at-a-time left-looking loop took half of
      // this kernel's time for stereo (41 dependent steps of dot products through shared memory).
      {
        auto A = [&](int i, int j) -> double& { return sS[SP(3 + i, 3 + j)]; };   // S'(i, j), i >= j; row q: right-hand side
        const int fk = lane & 3, fc = lane >> 2;
        for (int jb = 0; jb < q; jb += 8) {
          const int nb = min(8, q - jb);
          double arow[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) arow[c] = (lane < nb && c <= lane) ? A(jb + lane, jb + c) : 0.0;
          double rd = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k >= nb) break;
            const double d = __shfl_sync(0xffffffffu, arow[k], k);
            if (!(d > 0.0)) { pd = false; break; }
            const double inv = rsqrt_nobranch(d);   // d > 0 checked above
            const double l = arow[k] * inv;          // L[lane][k] (lane == k: sqrt(d))
            arow[k] = l;
            if (lane == k) rd = inv;
#pragma unroll
            for (int c = k + 1; c < 8; ++c) {
              const double lc = __shfl_sync(0xffffffffu, l, c);
              if (c <= lane) arow[c] = fma(-l, lc, arow[c]);
            }
          }
          if (!pd) break;
          if (lane < nb) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c <= lane && c < nb) A(jb + lane, jb + c) = arow[c];
          }
          double rdg[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) rdg[k] = __shfl_sync(0xffffffffu, rd, k);
          __syncwarp();
          const int i0 = jb + nb;
          for (int i = i0 + lane; i <= q; i += 32) {   // (b) panel: L21[i,:] = A21[i,:] L11^-T
            double x[8];
            double* row = &A(i, jb);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < nb) {
                double acc = row[k];
#pragma unroll
                for (int p2 = 0; p2 < 8; ++p2) if (p2 < k) acc = fma(-x[p2], A(jb + k, jb + p2), acc);
                x[k] = acc * rdg[k];
              }
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < nb) row[k] = x[k];
          }
          __syncwarp();
          if (i0 >= q) break;   // no columns left: the right-hand side row is solved
          const int nrt = (q + 1 - i0 + 7) / 8;   // (c) trailing update, 8 x 8 tiles of the lower triangle (+ the rhs row)
          for (int it = 0; it < nrt; ++it)
            for (int jt = 0; jt <= it; ++jt) {
              const int ri = i0 + 8 * it + fc, cj = i0 + 8 * jt + fc, c0 = i0 + 8 * jt + 2 * fk;
              double cx = 0.0, cy = 0.0;
              for (int k0 = 0; k0 < nb; k0 += 4) {
                const bool kv = k0 + fk < nb;
                const double av = (kv && ri <= q) ? A(ri, jb + k0 + fk) : 0.0;
                const double bv = (kv && cj < q) ? A(cj, jb + k0 + fk) : 0.0;
                mma884(cx, cy, av, bv);
              }
              if (ri <= q) {
                if (c0 < q && c0 <= ri) A(ri, c0) -= cx;
                if (c0 + 1 < q && c0 + 1 <= ri) A(ri, c0 + 1) -= cy;
              }
            }
          __syncwarp();
        }
      }
      if (pd) {
        double g = 0.0;
        for (int c = 3 + lane; c < M; c += 32) { const double y = sS[SP(M, c)]; g = fma(y, y, g); }
        gamma = warp_sum(g);
      }
    }
    const int dof = a.dof[bf];
    const bool accept = pd && dof >= 1 && dof <= a.chi2_n && (gamma < a.chi2[dof - 1]);
    if (lane == 0) { a.f_gamma[bo] = gamma; a.f_rows[bo] = accept ? q : 0; }
    // ---- (6') FUSE: the track's term of G = [H r]^T [H r] without forming the projected block ---------------
    // With W = [H_x | r] (M rows) and Q = [Q1 | Q2] the projected block is Q2^T W, so its Gram matrix is
    //     W^T Q2 Q2^T W = W^T W - Z^T Z,   Z = Q1^T W = rows 0..2 of Q^T W   (3 x (n+1)).
    // W^T W is block sparse: with U_s = [B | -A] the rows of the observation at clone s (6 columns) and the
    // anchor a's rotation columns carrying -B,  D_s = U_s^T U_s, d_s = U_s^T r_s:
    //     G[s,s] += D_s,  G[s,a_rot] -= D_s[:, :3],  G[a_rot,a_rot] += D_s[:3,:3],  g[s] += d_s,  g[a_rot] -= d_s[:3].
    // The warp leaves Z (3 rows) and the D_s | d_s (27 numbers per clone) in its own S region; after the block
    // barrier the CTA folds the Z rows of all warps into G with DMMA and the D_s into S_acc[anchor][s] in warp
    // order, so the sum is independent of the schedule (bitwise reproducible).
    if (accept && FUSE) {
      double* zrow = sS;                      // [3][ldz]
      double* dblk = sS + 3 * c_ldz;          // [ncl][27]
      const int ncolp = 8 * c_nt;
      // V^T (rows of the anchor's rotation columns): lanes 0..8 hold entry (q = lane / 3, comp = lane % 3) of
      // -sum_{rows not at the anchor} V[row][q] B[row][comp]
      double yanc = 0.0;
      if (ca >= 0 && lane < 9) {
        const int q_ = lane / 3, cp = lane - 3 * q_;
        for (int row = 0; row < M; ++row)
          if (row / RHO != kanc) yanc = fma(sV[row * 3 + q_], -sB[row * 3 + cp], yanc);
      }
      for (int j0 = 0; j0 < ncolp; j0 += 32) {
        const int j = j0 + lane;
        const int c = (j < n) ? j / 6 : 0, comp = (j < n) ? j - 6 * c : 0;
        const int cq = (comp < 3) ? comp : 0;
        const double ya0 = __shfl_sync(0xffffffffu, yanc, cq), ya1 = __shfl_sync(0xffffffffu, yanc, 3 + cq),
                     ya2 = __shfl_sync(0xffffffffu, yanc, 6 + cq);
        double zr[3] = {0.0, 0.0, 0.0};
        if (j < n) {
          const int kown = slot2k[c];
          const bool anc_rot = (c == anc) && (comp < 3);
          double y0 = 0.0, y1 = 0.0, y2 = 0.0, ownv[RHO];
#pragma unroll
          for (int t = 0; t < RHO; ++t) {
            double v = 0.0;
            if (kown >= 0) {
              const int row = kown * RHO + t;
              if (comp < 3) v = (kown == kanc) ? 0.0 : sB[row * 3 + comp];
              else v = ((kown == kanc) && drop) ? 0.0 : -sA[row * 3 + comp - 3];
              y0 = fma(sV[row * 3], v, y0); y1 = fma(sV[row * 3 + 1], v, y1); y2 = fma(sV[row * 3 + 2], v, y2);
            }
            ownv[t] = v;
          }
          if (anc_rot) { y0 += ya0; y1 += ya1; y2 += ya2; }
          const double z0 = T00 * y0, z1 = T01 * y0 + T11 * y1, z2 = T02 * y0 + T12 * y1 + T22 * y2;
          const int own_lo = (kown >= 0) ? kown * RHO : -(1 << 20);
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            double v = -(sV[p * 3] * z0 + sV[p * 3 + 1] * z1 + sV[p * 3 + 2] * z2);
            const int t = p - own_lo;
#pragma unroll
            for (int tt = 0; tt < RHO; ++tt) if (tt == t) v += ownv[tt];
            if (anc_rot && (p / RHO != kanc)) v -= sB[p * 3 + comp];
            zr[p] = v;
          }
        } else if (j == n) {
          zr[0] = sqr[0]; zr[1] = sqr[1]; zr[2] = sqr[2];
        }
        if (j < ncolp) { zrow[j] = zr[0]; zrow[c_ldz + j] = zr[1]; zrow[2 * c_ldz + j] = zr[2]; }
      }
      for (int sl = lane; sl < ncl; sl += 32) {
        double* dd = dblk + sl * 27;
        const int k = slot2k[sl];
        if (k < 0) {
#pragma unroll
          for (int e = 0; e < 27; ++e) dd[e] = 0.0;
        } else {
          const bool isanc = (k == kanc);
          double u[RHO][6], rr[RHO];
#pragma unroll
          for (int t = 0; t < RHO; ++t) {
            const int row = k * RHO + t;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              u[t][j] = isanc ? 0.0 : sB[row * 3 + j];
              u[t][3 + j] = (isanc && drop) ? 0.0 : -sA[row * 3 + j];
            }
            rr[t] = sr[row];
          }
          int e = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j, ++e) {
              double acc = 0.0;
#pragma unroll
              for (int t = 0; t < RHO; ++t) acc = fma(u[t][i], u[t][j], acc);
              dd[e] = acc;
            }
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < RHO; ++t) acc = fma(u[t][i], rr[t], acc);
            dd[21 + i] = acc;
          }
        }
      }
      flag = (ca >= 0) ? ca + 1 : ncl + 1;
    }
    // ---- (6) write the projected block [H | r] rows 3..M-1, row-major, lanes over columns ----------
    if (accept && !FUSE) {
      double* out = a.Hs + (size_t)b * a.hs_seq_stride + (size_t)f * a.qmax * a.ldo;
      // three column groups (j, j + 32, j + 64) per pass so that the broadcast reads of V's row i and the
      // row loop are shared; per group the lane keeps z = T^T V^T a_j and its column's own entries
      for (int j0 = 0; j0 < n + 1; j0 += 96) {
        double z[3][3], ownv[3][RHO];
        int own_lo[3], acomp[3];   // acomp >= 0: anchor rotation column (component), else -1
        bool isres[3], live[3];
#pragma unroll
        for (int gq = 0; gq < 3; ++gq) {
          const int j = j0 + 32 * gq + lane;
          live[gq] = j < n + 1;
          isres[gq] = (j == n);
          const int jc = (j < n) ? j : 0;
          const int c = jc / 6, comp = jc % 6;
          const int kown = (j < n) ? slot2k[c] : -1;
          const bool anc_rot = (j < n) && (c == anc) && (comp < 3);
          acomp[gq] = anc_rot ? comp : -1;
          double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll
          for (int t = 0; t < RHO; ++t) {
            double v = 0.0;
            if (kown >= 0) {
              const int row = kown * RHO + t;
              if (comp < 3) v = (kown == kanc) ? 0.0 : sB[row * 3 + comp];
              else v = ((kown == kanc) && drop) ? 0.0 : -sA[row * 3 + comp - 3];
              y0 = fma(sV[row * 3], v, y0); y1 = fma(sV[row * 3 + 1], v, y1); y2 = fma(sV[row * 3 + 2], v, y2);
            }
            ownv[gq][t] = v;
          }
          if (anc_rot) {
            for (int row = 0; row < M; ++row) {
              if (row / RHO == kanc) continue;
              const double v = -sB[row * 3 + comp];
              y0 = fma(sV[row * 3], v, y0); y1 = fma(sV[row * 3 + 1], v, y1); y2 = fma(sV[row * 3 + 2], v, y2);
            }
          }
          z[gq][0] = T00 * y0;
          z[gq][1] = T01 * y0 + T11 * y1;
          z[gq][2] = T02 * y0 + T12 * y1 + T22 * y2;
          own_lo[gq] = (kown >= 0) ? kown * RHO : -(1 << 20);
        }
        for (int i = 3; i < M; ++i) {
          const double s0 = sV[i * 3], s1 = sV[i * 3 + 1], s2 = sV[i * 3 + 2];
          const double qri = sqr[i];
          const bool not_anchor_row = (i / RHO != kanc);
          double* orow = out + (size_t)(i - 3) * a.ldo + j0 + lane;
#pragma unroll
          for (int gq = 0; gq < 3; ++gq) {
            // (the column's own entries, which live in RHO of the M rows, are added by the short pass below: a select chain
            // over RHO candidates in this loop was 12 % of the stereo kernel's instructions)
            double v = -(s0 * z[gq][0] + s1 * z[gq][1] + s2 * z[gq][2]);
            if (acomp[gq] >= 0 && not_anchor_row) v -= sB[i * 3 + acomp[gq]];
            if (isres[gq]) v = qri;
            if (live[gq]) {
              if (a.hs_f32)   // float elements with the SAME element strides, counted from the sequence's base
                reinterpret_cast<float*>(a.Hs + (size_t)b * a.hs_seq_stride)[((size_t)f * a.qmax + (i - 3)) * a.ldo + j0 + lane + 32 * gq] = (float)v;
              else orow[32 * gq] = v;
            }
          }
        }
        // the RHO rows of the observation at the column's own clone: the same value plus the column's own entry (same order
        // of operations as if it had been added in the loop above)
#pragma unroll
        for (int gq = 0; gq < 3; ++gq) {
          if (!live[gq] || isres[gq]) continue;
#pragma unroll
          for (int tt = 0; tt < RHO; ++tt) {
            const int i = own_lo[gq] + tt;
            if (i >= 3 && i < M) {
              double v = -(sV[i * 3] * z[gq][0] + sV[i * 3 + 1] * z[gq][1] + sV[i * 3 + 2] * z[gq][2]);
              v += ownv[gq][tt];
              if (acomp[gq] >= 0 && (i / RHO != kanc)) v -= sB[i * 3 + acomp[gq]];
              const size_t at = ((size_t)f * a.qmax + (i - 3)) * a.ldo + j0 + lane + 32 * gq;
              if (a.hs_f32) reinterpret_cast<float*>(a.Hs + (size_t)b * a.hs_seq_stride)[at] = (float)v;
              else (a.Hs + (size_t)b * a.hs_seq_stride)[at] = v;
            }
          }
        }
      }
    }
    __syncwarp();
    } while (0);
    if (FUSE) {
      const int fk = lane & 3, fc = lane >> 2;
      if (lane == 0) flag_s[warp] = flag;
      __syncthreads();
      // warps whose track enters the update: accepted by the gate and inside the max_valid cap (in track order)
      const unsigned accm = __ballot_sync(0xffffffffu, lane < nwarps && flag_s[lane < nwarps ? lane : 0] != 0);
      unsigned onmask = accm;
      if (a.max_valid > 0 && a.max_valid - acc_count < __popc(accm)) {   // the cap cuts into this round
        onmask = 0;
        unsigned mm = accm;
        for (int left = a.max_valid - acc_count; mm && left > 0; --left) { onmask |= mm & (0u - mm); mm &= mm - 1; }
      }
      acc_count += __popc(accm);
      if (onmask) {
        // Gt += Z^T Z over the 3 * nwarps staged rows (four per DMMA k-step; rows of warps that do not contribute
        // are masked to zero), subtracted from G at the end. The upper-triangle tiles are dealt to the warps, at
        // most four each (nwarps >= 12, <= 45 tiles), and stay in registers across the k-steps.
        double2 g[4];
        int ci8[4], cj8[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ti = warp + u * nwarps;
          if (ti < ntt) {
            const int cc = tile_cc[ti];
            ci8[u] = 8 * (cc & 0xff) + fc; cj8[u] = 8 * (cc >> 8) + fc;
            g[u] = Gt[ti * 32 + lane];
          }
        }
        const int nrows = 3 * nwarps;
        for (int k0 = 0; k0 < nrows; k0 += 4) {
          const int r = k0 + fk, w = (r * 43) >> 7, p = r - 3 * w;   // w = r / 3 for r < 64
          const bool on = r < nrows && ((onmask >> w) & 1u);
          if (!__any_sync(0xffffffffu, on)) continue;
          // masked rows read a row of zeros instead of being selected away after the load
          const double* zp = on ? wsbase + (size_t)w * c_per_warp + zoff + p * c_ldz : zero_row;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (warp + u * nwarps < ntt) mma884(g[u].x, g[u].y, zp[ci8[u]], zp[cj8[u]]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (warp + u * nwarps < ntt) Gt[(warp + u * nwarps) * 32 + lane] = g[u];
        for (int idx = threadIdx.x; idx < ncl * 27; idx += blockDim.x)
          for (unsigned mm = onmask; mm; mm &= mm - 1) {
            const int w = __ffs(mm) - 1;
            S_acc[(size_t)(flag_s[w] - 1) * ncl * 27 + idx] += (wsbase + (size_t)w * c_per_warp + zoff + 3 * c_ldz)[idx];
          }
      }
      __syncthreads();
    }
  }
  if (FUSE) {
    // ---- assemble the upper triangle of G in shared memory (over the dead per-warp regions) and write it -------
    const int fk = lane & 3, fc = lane >> 2;
    const int n1 = n + 1, ldg = n1 | 1;
    double* Gf = wsbase;
    for (int ti = warp; ti < ntt; ti += nwarps) {
      const int ci = tile_cc[ti] & 0xff, cj = tile_cc[ti] >> 8;
      const double2 g = Gt[ti * 32 + lane];
      const int row = 8 * ci + fc, col = 8 * cj + 2 * fk;
      if (row < n1) {   // Gt holds + sum Z^T Z
        if (col < n1) Gf[row * ldg + col] = -g.x;
        if (col + 1 < n1) Gf[row * ldg + col + 1] = -g.y;
      }
    }
    __syncthreads();
    auto uidx = [](int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); };   // i <= j < 6
    // clone-diagonal blocks and the right-hand side: sum over the anchors in a fixed order
    for (int idx = threadIdx.x; idx < ncl * 27; idx += blockDim.x) {
      const int sl = idx / 27, e = idx - 27 * sl;
      double tot = 0.0;
      for (int an = 0; an <= ncl; ++an) tot += S_acc[(size_t)an * ncl * 27 + idx];
      if (e < 21) {
        int i = 0, rem = e;
        while (rem >= 6 - i) { rem -= 6 - i; ++i; }
        Gf[(6 * sl + i) * ldg + 6 * sl + i + rem] += tot;
      } else {
        Gf[(6 * sl + e - 21) * ldg + n] += tot;
      }
    }
    __syncthreads();
    // anchor coupling: block (s, a_rot) -= D[:, :3]. Blocks {s, a} and {a, s} share the rot-rot entries of the
    // upper triangle, so s < a and s > a are two phases; the anchor's own rot-rot block and right-hand side ride
    // in the first one (disjoint targets).
    for (int phase = 0; phase < 2; ++phase) {
      for (int idx = threadIdx.x; idx < ncl * ncl * 18; idx += blockDim.x) {
        const int an = idx / (ncl * 18), rem = idx - an * ncl * 18, sl = rem / 18, q18 = rem - 18 * sl;
        const int i = q18 / 3, jp = q18 - 3 * i;
        if (sl == an || ((sl < an) != (phase == 0))) continue;
        const double val = S_acc[((size_t)an * ncl + sl) * 27 + uidx(min(i, jp), max(i, jp))];
        const int r_ = 6 * sl + i, c_ = 6 * an + jp;
        Gf[min(r_, c_) * ldg + max(r_, c_)] -= val;
      }
      if (phase == 0) {
        for (int idx = threadIdx.x; idx < ncl * 9; idx += blockDim.x) {
          const int an = idx / 9, e9 = idx - 9 * an;
          int i = 0, j = 0, src;
          if (e9 < 6) { int rem = e9; while (rem >= 3 - i) { rem -= 3 - i; ++i; } j = i + rem; src = uidx(i, j); }
          else { i = e9 - 6; src = 21 + i; }
          double tot = 0.0;
          for (int sl = 0; sl < ncl; ++sl)
            if (sl != an) tot += S_acc[((size_t)an * ncl + sl) * 27 + src];
          if (e9 < 6) Gf[(6 * an + i) * ldg + 6 * an + j] += tot;
          else Gf[(6 * an + i) * ldg + n] -= tot;
        }
      }
      __syncthreads();
    }
    double* G = a.G + (size_t)b * a.g_seq_stride;
    for (int row = warp; row < n1; row += nwarps)
      for (int col = row + lane; col < n1; col += 32) G[(size_t)row * a.n1p + col] = Gf[row * ldg + col];
    if (threadIdx.x == 0 && a.n_acc) a.n_acc[b] = (a.max_valid > 0) ? min(acc_count, a.max_valid) : acc_count;
  }
}

template <int RHO, bool PS, int QT, bool FUSE, int NCL>
void launch_feat_q(const FeatArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  IGV_SMEM_OPTIN((k_msckf_features<RHO, PS, QT, FUSE, NCL>), 224 * 1024);
  k_msckf_features<RHO, PS, QT, FUSE, NCL><<<grid, threads, smem, st>>>(a);
}

template <int RHO, bool PS, bool FUSE>
void launch_feat(const FeatArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  // register-resident gate sized to the largest block this window can produce (rho * clones - 3 rows)
  const int qmax = a.Mmax - 3;
  // the benchmarked window (SW = 11, BASELINE configs c2 / c3 / c4) gets an instance with compile-time sizes
  if (PS && a.L.n_clones == 11 && a.const_sizes) {
    if (RHO == 2) launch_feat_q<RHO, PS, 21, FUSE, 11>(a, grid, threads, smem, st);
    else launch_feat_q<RHO, PS, 25, FUSE, 11>(a, grid, threads, smem, st);
    return;
  }
  if (qmax <= 9) launch_feat_q<RHO, PS, 9, FUSE, 0>(a, grid, threads, smem, st);
  else if (qmax <= 17) launch_feat_q<RHO, PS, 17, FUSE, 0>(a, grid, threads, smem, st);
  else if (qmax <= 21) launch_feat_q<RHO, PS, 21, FUSE, 0>(a, grid, threads, smem, st);
  else launch_feat_q<RHO, PS, 25, FUSE, 0>(a, grid, threads, smem, st);
}

}  // namespace

// Sets h->feat_fused when the kernel also accumulated the Gram matrix of the stack (FUSE): igv_launch_qr_compress
// then only runs the factorisation.
void igv_launch_msckf_features(igv_batch* h, const IgvMsckfLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_FEATURES);
  FeatArgs a;
  a.P = h->Pc(); a.ld = h->ld;
  a.X = h->Xc(); a.xsize = h->xsize; a.L = h->layout();
  a.mode = l.mode; a.F = l.F; a.obs_slots = l.obs_slots;
  a.pf = l.pf; a.anchor = l.anchor; a.obs = l.obs; a.mask = l.mask; a.dof = l.dof; a.feat_ok = l.feat_ok;
  a.noise2 = l.noise * l.noise;
  a.prm = h->params;
  a.chi2 = h->chi2; a.chi2_n = h->chi2_n;
  a.Hs = h->Hs; a.qmax = h->qmax; a.ldo = 6 * a.L.n_clones + 1;
  a.f_rows = h->f_rows; a.f_gamma = h->f_gamma;
  a.hs_seq_stride = (size_t)h->cfg.max_feats * h->qmax * (h->ncols_max + 1); a.F_alloc = h->cfg.max_feats;
  a.Mmax = h->rho * a.L.n_clones;
  a.ldm = 0;
  a.ssz = feat_ssz(a.Mmax);
  a.per_warp = feat_per_warp(a.Mmax, a.ssz);
  const int ncl = a.L.n_clones, n = 6 * ncl;
  a.nt = (n + 1 + 7) / 8; a.ldz = 8 * a.nt + 4;
  a.G = h->Gws; a.g_seq_stride = (long)h->qr_split_cap * h->gram_n1p * h->gram_n1p; a.n1p = 24 * ((n + 1 + 23) / 24) + 8;
  a.n_acc = h->n_acc; a.max_valid = l.max_valid;
  a.const_sizes = (h->knobs.feat_const != 0) ? 1 : 0;
  a.hs_f32 = h->stack_f32;
  const bool ps = ((size_t)n * n * sizeof(double) <= 72 * 1024) && h->knobs.feat_ps != 0;
  const size_t fixed = 12 * IGV_MAX_CLONES + (ps ? (size_t)n * n : 0);
  // ---- fused Gram accumulation: one CTA of up to 16 warps per sequence -----------------------------------------
  h->feat_fused = false;
  {
    const int ef = h->knobs.fuse;                // test knob: 1 forces the fused kernel, 0 forbids it
    const int qc = h->knobs.qr_cfg;              // a forced compression kernel needs the materialised stack
    const bool allowed = ps && a.nt <= 9 && ncl <= 32 && l.F >= 1 && h->compress != IGV_COMPRESS_HOUSEHOLDER && qc == 0 &&
                         !h->stack_f32;   // the FP32-stack mode is about the materialised stack
    if (allowed && (ef == 1 || (ef != 0 && h->B >= 296))) {
      const int ntt = a.nt * (a.nt + 1) / 2;
      const int ssz = max(a.ssz, 3 * a.ldz + ncl * 27);
      const int pw = feat_per_warp(a.Mmax, ssz);
      const size_t extra = (size_t)(ncl + 1) * ncl * 27 + 2 + (size_t)ntt * 64 + a.ldz + 8 + (ntt + 1) / 2;
      int W = 16;
      while (W > 1 && sizeof(double) * (fixed + (size_t)W * pw + extra) > 222 * 1024) --W;
      const bool room = (size_t)W * pw >= (size_t)(n + 1) * ((n + 1) | 1);   // G is assembled over the per-warp regions
      if (W >= h->knobs.fuse_minw && room) {
        W = (l.F + (l.F + W - 1) / W - 1) / ((l.F + W - 1) / W);   // fewest warps with the same number of rounds
        a.ssz = ssz; a.per_warp = pw;
        const size_t smem = sizeof(double) * (fixed + (size_t)W * pw + extra);
        dim3 grid(1, h->B);
        if (h->rho == 2) launch_feat<2, true, true>(a, grid, W * 32, smem, h->stream);
        else launch_feat<4, true, true>(a, grid, W * 32, smem, h->stream);
        h->feat_fused = true;
        h->launches++;
        return;
      }
    }
  }
  // unfused: mono runs two CTAs of 8 warps per SM; stereo (one CTA per SM: its per-warp scratch is twice as large) takes
  // as many warps as fit, up to 12 (IGV_FEAT_WARPS caps it, for A/B runs)
  int W = (h->rho == 4) ? IGV_STEREO_THREADS / 32 : kWarps;
  if (h->knobs.feat_warps > 0) W = min(W, h->knobs.feat_warps);
  while (W > 1 && sizeof(double) * (fixed + (size_t)W * a.per_warp) > ((h->rho == 4) ? 222 : 200) * 1024) --W;
  const size_t smem = sizeof(double) * (fixed + (size_t)W * a.per_warp);
  // each CTA stages P_s once and its warps loop over tracks: a few CTAs per sequence are enough
  // (staging costs ~10% of the kernel with 5 CTAs per sequence: one CTA per sequence once the batch alone
  // fills the chip, B = 148 * 8 is exactly four waves of the 2 resident CTAs per SM)
  const int per_cta = (h->B >= 296) ? max(l.F, 1) : W;
  const int blocks_x = max(1, min((l.F + per_cta - 1) / per_cta, 64));
  dim3 grid(blocks_x, h->B);
  if (h->rho == 2) {
    if (ps) launch_feat<2, true, false>(a, grid, W * 32, smem, h->stream);
    else launch_feat<2, false, false>(a, grid, W * 32, smem, h->stream);
  } else {
    if (ps) launch_feat<4, true, false>(a, grid, W * 32, smem, h->stream);
    else launch_feat<4, false, false>(a, grid, W * 32, smem, h->stream);
  }
  h->launches++;
}
