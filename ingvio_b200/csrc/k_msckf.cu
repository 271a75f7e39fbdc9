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
#include "igv_device.cuh"

using namespace igv;

namespace {

constexpr int kWarps = 8;

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
  int ldm;                           // row stride of the per-warp S matrix (odd)
  int per_warp;                      // doubles of per-warp scratch
};

// per-warp scratch layout (doubles): A[M][3] B[M][3] V[M][3] Am[M][3] E[M][3] r[M] qr[M] S[(M+1)][ldm] maps
__host__ __device__ inline int feat_per_warp(int Mmax, int ldm) {
  return Mmax * 15 + 2 * Mmax + (Mmax + 1) * ldm + 2 * Mmax + 8;
}

template <int RHO, bool PS_SMEM, int QT>   // QT: largest projected block (rows) whose gate runs in registers, see (5a)
__global__ void __launch_bounds__(kWarps * 32, 2) k_msckf_features(FeatArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int ncl = a.L.n_clones, n = 6 * ncl, Mmax = a.Mmax, ldm = a.ldm;
  double* sPose = sm;                                   // 12 per clone
  double* sPs = sm + 12 * IGV_MAX_CLONES;               // [n][n] clone block of P (symmetric), if staged
  double* ws = sPs + (PS_SMEM ? n * n : 0) + (size_t)warp * a.per_warp;
  double* sA = ws;                 // [M][3] rows of A_k (== H_f)
  double* sB = sA + Mmax * 3;      // [M][3] rows of B_k = A_k [pf]x
  double* sV = sB + Mmax * 3;      // [M][3] Householder vectors
  double* sAm = sV + Mmax * 3;     // [M][3] W T
  double* sE = sAm + Mmax * 3;     // [M][3] W T - V D
  double* sr = sE + Mmax * 3;      // [M] residual
  double* sqr = sr + Mmax;         // [M] Q^T r
  double* sS = sqr + Mmax;         // [(M+1)][ldm] : H_x P_s H_x^T, then S and its Cholesky factor (+ rhs row)
  int* k2slot = reinterpret_cast<int*>(sS + (Mmax + 1) * ldm);  // [ncl] obs k -> slot
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
  __syncthreads();
  const bool drop = (a.mode == IGV_VIS_SELECTED);

  for (int f = blockIdx.x * nwarps + warp; f < a.F; f += gridDim.x * nwarps) {
    const size_t bf = (size_t)b * a.F + f;          // index into the caller's arrays
    const size_t bo = (size_t)b * a.F_alloc + f;    // index into f_rows / f_gamma
    if (a.feat_ok && !a.feat_ok[bf]) {              // track rejected upstream (e.g. triangulation failed)
      if (lane == 0) { a.f_rows[bo] = 0; a.f_gamma[bo] = nan(""); }
      continue;
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
      continue;
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
    // ---- (3) M0 = H_x P_s H_x^T into sS (full storage), one lane per observation pair -------------
    {
      const int npair = nobs * (nobs + 1) / 2;
      for (int pr = lane; pr < npair; pr += 32) {
        int k1 = (int)((sqrtf(8.f * pr + 1.f) - 1.f) * 0.5f);
        while (k1 * (k1 + 1) / 2 > pr) --k1;
        while ((k1 + 1) * (k1 + 2) / 2 <= pr) ++k1;
        const int k2 = pr - k1 * (k1 + 1) / 2;  // k1 >= k2
        const int c1 = k2slot[k1], c2 = k2slot[k2];
        // strides of the P blocks: SMEM copy is [n][n]; the global matrix is column-major (symmetric)
        const int rs = PS_SMEM ? n : 1;
        const long cs = PS_SMEM ? 1 : a.ld;
        const double* p12 = PS_SMEM ? sPs + (6 * c1) * n + 6 * c2 : Pb + a.L.idx_clone[c1] + (long)a.L.idx_clone[c2] * a.ld;
        const double* pa2 = nullptr; const double* p1a = nullptr; const double* paa = nullptr;
        if (ca >= 0) {
          pa2 = PS_SMEM ? sPs + (6 * ca) * n + 6 * c2 : Pb + a.L.idx_clone[ca] + (long)a.L.idx_clone[c2] * a.ld;
          p1a = PS_SMEM ? sPs + (6 * c1) * n + 6 * ca : Pb + a.L.idx_clone[c1] + (long)a.L.idx_clone[ca] * a.ld;
          paa = PS_SMEM ? sPs + (6 * ca) * n + 6 * ca : Pb + a.L.idx_clone[ca] + (long)a.L.idx_clone[ca] * a.ld;
        }
        const bool a1 = (k1 == kanc), a2 = (k2 == kanc);
#pragma unroll
        for (int t1 = 0; t1 < RHO; t1 += 2) {
          const int r1 = k1 * RHO + t1;
          // rows r1, r1+1 of H_x: u (6 on clone c1), w (3 on the anchor's rotation columns)
          double u[2][6], w[2][3];
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double bb = sB[(r1 + t) * 3 + j], aa = sA[(r1 + t) * 3 + j];
              u[t][j] = a1 ? 0.0 : bb;
              u[t][3 + j] = (a1 && drop) ? 0.0 : -aa;
              w[t][j] = (a1 || ca < 0) ? 0.0 : -bb;
            }
          // g = row * P[:, clone c2 block] (6),  ga = row * P[:, anchor rot] (3)
          double g[2][6], ga[2][3];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int j = 0; j < 6; ++j) g[t][j] = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) ga[t][j] = 0.0;
          }
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const double p = p12[i * rs + j * cs];
              g[0][j] = fma(u[0][i], p, g[0][j]);
              g[1][j] = fma(u[1][i], p, g[1][j]);
            }
          if (ca >= 0 && !a1) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                const double p = pa2[i * rs + j * cs];
                g[0][j] = fma(w[0][i], p, g[0][j]);
                g[1][j] = fma(w[1][i], p, g[1][j]);
              }
          }
          if (ca >= 0 && !a2) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const double p = p1a[i * rs + j * cs];
                ga[0][j] = fma(u[0][i], p, ga[0][j]);
                ga[1][j] = fma(u[1][i], p, ga[1][j]);
              }
            if (!a1) {
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  const double p = paa[i * rs + j * cs];
                  ga[0][j] = fma(w[0][i], p, ga[0][j]);
                  ga[1][j] = fma(w[1][i], p, ga[1][j]);
                }
            }
          }
#pragma unroll
          for (int t2 = 0; t2 < RHO; ++t2) {
            const int r2 = k2 * RHO + t2;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double bb = sB[r2 * 3 + j], aa = sA[r2 * 3 + j];
              const double u2r = a2 ? 0.0 : bb;
              const double u2t = (a2 && drop) ? 0.0 : -aa;
              const double w2 = (a2 || ca < 0) ? 0.0 : -bb;
              acc0 = fma(g[0][j], u2r, acc0); acc0 = fma(g[0][3 + j], u2t, acc0); acc0 = fma(ga[0][j], w2, acc0);
              acc1 = fma(g[1][j], u2r, acc1); acc1 = fma(g[1][3 + j], u2t, acc1); acc1 = fma(ga[1][j], w2, acc1);
            }
            if (r2 <= r1) { sS[r1 * ldm + r2] = acc0; sS[r2 * ldm + r1] = acc0; }
            if (r2 <= r1 + 1) { sS[(r1 + 1) * ldm + r2] = acc1; sS[r2 * ldm + r1 + 1] = acc1; }
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
        const double* mrow = sS + i * ldm;
        for (int k = 0; k < M; ++k) {
          const double m = mrow[k];
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
        if (i >= 3) sS[M * ldm + i] = qv;       // extra row M of S: the right-hand side of L y = r_proj
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
      const double* srow = sS + ii * ldm;
      double ar[QT];
#pragma unroll
      for (int c = 0; c < QT; ++c) {
        const int j = min(c + 3, M - 1);
        double val = srow[j] - (v0 * sAm[j * 3] + v1 * sAm[j * 3 + 1] + v2 * sAm[j * 3 + 2] +
                                e0 * sV[j * 3] + e1 * sV[j * 3 + 1] + e2 * sV[j * 3 + 2]);
        if (c == row) val += a.noise2;
        if (row == q) val = sqr[j];
        ar[c] = (c < q) ? val : 0.0;
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
        double* srow = sS + ii * ldm;
        const int jmax = min(M, i0 + 32);
        for (int j = 3; j < jmax; ++j) {
          const double val = srow[j] - (v0 * sAm[j * 3] + v1 * sAm[j * 3 + 1] + v2 * sAm[j * 3 + 2] +
                                        e0 * sV[j * 3] + e1 * sV[j * 3 + 1] + e2 * sV[j * 3 + 2]);
          if (on && j <= i) srow[j] = (j == i) ? val + a.noise2 : val;
        }
      }
      __syncwarp();
      // ---- (5b) Cholesky of S = sS[3:M,3:M] (lower) with the rhs as row M: L y = r_proj rides along ----
      for (int j = 3; j < M; ++j) {
        const double d = sS[j * ldm + j];
        if (!(d > 0.0)) { pd = false; break; }
        const double inv = rsqrt(d);
        __syncwarp();
        for (int i = j + 1 + lane; i <= M; i += 32) sS[i * ldm + j] *= inv;   // rows j+1..M (incl. rhs row)
        if (lane == 0) sS[j * ldm + j] = d * inv;
        __syncwarp();
        for (int i0 = j + 1; i0 <= M; i0 += 32) {
          const int i = i0 + lane;
          const bool on = i <= M;
          const int ii = on ? i : j + 1;
          double* srow = sS + ii * ldm;
          const double lij = srow[j];
          const int cmax = min(M - 1, i0 + 31);   // columns j+1..min(i, M-1): rhs row only has columns < M
          for (int c = j + 1; c <= cmax; ++c) {
            const double lcj = sS[c * ldm + j];
            if (on && c <= i) srow[c] = fma(-lij, lcj, srow[c]);
          }
        }
        __syncwarp();
      }
      if (pd) {
        double g = 0.0;
        for (int c = 3 + lane; c < M; c += 32) { const double y = sS[M * ldm + c]; g = fma(y, y, g); }
        gamma = warp_sum(g);
      }
    }
    const int dof = a.dof[bf];
    const bool accept = pd && dof >= 1 && dof <= a.chi2_n && (gamma < a.chi2[dof - 1]);
    if (lane == 0) { a.f_gamma[bo] = gamma; a.f_rows[bo] = accept ? q : 0; }
    // ---- (6) write the projected block [H | r] rows 3..M-1, row-major, lanes over columns ----------
    if (accept) {
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
            double v = -(s0 * z[gq][0] + s1 * z[gq][1] + s2 * z[gq][2]);
            const int t = i - own_lo[gq];
#pragma unroll
            for (int tt = 0; tt < RHO; ++tt) if (tt == t) v += ownv[gq][tt];
            if (acomp[gq] >= 0 && not_anchor_row) v -= sB[i * 3 + acomp[gq]];
            if (isres[gq]) v = qri;
            if (live[gq]) orow[32 * gq] = v;
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int RHO, bool PS, int QT>
void launch_feat_q(const FeatArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_msckf_features<RHO, PS, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  k_msckf_features<RHO, PS, QT><<<grid, threads, smem, st>>>(a);
}

template <int RHO, bool PS>
void launch_feat(const FeatArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  // register-resident gate sized to the largest block this window can produce (rho * clones - 3 rows)
  const int qmax = a.Mmax - 3;
  if (qmax <= 9) launch_feat_q<RHO, PS, 9>(a, grid, threads, smem, st);
  else if (qmax <= 17) launch_feat_q<RHO, PS, 17>(a, grid, threads, smem, st);
  else if (qmax <= 21) launch_feat_q<RHO, PS, 21>(a, grid, threads, smem, st);
  else launch_feat_q<RHO, PS, 25>(a, grid, threads, smem, st);
}

}  // namespace

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
  a.ldm = a.Mmax | 1;
  a.per_warp = feat_per_warp(a.Mmax, a.ldm);
  const int n = 6 * a.L.n_clones;
  const bool ps = ((size_t)n * n * sizeof(double) <= 72 * 1024);
  const size_t fixed = 12 * IGV_MAX_CLONES + (ps ? (size_t)n * n : 0);
  int W = kWarps;
  while (W > 1 && sizeof(double) * (fixed + (size_t)W * a.per_warp) > 200 * 1024) --W;
  const size_t smem = sizeof(double) * (fixed + (size_t)W * a.per_warp);
  // each CTA stages P_s once and its warps loop over tracks: a few CTAs per sequence are enough
  // (staging costs ~10% of the kernel with 5 CTAs per sequence: one CTA per sequence once the batch alone
  // fills the chip, B = 148 * 8 is exactly four waves of the 2 resident CTAs per SM)
  const int per_cta = (h->B >= 296) ? max(l.F, 1) : W;
  const int blocks_x = max(1, min((l.F + per_cta - 1) / per_cta, 64));
  dim3 grid(blocks_x, h->B);
  if (h->rho == 2) {
    if (ps) launch_feat<2, true>(a, grid, W * 32, smem, h->stream);
    else launch_feat<2, false>(a, grid, W * 32, smem, h->stream);
  } else {
    if (ps) launch_feat<4, true>(a, grid, W * 32, smem, h->stream);
    else launch_feat<4, false>(a, grid, W * 32, smem, h->stream);
  }
  h->launches++;
}
