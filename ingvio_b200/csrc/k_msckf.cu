// MSCKF per-feature kernel: residual + Jacobian over the clone poses, left null-space projection of
// the landmark, chi^2 gate.  One WARP per feature; lanes own columns of the projected block, so the
// write of the (M-3) x (n+1) block to HBM is coalesced and the clone poses are read once per CTA.
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
// "SELECTED" mode reproduces the reference's overwrite of the anchor's six columns
// (KeyframeUpdate.cpp:523, SwMargUpdate.cpp:127): the anchor's own observation loses its -A_k block.
//
// Algorithmic bytes per feature: reads 24 + 8*rho*nobs (+ 96*SW poses per CTA, P_s through L1/L2),
// writes 8*(M-3)*(n+1).
#include "igv_device.cuh"

using namespace igv;

namespace {

constexpr int kWarps = 8;

struct FeatArgs {
  const double* P; int ld;
  const double* X; int xsize; IgvLayout L;
  int mode, F, obs_slots, rho;
  const double* pf; const int* anchor; const double* obs; const unsigned char* mask; const int* dof;
  double noise2;
  IgvDevParams prm;
  const double* chi2; int chi2_n;
  double* Hs; int qmax; int ldo;     // out block [f][qmax][ldo], ldo = 6*n_clones + 1
  size_t hs_seq_stride; int F_alloc; // per-sequence strides of Hs and of f_rows / f_gamma
  int* f_rows; double* f_gamma;
  int Mmax;                          // rho * n_clones
  int ps_in_smem;                    // clone block of P staged in shared memory (n*n doubles)
};

__device__ __forceinline__ int tri(int i, int j) {  // packed lower index, i >= j
  return i * (i + 1) / 2 + j;
}

__global__ void __launch_bounds__(kWarps * 32, 2) k_msckf_features(FeatArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int ncl = a.L.n_clones, rho = a.rho, n = 6 * ncl, Mmax = a.Mmax;
  // shared: clone poses (12*ncl) | per-warp scratch
  double* sPose = sm;
  const int per_warp = Mmax * 6 /*A,B*/ + Mmax * 3 /*V*/ + Mmax /*r*/ + 16 /*T, tau*/ + Mmax * (Mmax + 1) / 2 /*M0*/ +
                       3 * Mmax /*Y*/ + 2 * Mmax /*slot/k maps as doubles? no: ints below*/;
  double* sPs = sm + 12 * IGV_MAX_CLONES;   // [n][n] clone block of P (symmetric), if staged
  double* ws = sPs + (a.ps_in_smem ? n * n : 0) + (size_t)warp * per_warp;
  double* sA = ws;                     // [M][3]   rows of A_k  (== H_f)
  double* sB = sA + Mmax * 3;          // [M][3]   rows of B_k = A_k [pf]x
  double* sV = sB + Mmax * 3;          // [M][3]   Householder vectors (unit leading entries implicit-explicit)
  double* sr = sV + Mmax * 3;          // [M]      residual
  double* sT = sr + Mmax;              // 9 (T upper-tri) + 3 tau
  double* sM0 = sT + 16;               // packed lower M x M
  double* sY = sM0 + Mmax * (Mmax + 1) / 2;  // [3][M]
  int* sSlot = reinterpret_cast<int*>(sY + 3 * Mmax);  // [ncl] obs k -> slot ; then [ncl] slot -> obs k
  const double* Xb = a.X + (size_t)b * a.xsize;
  for (int t = threadIdx.x; t < 12 * ncl; t += blockDim.x) sPose[t] = Xb[IGV_X_CORE + t];
  const double* Pb = a.P + (size_t)b * a.ld * a.ld;
  if (a.ps_in_smem) {
    // getMarginalCov of the window (StateManager.cpp:128-153), once per CTA; i is the fast index so the
    // column-major global reads are contiguous, and P's symmetry makes the transposed store exact.
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t % n, j = t / n;
      sPs[t] = Pb[(a.L.idx_clone[i / 6] + i % 6) + (size_t)(a.L.idx_clone[j / 6] + j % 6) * a.ld];
    }
  }
  __syncthreads();
  // P element between (clone slot s1, component i) and (clone slot s2, component j)
  auto Pel = [&](int s1, int i, int s2, int j) -> double {
    if (a.ps_in_smem) return sPs[(6 * s1 + i) * n + 6 * s2 + j];
    return __ldg(&Pb[(a.L.idx_clone[s1] + i) + (size_t)(a.L.idx_clone[s2] + j) * a.ld]);
  };

  for (int f = blockIdx.x * nwarps + warp; f < a.F; f += gridDim.x * nwarps) {
    const size_t bf = (size_t)b * a.F + f;          // index into the caller's arrays
    const size_t bo = (size_t)b * a.F_alloc + f;    // index into f_rows / f_gamma
    const double pf[3] = {a.pf[bf * 3], a.pf[bf * 3 + 1], a.pf[bf * 3 + 2]};
    const int anc = a.anchor[bf];
    const unsigned char* mk = a.mask + bf * a.obs_slots;
    const double* ob = a.obs + bf * a.obs_slots * rho;
    int* k2slot = sSlot;
    int* slot2k = sSlot + ncl;
    // ---- (1) per-observation quantities; lanes over slots; NaN rows dropped like the reference ----
    int nobs = 0;
    for (int base = 0; base < ncl; base += 32) {
      const int s = base + lane;
      bool valid = (s < ncl) && mk[s];
      double Ak[12], Bk[12], rk[4];
      if (valid) {
        const double* R = sPose + 12 * s;
        const double d[3] = {pf[0] - R[9], pf[1] - R[10], pf[2] - R[11]};
        double pc[3];
        mat3T_vec(R, d, pc);
        const double iz = 1.0 / pc[2];
        const double h02 = -pc[0] / (pc[2] * pc[2]), h12 = -pc[1] / (pc[2] * pc[2]);
        if (isnan(iz) || isnan(h02) || isnan(h12)) valid = false;  // H_proj.hasNaN()
        // A = Hproj * R^T : row0 = iz*R[:,0]^T + h02*R[:,2]^T ; (R^T)[i][j] = R[j][i]
        for (int j = 0; j < 3; ++j) {
          Ak[j] = iz * R[3 * j + 0] + h02 * R[3 * j + 2];
          Ak[3 + j] = iz * R[3 * j + 1] + h12 * R[3 * j + 2];
        }
        rk[0] = ob[s * rho + 0] - pc[0] * iz;
        rk[1] = ob[s * rho + 1] - pc[1] * iz;
        if (rho == 4) {
          double pr[3];
          mat3_vec(a.prm.Rc, pc, pr);
          for (int i = 0; i < 3; ++i) pr[i] += a.prm.pc[i];
          const double izr = 1.0 / pr[2];
          const double g02 = -pr[0] / (pr[2] * pr[2]), g12 = -pr[1] / (pr[2] * pr[2]);
          // Hproj_r * Rc (2x3), then * R^T
          double Hr[6];
          for (int j = 0; j < 3; ++j) {
            Hr[j] = izr * a.prm.Rc[j] + g02 * a.prm.Rc[6 + j];
            Hr[3 + j] = izr * a.prm.Rc[3 + j] + g12 * a.prm.Rc[6 + j];
          }
          for (int t = 0; t < 2; ++t)
            for (int j = 0; j < 3; ++j)
              Ak[(2 + t) * 3 + j] = Hr[3 * t] * R[3 * j] + Hr[3 * t + 1] * R[3 * j + 1] + Hr[3 * t + 2] * R[3 * j + 2];
          rk[2] = ob[s * rho + 2] - pr[0] * izr;
          rk[3] = ob[s * rho + 3] - pr[1] * izr;
        }
        // B = A * skew(pf):  (a x-matrix) row: a^T [pf]x = (pf x a)^T ... explicit:
        for (int t = 0; t < rho; ++t) {
          const double* ar = Ak + 3 * t;
          Bk[3 * t + 0] = ar[1] * pf[2] - ar[2] * pf[1];
          Bk[3 * t + 1] = ar[2] * pf[0] - ar[0] * pf[2];
          Bk[3 * t + 2] = ar[0] * pf[1] - ar[1] * pf[0];
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      const int k = nobs + __popc(bal & ((1u << lane) - 1u));
      if (s < ncl) slot2k[s] = valid ? k : -1;
      if (valid) {
        k2slot[k] = s;
        for (int t = 0; t < rho; ++t) {
          const int row = k * rho + t;
          for (int j = 0; j < 3; ++j) {
            sA[row * 3 + j] = Ak[3 * t + j];
            sB[row * 3 + j] = Bk[3 * t + j];
            sV[row * 3 + j] = Ak[3 * t + j];
          }
          sr[row] = rk[t];
        }
      }
      nobs += __popc(bal);
    }
    __syncwarp();
    const int M = nobs * rho, q = M - 3;
    if (q < 1) {
      if (lane == 0) { a.f_rows[bo] = 0; a.f_gamma[bo] = nan(""); }
      __syncwarp();
      continue;
    }
    const int kanc = (anc >= 0 && anc < ncl) ? slot2k[anc] : -1;  // observation taken at the anchor clone
    // ---- (2) Householder QR of H_f (M x 3) in sV; T for the compact WY form ----------------------
    double tau[3];
    for (int j = 0; j < 3; ++j) {
      double ss = 0.0;
      for (int i = j + 1 + lane; i < M; i += 32) ss = fma(sV[i * 3 + j], sV[i * 3 + j], ss);
      ss = warp_sum(ss);
      const double alpha = sV[j * 3 + j];
      double beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      double tj = 0.0, scale = 0.0;
      if (ss > 0.0 || alpha != 0.0) {
        if (ss == 0.0) { tj = 0.0; beta = alpha; }
        else { tj = (beta - alpha) / beta; scale = 1.0 / (alpha - beta); }
      }
      tau[j] = tj;
      __syncwarp();
      for (int i = j + 1 + lane; i < M; i += 32) sV[i * 3 + j] *= scale;
      if (lane == 0) sV[j * 3 + j] = 1.0;
      for (int i = lane; i < j; i += 32) sV[i * 3 + j] = 0.0;
      __syncwarp();
      // apply to the remaining columns of H_f
      for (int c = j + 1; c < 3; ++c) {
        double w = 0.0;
        for (int i = j + lane; i < M; i += 32) w = fma(sV[i * 3 + j], sV[i * 3 + c], w);
        w = warp_sum(w) * tj;
        __syncwarp();
        for (int i = j + lane; i < M; i += 32) sV[i * 3 + c] = fma(-w, sV[i * 3 + j], sV[i * 3 + c]);
        __syncwarp();
      }
      (void)beta;
    }
    // note: column c>j of sV now holds the transformed H_f above the diagonal rows (< c); those entries
    // are cleared when column c is processed (rows < c set to 0), so sV ends as the pure V.
    // T (3x3 upper): T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j,0:j] V[:,0:j]^T v_j
    double g01 = 0.0, g02 = 0.0, g12 = 0.0;
    for (int i = lane; i < M; i += 32) {
      g01 = fma(sV[i * 3 + 0], sV[i * 3 + 1], g01);
      g02 = fma(sV[i * 3 + 0], sV[i * 3 + 2], g02);
      g12 = fma(sV[i * 3 + 1], sV[i * 3 + 2], g12);
    }
    g01 = warp_sum(g01); g02 = warp_sum(g02); g12 = warp_sum(g12);
    const double T00 = tau[0], T11 = tau[1], T22 = tau[2];
    const double T01 = -tau[1] * T00 * g01;
    const double T02 = -tau[2] * (T00 * g02 + T01 * g12);
    const double T12 = -tau[2] * (T11 * g12);
    // Q^T x = x - V T^T (V^T x):  z = T^T y, z0 = T00 y0, z1 = T01 y0 + T11 y1, z2 = T02 y0 + T12 y1 + T22 y2
    // ---- (3) M0 = H_x P_s H_x^T (packed lower), one lane per observation pair ---------------------
    {
      const bool drop = (a.mode == IGV_VIS_SELECTED);
      const int npair = nobs * (nobs + 1) / 2;
      for (int pr = lane; pr < npair; pr += 32) {
        int k1 = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
        while (k1 * (k1 + 1) / 2 > pr) --k1;
        while ((k1 + 1) * (k1 + 2) / 2 <= pr) ++k1;
        const int k2 = pr - k1 * (k1 + 1) / 2;  // k1 >= k2
        const int c1 = k2slot[k1], c2 = k2slot[k2];          // clone slots
        const int ca = (anc >= 0 && anc < ncl) ? anc : -1;
        for (int t1 = 0; t1 < rho; ++t1) {
          const int r1 = k1 * rho + t1;
          // row vector of H_x at r1: u1 (6 on clone c1), w1 (3 on anchor rot)
          double u1[6], w1[3];
          const bool a1 = (k1 == kanc);
          for (int j = 0; j < 3; ++j) {
            u1[j] = a1 ? 0.0 : sB[r1 * 3 + j];
            u1[3 + j] = (a1 && drop) ? 0.0 : -sA[r1 * 3 + j];
            w1[j] = (a1 || ca < 0) ? 0.0 : -sB[r1 * 3 + j];
          }
          // g = u1^T P[c1, c2-block] (6) + w1^T P[ca, c2-block] ; ga = u1^T P[c1, ca-rot] + w1^T P[ca-rot, ca-rot]
          double g[6], ga[3];
          for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int i = 0; i < 6; ++i) acc = fma(u1[i], Pel(c1, i, c2, j), acc);
            if (ca >= 0) for (int i = 0; i < 3; ++i) acc = fma(w1[i], Pel(ca, i, c2, j), acc);
            g[j] = acc;
          }
          for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            if (ca >= 0) {
              for (int i = 0; i < 6; ++i) acc = fma(u1[i], Pel(c1, i, ca, j), acc);
              for (int i = 0; i < 3; ++i) acc = fma(w1[i], Pel(ca, i, ca, j), acc);
            }
            ga[j] = acc;
          }
          for (int t2 = 0; t2 < rho; ++t2) {
            const int r2 = k2 * rho + t2;
            if (r2 > r1) continue;
            const bool a2 = (k2 == kanc);
            double acc = 0.0;
            for (int j = 0; j < 3; ++j) {
              const double u2r = a2 ? 0.0 : sB[r2 * 3 + j];
              const double u2t = (a2 && drop) ? 0.0 : -sA[r2 * 3 + j];
              const double w2 = (a2 || ca < 0) ? 0.0 : -sB[r2 * 3 + j];
              acc = fma(g[j], u2r, acc);
              acc = fma(g[3 + j], u2t, acc);
              acc = fma(ga[j], w2, acc);
            }
            sM0[tri(r1, r2)] = acc;
          }
        }
      }
    }
    __syncwarp();
    // ---- (4) two-sided projection  S' = Q^T M0 Q  (only rows/cols >= 3 are needed afterwards) -----
    // left:  M1 = M0 - V Z,  Z = T^T (V^T M0)   (3 x M);  right: M2 = M1 - (M1 V) T V^T.
    // With W = M0 V (M x 3, by symmetry (V^T M0)^T) and C = V^T M0 V (3x3):
    //   Q^T M0 Q = M0 - V T^T W^T - W T V^T + V T^T C T V^T
    {
      for (int i = lane; i < M; i += 32) {
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        for (int k = 0; k < M; ++k) {
          const double m = (i >= k) ? sM0[tri(i, k)] : sM0[tri(k, i)];
          w0 = fma(m, sV[k * 3 + 0], w0);
          w1 = fma(m, sV[k * 3 + 1], w1);
          w2 = fma(m, sV[k * 3 + 2], w2);
        }
        sY[0 * Mmax + i] = w0; sY[1 * Mmax + i] = w1; sY[2 * Mmax + i] = w2;
      }
      __syncwarp();
      double C[9];
      for (int e = 0; e < 9; ++e) {
        const int p = e / 3, qq = e % 3;
        double acc = 0.0;
        for (int i = lane; i < M; i += 32) acc = fma(sV[i * 3 + p], sY[qq * Mmax + i], acc);
        C[e] = warp_sum(acc);
      }
      // D = T^T C T (3x3), with T upper triangular
      const double Tm[9] = {T00, T01, T02, 0.0, T11, T12, 0.0, 0.0, T22};
      double E[9], D[9];
      for (int p = 0; p < 3; ++p) for (int qq = 0; qq < 3; ++qq) {  // E = T^T C
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += Tm[k * 3 + p] * C[k * 3 + qq];
        E[p * 3 + qq] = acc;
      }
      for (int p = 0; p < 3; ++p) for (int qq = 0; qq < 3; ++qq) {  // D = E T
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += E[p * 3 + k] * Tm[k * 3 + qq];
        D[p * 3 + qq] = acc;
      }
      __syncwarp();
      // WT = W T (M x 3):  WT[i][q] = sum_p W[i][p] T[p][q]
      const int nlow = M * (M + 1) / 2;
      for (int e = lane; e < nlow; e += 32) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        if (j < 3) continue;  // only the trailing block is used
        double vi[3], vj[3], wti[3], wtj[3];
        for (int p = 0; p < 3; ++p) { vi[p] = sV[i * 3 + p]; vj[p] = sV[j * 3 + p]; }
        for (int qq = 0; qq < 3; ++qq) {
          double ai = 0.0, aj = 0.0;
          for (int p = 0; p < 3; ++p) { ai += sY[p * Mmax + i] * Tm[p * 3 + qq]; aj += sY[p * Mmax + j] * Tm[p * 3 + qq]; }
          wti[qq] = ai; wtj[qq] = aj;
        }
        double val = sM0[e];
        for (int p = 0; p < 3; ++p) {
          val -= vi[p] * wtj[p];      // (V T^T W^T)[i][j] = sum_p V[i][p] (W T)[j][p]
          val -= wti[p] * vj[p];      // (W T V^T)[i][j]
          for (int qq = 0; qq < 3; ++qq) val += vi[p] * D[p * 3 + qq] * vj[qq];
        }
        sM0[e] = val;
      }
      __syncwarp();
      for (int i = 3 + lane; i < M; i += 32) sM0[tri(i, i)] += a.noise2;
      // r_proj = (Q^T r)[3:]
      double y0 = 0.0, y1 = 0.0, y2 = 0.0;
      for (int i = lane; i < M; i += 32) {
        y0 = fma(sV[i * 3 + 0], sr[i], y0);
        y1 = fma(sV[i * 3 + 1], sr[i], y1);
        y2 = fma(sV[i * 3 + 2], sr[i], y2);
      }
      y0 = warp_sum(y0); y1 = warp_sum(y1); y2 = warp_sum(y2);
      const double z0 = T00 * y0, z1 = T01 * y0 + T11 * y1, z2 = T02 * y0 + T12 * y1 + T22 * y2;
      __syncwarp();
      for (int i = lane; i < M; i += 32)
        sY[i] = sr[i] - (sV[i * 3 + 0] * z0 + sV[i * 3 + 1] * z1 + sV[i * 3 + 2] * z2);   // sY[0:M] = Q^T r
      __syncwarp();
    }
    // ---- (5) Cholesky of S = M0[3:,3:] (q x q, packed lower) and gamma ---------------------------
    bool pd = true;
    for (int j = 0; j < q; ++j) {
      const double d = sM0[tri(3 + j, 3 + j)];
      if (!(d > 0.0)) { pd = false; break; }
      const double dj = sqrt(d);
      __syncwarp();
      if (lane == 0) sM0[tri(3 + j, 3 + j)] = dj;
      for (int i = j + 1 + lane; i < q; i += 32) sM0[tri(3 + i, 3 + j)] /= dj;
      __syncwarp();
      for (int i = j + 1 + lane; i < q; i += 32) {
        const double lij = sM0[tri(3 + i, 3 + j)];
        for (int c = j + 1; c <= i; ++c) sM0[tri(3 + i, 3 + c)] = fma(-lij, sM0[tri(3 + c, 3 + j)], sM0[tri(3 + i, 3 + c)]);
      }
      __syncwarp();
    }
    double gamma = nan("");
    if (pd) {
      if (lane == 0) {
        double g = 0.0;
        for (int i = 0; i < q; ++i) {
          double acc = sY[3 + i];
          for (int k = 0; k < i; ++k) acc = fma(-sM0[tri(3 + i, 3 + k)], sr[k], acc);   // sr reused as y
          acc /= sM0[tri(3 + i, 3 + i)];
          sr[i] = acc;
          g = fma(acc, acc, g);
        }
        gamma = g;
      }
      gamma = __shfl_sync(0xffffffffu, gamma, 0);
    }
    const int dof = a.dof[bf];
    bool accept = pd && dof >= 1 && dof <= a.chi2_n && (gamma < a.chi2[dof - 1]);
    if (lane == 0) { a.f_gamma[bo] = gamma; a.f_rows[bo] = accept ? q : 0; }
    // ---- (6) write the projected block [H | r] rows 3..M-1, row-major, lanes over columns ----------
    if (accept) {
      double* out = a.Hs + (size_t)b * a.hs_seq_stride + (size_t)f * a.qmax * a.ldo;
      const bool drop = (a.mode == IGV_VIS_SELECTED);
      for (int j = lane; j < n + 1; j += 32) {
        if (j == n) {
          for (int i = 3; i < M; ++i) out[(size_t)(i - 3) * a.ldo + j] = sY[i];
          continue;
        }
        const int c = j / 6, comp = j % 6;
        const int kown = slot2k[c];
        const bool anc_rot = (c == anc) && (comp < 3);
        // y = V^T a_j  (sparse column)
        double y0 = 0.0, y1 = 0.0, y2 = 0.0;
        auto col_entry = [&](int k, int t) -> double {
          const int row = k * rho + t;
          double v = 0.0;
          if (k == kown) {
            if (comp < 3) v = (k == kanc) ? 0.0 : sB[row * 3 + comp];
            else v = ((k == kanc) && drop) ? 0.0 : -sA[row * 3 + comp - 3];
          }
          if (anc_rot && k != kanc) v -= sB[row * 3 + comp];
          return v;
        };
        if (anc_rot) {
          for (int k = 0; k < nobs; ++k)
            for (int t = 0; t < rho; ++t) {
              const double v = col_entry(k, t);
              const int row = k * rho + t;
              y0 = fma(sV[row * 3], v, y0); y1 = fma(sV[row * 3 + 1], v, y1); y2 = fma(sV[row * 3 + 2], v, y2);
            }
        } else if (kown >= 0) {
          for (int t = 0; t < rho; ++t) {
            const double v = col_entry(kown, t);
            const int row = kown * rho + t;
            y0 = fma(sV[row * 3], v, y0); y1 = fma(sV[row * 3 + 1], v, y1); y2 = fma(sV[row * 3 + 2], v, y2);
          }
        }
        const double z0 = T00 * y0, z1 = T01 * y0 + T11 * y1, z2 = T02 * y0 + T12 * y1 + T22 * y2;
        for (int i = 3; i < M; ++i) {
          const int k = i / rho, t = i % rho;
          double v = 0.0;
          if (anc_rot || k == kown) v = col_entry(k, t);
          v -= sV[i * 3] * z0 + sV[i * 3 + 1] * z1 + sV[i * 3 + 2] * z2;
          out[(size_t)(i - 3) * a.ldo + j] = v;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace

void igv_launch_msckf_features(igv_batch* h, const IgvMsckfLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_FEATURES);
  FeatArgs a;
  a.P = h->Pc(); a.ld = h->ld;
  a.X = h->Xc(); a.xsize = h->xsize; a.L = h->layout();
  a.mode = l.mode; a.F = l.F; a.obs_slots = l.obs_slots; a.rho = h->rho;
  a.pf = l.pf; a.anchor = l.anchor; a.obs = l.obs; a.mask = l.mask; a.dof = l.dof;
  a.noise2 = l.noise * l.noise;
  a.prm = h->params;
  a.chi2 = h->chi2; a.chi2_n = h->chi2_n;
  a.Hs = h->Hs; a.qmax = h->qmax; a.ldo = 6 * a.L.n_clones + 1;
  a.f_rows = h->f_rows; a.f_gamma = h->f_gamma;
  a.hs_seq_stride = (size_t)h->cfg.max_feats * h->qmax * (h->ncols_max + 1); a.F_alloc = h->cfg.max_feats;
  a.Mmax = h->rho * a.L.n_clones;
  const int Mmax = a.Mmax;
  const size_t per_warp = (size_t)Mmax * 6 + Mmax * 3 + Mmax + 16 + (size_t)Mmax * (Mmax + 1) / 2 + 3 * Mmax + 2 * Mmax;
  const int n = 6 * a.L.n_clones;
  a.ps_in_smem = ((size_t)n * n * sizeof(double) <= 72 * 1024) ? 1 : 0;
  const size_t fixed = 12 * IGV_MAX_CLONES + (a.ps_in_smem ? (size_t)n * n : 0);
  int W = kWarps;
  while (W > 1 && sizeof(double) * (fixed + W * per_warp) > 200 * 1024) --W;
  size_t smem = sizeof(double) * (fixed + W * per_warp);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_msckf_features, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  // each CTA stages P_s once and its warps loop over tracks: a few CTAs per sequence are enough
  const int per_cta = (h->B >= 296) ? 4 * W : W;
  const int blocks_x = max(1, min((l.F + per_cta - 1) / per_cta, 64));
  dim3 grid(blocks_x, h->B);
  k_msckf_features<<<grid, W * 32, smem, h->stream>>>(a);
  h->launches++;
}
