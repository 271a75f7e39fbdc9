// EKF update on the type-indexed covariance: Z = H P_s^T-gather, S = H Z_s + R, Cholesky,
// Y = L^-1 Z, P -= Y^T Y, dx = Y^T (L^-1 res), box-plus.  One CTA per sequence.
//
// Reference: StateManager::ekfUpdate (StateManager.cpp:359-426) -- PH^T by looping all variable
// blocks x measured blocks (:385-397), S = H P_s H^T + R (:399-403), K = PH^T S^-1 via an explicit
// LU inverse (:405), P <- sym(P - K (PH^T)^T) (:407-411), negative-diagonal check (:413-421),
// dx = K res (:423), boxPlus (:425); UpdateBase::whitenResidual (Update.cpp:36-79) for gamma_only.
//
// With S = L L^T and Y = L^-1 (PH^T)^T:  K (PH^T)^T = Y^T Y (symmetric by construction, so the
// reference's explicit symmetrisation is a no-op here) and dx = Y^T L^-1 res.  Algebraically
// identical to the reference; rounding differs at the 1e-15 relative level.
//
// Algorithmic traffic per sequence: read P (8N^2) + write P (8N^2) + H/res (8 r (n+1)) + dx (8N).
#include "igv_device.cuh"

using namespace igv;

namespace {

struct EkfArgs {
  double* P; int ld; int N;
  double* X; int xsize; IgvLayout L;
  IgvBlocks blk;
  int rows;
  const double* H; long strideH; int h_ld; int h_rowmajor;
  const double* res; long strideRes; int res_inc;
  const double* R; long strideR; int r_kind; double r_iso_value; const int* only_if;
  double* Zws; long strideZ; double* Sws; long strideS;
  int z_in_smem, s_in_smem, ld_pad;
  int h_upper;               // H is upper triangular (rows = compressed [R | y]): leading zeros are skipped
  double* dx_out; double* dxws;
  int gamma_only; double* gamma_out;
  const int* gate_rows; const double* chi2; int chi2_n;
  int apply_boxplus;
  int* flags;
};

// THREADS per CTA: 384 (12 warps, two CTAs per SM at 80 registers) for the wide visual update; few measurement rows
// (GNSS, delayed initialisation, gates) are bound by the barriers of the serial Cholesky, so they run as smaller CTAs,
// more of which are resident per SM (independent barrier domains overlap).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 384 ? 2 : (THREADS >= 256 ? 3 : 6)) k_ekf_update(EkfArgs a) {
  extern __shared__ double sm[];
  __shared__ int cols[6 * IGV_MAX_BLOCKS];
  __shared__ int s_ok;
  __shared__ double s_red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int N = a.N, ld = a.ld, r = a.rows, n = a.blk.n;
  double* Pb = a.P + (size_t)b * ld * ld;
  const double* Hb = a.H + (size_t)b * a.strideH;
  const double* resb = a.res + (size_t)b * a.strideRes;
  // Z is r x (N+1) column-major: column i = (H P[:, i]) ; column N = res. In shared memory the leading dimensions
  // are padded to 4 (mod 8) doubles: the 4 x 8 operand fragments of the DMMA stages then hit distinct banks.
  const int ldz = a.z_in_smem ? a.ld_pad : r;
  const int lds = a.s_in_smem ? a.ld_pad : r;
  if (a.only_if && !a.only_if[b]) {
    if (a.dx_out) for (int i = tid; i < N; i += blockDim.x) a.dx_out[(size_t)b * N + i] = 0.0;
    return;
  }
  double* smp = sm;
  double* Z = a.z_in_smem ? smp : a.Zws + (size_t)b * a.strideZ;
  if (a.z_in_smem) smp += (size_t)ldz * (N + 1);
  double* S = a.s_in_smem ? smp : a.Sws + (size_t)b * a.strideS;

  for (int q = tid; q < a.blk.n_blocks; q += blockDim.x) {
    int off = 0;
    for (int t = 0; t < q; ++t) off += a.blk.size[t];
    for (int k = 0; k < a.blk.size[q]; ++k) cols[off + k] = a.blk.idx[q] + k;
  }
  __syncthreads();
  auto Hat = [&](int i, int k) -> double {
    return a.h_rowmajor ? Hb[(size_t)i * a.h_ld + k] : Hb[i + (size_t)k * a.h_ld];
  };
  // (1) Z[a_, i] = sum_c H[a_, c] * P[i, cols[c]]     (StateManager.cpp:385-397, transposed)
  //     the thread-fast index is i so that P reads are contiguous.
  {
    const int Ni = a.gamma_only ? n : N;  // gate only needs the measured rows of P
    // H upper triangular (the compressed [R | y]): row j of H is zero left of column j, so the contraction of
    // the output block of rows j0.. starts at k = j0
    cta_gemm_mma_ex<2, 2>(Ni, r, n,
                          [&](int i, int k) { const int ii = a.gamma_only ? cols[i] : i; return Pb[ii + (size_t)cols[k] * ld]; },
                          [&](int k, int j) { return Hat(j, k); },
                          [&](int i, int j, double v, double) { Z[j + (size_t)i * ldz] = v; },
                          [](int, int) { return false; },
                          [&](int, int j0) { return a.h_upper ? (j0 & ~3) : 0; }, [](int, int) { return 0.0; });
    for (int t = tid; t < r; t += blockDim.x) Z[t + (size_t)N * ldz] = resb[(size_t)t * a.res_inc];
  }
  __syncthreads();
  // (2) S = H * Z[:, cols] + R                          (StateManager.cpp:399-403)
  {
    const double* Rb = a.R ? a.R + (size_t)b * a.strideR : nullptr;
    // only the lower triangle of S is used by the factorisation: blocks above the diagonal are skipped
    cta_gemm_mma_ex<2, 2>(r, r, n, [&](int i, int k) { return Hat(i, k); },
                       [&](int k, int j) { const int cc = a.gamma_only ? k : cols[k]; return Z[j + (size_t)cc * ldz]; },
                       [&](int i, int j, double v, double) {
                         double rr = 0.0;
                         if (a.r_kind == IGV_R_ISO) rr = (i == j) ? (Rb ? Rb[0] : a.r_iso_value) : 0.0;
                         else if (a.r_kind == IGV_R_DIAG) rr = (i == j) ? Rb[i] : 0.0;
                         else rr = Rb[i + (size_t)j * r];
                         S[i + (size_t)j * lds] = v + rr;
                       },
                       [](int bi, int bj) { return bj > bi + 15; },
                       [&](int i0, int) { return a.h_upper ? (i0 & ~3) : 0; }, [](int, int) { return 0.0; });
  }
  __syncthreads();
  // (3) S = L L^T
  // (3)+(4)+(5) S = L L^T fused with Y = L^-1 Z (all N columns) and w = L^-1 res (column N)
  const bool ok = a.gamma_only ? cta_chol_solve_fused<IGV_EKF_NB>(S, r, lds, Z, ldz, N, 1, &s_ok)
                               : cta_chol_solve_fused<IGV_EKF_NB>(S, r, lds, Z, ldz, 0, N + 1, &s_ok);
  if (!ok) {
    if (tid == 0) {
      atomicOr(&a.flags[b], IGV_FLAG_CHOL_FAIL);
      if (a.gamma_out) a.gamma_out[b] = nan("");
    }
    if (a.dx_out) for (int i = tid; i < N; i += blockDim.x) a.dx_out[(size_t)b * N + i] = 0.0;
    return;
  }
  if (tid < 32) {  // gamma = |w|^2
    const double* w = Z + (size_t)N * ldz;
    double g = 0.0;
    for (int i = tid; i < r; i += 32) g = fma(w[i], w[i], g);
    g = warp_sum(g);
    if (tid == 0) { s_red[0] = g; if (a.gamma_out) a.gamma_out[b] = g; }
  }
  __syncthreads();
  if (a.gamma_only) return;
  if (a.gate_rows) {  // GnssUpdate.cpp:286-287 joint "strong reject"
    const int cnt = a.gate_rows[b];
    bool reject = false;
    if (cnt <= 0) reject = true;
    else if (cnt <= 14) {
      const double thr = (cnt <= a.chi2_n) ? a.chi2[cnt - 1] : INFINITY;
      reject = !(s_red[0] < thr);
    }
    if (reject) {
      if (tid == 0 && cnt > 0) atomicOr(&a.flags[b], IGV_FLAG_GNSS_REJECTED);
      if (a.dx_out) for (int i = tid; i < N; i += blockDim.x) a.dx_out[(size_t)b * N + i] = 0.0;
      return;
    }
  }
  // (6) dx = Y^T w
  double* dxb = a.dxws + (size_t)b * ld;
  for (int i = tid; i < N; i += blockDim.x) {
    const double* y = Z + (size_t)i * ldz;
    const double* w = Z + (size_t)N * ldz;
    double acc = 0.0;
    for (int k = 0; k < r; ++k) acc = fma(y[k], w[k], acc);
    dxb[i] = acc;
    if (a.dx_out) a.dx_out[(size_t)b * N + i] = acc;
  }
  // (7) P -= Y^T Y : blocks of the lower triangle on DMMA, mirrored on store
  cta_gemm_mma_ex<2, 2>(N, N, r, [&](int i, int k) { return Z[k + (size_t)i * ldz]; },
                        [&](int k, int j) { return Z[k + (size_t)j * ldz]; },
                        [&](int i, int j, double v, double old) {
                          if (j <= i) {
                            const double val = old - v;
                            Pb[i + (size_t)j * ld] = val;
                            if (i != j) Pb[j + (size_t)i * ld] = val;
                          }
                        },
                        [](int bi, int bj) { return bj > bi + 15; }, [](int, int) { return 0; },
                        [&](int i, int j) { return Pb[i + (size_t)j * ld]; });
  __syncthreads();
  // (8) negative diagonal check (StateManager.cpp:413-421) and box-plus (:425)
  for (int i = tid; i < N; i += blockDim.x)
    if (Pb[i + (size_t)i * ld] < 0.0) atomicOr(&a.flags[b], IGV_FLAG_NEG_DIAG);
  if (a.apply_boxplus) boxplus_all(a.X + (size_t)b * a.xsize, dxb, a.L);
}

}  // namespace

void igv_launch_ekf(igv_batch* h, const IgvEkfLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_EKF);
  EkfArgs a;
  a.P = h->Pc(); a.ld = h->ld; a.N = h->N;
  a.X = h->Xc(); a.xsize = h->xsize; a.L = h->layout();
  a.blk = l.blk; a.rows = l.rows;
  a.H = l.H; a.strideH = l.strideH; a.h_ld = l.h_ld; a.h_rowmajor = l.h_rowmajor;
  a.res = l.res; a.strideRes = l.strideRes; a.res_inc = l.res_inc;
  a.R = l.R; a.strideR = l.strideR; a.r_kind = l.r_kind; a.r_iso_value = l.r_iso_value; a.only_if = l.only_if;
  a.Zws = h->Zws; a.strideZ = (long)h->max_rows * (h->ld + 1);
  a.Sws = h->Sws; a.strideS = (long)h->max_rows * h->max_rows;
  a.dx_out = l.dx_out; a.dxws = h->dxws;
  a.gamma_only = l.gamma_only; a.gamma_out = l.gamma_out;
  a.gate_rows = l.gate_rows; a.chi2 = h->chi2; a.chi2_n = h->chi2_n;
  a.apply_boxplus = l.apply_boxplus; a.flags = h->flags;
  a.h_upper = l.h_upper;
  // shared-memory placement: Z first, then S, as long as they fit
  const size_t cap = 200 * 1024;
  const int ld_pad = l.rows + ((4 - l.rows % 8) + 8) % 8;   // smallest >= rows that is 4 (mod 8)
  a.ld_pad = ld_pad;
  const size_t zb = sizeof(double) * (size_t)ld_pad * (h->N + 1), sb = sizeof(double) * (size_t)ld_pad * l.rows;
  size_t smem = 0;
  a.z_in_smem = (zb <= cap) ? 1 : 0;
  if (a.z_in_smem) smem += zb;
  a.s_in_smem = (smem + sb <= cap) ? 1 : 0;
  if (a.s_in_smem) smem += sb;
  // knobs: IGV_EKF_T_SMALL / IGV_EKF_T_BIG (threads per CTA for rows <= 32 / above)
  // measured (c2, B200): 256-thread CTAs for the 24-row GNSS update gain 4 % of the EKF time once the batch fills the
  // chip (B = 1184) and lose 2 % at B = 8, where one CTA per sequence is all there is
  int want = (l.rows <= 32) ? h->knobs.ekf_t_small : h->knobs.ekf_t_big;
  if (want == 0) want = (l.rows <= 32 && h->B >= 296) ? 256 : 384;
  if (want <= 128) {
    IGV_SMEM_OPTIN((k_ekf_update<128>), 220 * 1024);
    k_ekf_update<128><<<h->B, 128, smem, h->stream>>>(a);
  } else if (want <= 256) {
    IGV_SMEM_OPTIN((k_ekf_update<256>), 220 * 1024);
    k_ekf_update<256><<<h->B, 256, smem, h->stream>>>(a);
  } else {
    IGV_SMEM_OPTIN((k_ekf_update<384>), 220 * 1024);
    k_ekf_update<384><<<h->B, 384, smem, h->stream>>>(a);
  }
  h->launches++;
}
