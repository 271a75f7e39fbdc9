// SLAM landmarks kept in the state (mono): measurement rows of the delayed initialisation
// and of the per-frame landmark update, and the H of the anchor change.  One CTA per sequence.
//
// Reference: LandmarkUpdate::calcResJacobianSingleFeatAllMonoObs (LandmarkUpdate.cpp:426-500),
// LandmarkUpdate::calcResJacobianSingleLandmarkMono (:521-572), LandmarkUpdate::updateLandmarkMono (:32-149),
// FeatureInfoManager::changeAnchoredPose (MapServerManager.cpp:343-378).
#include "igv_device.cuh"

using namespace igv;

namespace {

// Rows of one new landmark over the window: per observing clone c with p_c = R_c^T (p_f - p_c), A = Hproj(p_c) R_c^T:
//   H_x[clone c] = [A [p_f]x | -A],  H_x[anchor rotation] -= A [p_f]x  (zero rotation block when c is the anchor),
//   H_f = A,  res = z - pi(p_c).   Valid observations are compacted to the front in clone order; the rest is padding.
__global__ void k_lm_init_rows(const double* X, int xsize, IgvLayout L, IgvLmInitLaunch g) {
  const int b = blockIdx.x, ncl = L.n_clones, n = 6 * ncl, R_ = g.rows_max;
  const double* Xb = X + (size_t)b * xsize;
  double* Hx = g.Hx + (size_t)b * R_ * n;
  double* Hf = g.Hf + (size_t)b * R_ * 3;
  double* rs = g.res + (size_t)b * R_;
  __shared__ int s_row[IGV_MAX_CLONES];
  __shared__ int s_n;
  const double pf[3] = {g.pf[(size_t)b * 3], g.pf[(size_t)b * 3 + 1], g.pf[(size_t)b * 3 + 2]};
  for (int t = threadIdx.x; t < R_ * n; t += blockDim.x) Hx[t] = 0.0;
  for (int t = threadIdx.x; t < R_ * 3; t += blockDim.x) Hf[t] = 0.0;
  for (int t = threadIdx.x; t < R_; t += blockDim.x) rs[t] = 0.0;
  if (threadIdx.x == 0) {
    int k = 0;
    for (int s = 0; s < ncl; ++s) {
      bool ok = g.mask[(size_t)b * g.SW + s] != 0;
      if (ok) {   // H_proj.hasNaN() / H_pf2x.hasNaN() -> the observation is skipped (:479-480)
        const double* Rc = Xb + IGV_X_CORE + 12 * s;
        const double d[3] = {pf[0] - Rc[9], pf[1] - Rc[10], pf[2] - Rc[11]};
        double pc[3];
        mat3T_vec(Rc, d, pc);
        const double iz = 1.0 / pc[2], h02 = -pc[0] / (pc[2] * pc[2]), h12 = -pc[1] / (pc[2] * pc[2]);
        if (isnan(iz) || isnan(h02) || isnan(h12)) ok = false;
      }
      s_row[s] = ok ? k++ : -1;
    }
    s_n = k;
    g.count[b] = 2 * k;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < ncl; s += blockDim.x) {
    const int k = s_row[s];
    if (k < 0) continue;
    const double* Rc = Xb + IGV_X_CORE + 12 * s;
    const double d[3] = {pf[0] - Rc[9], pf[1] - Rc[10], pf[2] - Rc[11]};
    double pc[3];
    mat3T_vec(Rc, d, pc);
    const double iz = 1.0 / pc[2], h02 = -pc[0] / (pc[2] * pc[2]), h12 = -pc[1] / (pc[2] * pc[2]);
    double A[6];   // Hproj R^T, 2 x 3
    for (int j = 0; j < 3; ++j) {
      A[j] = iz * Rc[3 * j + 0] + h02 * Rc[3 * j + 2];
      A[3 + j] = iz * Rc[3 * j + 1] + h12 * Rc[3 * j + 2];
    }
    for (int t = 0; t < 2; ++t) {
      const int row = 2 * k + t;
      const double* ar = A + 3 * t;
      const double bx[3] = {ar[1] * pf[2] - ar[2] * pf[1], ar[2] * pf[0] - ar[0] * pf[2], ar[0] * pf[1] - ar[1] * pf[0]};   // A [pf]x
      for (int j = 0; j < 3; ++j) {
        if (s != g.anchor_slot) {
          Hx[row + (size_t)(6 * s + j) * R_] = bx[j];
          Hx[row + (size_t)(6 * g.anchor_slot + j) * R_] = -bx[j];
        }
        Hx[row + (size_t)(6 * s + 3 + j) * R_] = -ar[j];
        Hf[row + (size_t)j * R_] = ar[j];
      }
      const double z = g.obs[((size_t)b * g.SW + s) * 2 + t];
      rs[row] = z - (t == 0 ? pc[0] : pc[1]) * iz;
    }
  }
}

// Rows of the per-frame landmark update: landmark l, rows 2l, 2l+1 over the columns
//   [SE23 (9) | extrinsics (6) | clones (6 each) | landmarks (3 each)]; gate with dof 2 on the 24 columns it touches.
__global__ void k_lm_update_rows(const double* P, int ld, const double* X, int xsize, IgvLayout L, IgvLmUpdateLaunch g,
                                 const double* chi2, int chi2_n) {
  const int b = blockIdx.x, nl = L.n_lm, ncl = L.n_clones;
  const double* Pb = P + (size_t)b * ld * ld;
  const double* Xb = X + (size_t)b * xsize;
  double* H = g.H + (size_t)b * g.ldh * g.ncols;
  double* rs = g.res + (size_t)b * g.ldh;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  for (int t = threadIdx.x; t < g.ldh * g.ncols; t += blockDim.x) H[t] = 0.0;
  for (int t = threadIdx.x; t < g.ldh; t += blockDim.x) rs[t] = 0.0;
  __syncthreads();
  for (int l = threadIdx.x; l < nl; l += blockDim.x) {
    double gam = nan("");
    const bool on = g.valid[(size_t)b * nl + l] != 0;
    if (on) {
      const double* pf = Xb + L.lm_off + 3 * l;
      const double* R = Xb;            // R_i2w
      const double* Re = Xb + 21;      // R_cl2i
      const double d[3] = {pf[0] - Xb[9], pf[1] - Xb[10], pf[2] - Xb[11]};
      double pi_[3], pcl[3];
      mat3T_vec(R, d, pi_);                                                   // R_i2w^T (pf - p)
      const double e[3] = {pi_[0] - Xb[30], pi_[1] - Xb[31], pi_[2] - Xb[32]};
      mat3T_vec(Re, e, pcl);                                                  // R_cl2i^T (pf_i - p_c2i)
      const double iz = 1.0 / pcl[2], h02 = -pcl[0] / (pcl[2] * pcl[2]), h12 = -pcl[1] / (pcl[2] * pcl[2]);
      double Rw2cl[9];   // R_cl2i^T R_i2w^T
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rw2cl[3 * i + j] = Re[0 + i] * R[3 * j + 0] + Re[3 + i] * R[3 * j + 1] + Re[6 + i] * R[3 * j + 2];
      double A[6], Ae[6];   // Hproj R_w2cl ; Hproj R_cl2i^T
      for (int j = 0; j < 3; ++j) {
        A[j] = iz * Rw2cl[j] + h02 * Rw2cl[6 + j];
        A[3 + j] = iz * Rw2cl[3 + j] + h12 * Rw2cl[6 + j];
        Ae[j] = iz * Re[3 * j + 0] + h02 * Re[3 * j + 2];
        Ae[3 + j] = iz * Re[3 * j + 1] + h12 * Re[3 * j + 2];
      }
      double h[2][24], r2[2];
      int col[24];
      const int a = L.lm_anchor[l];
      for (int t = 0; t < 2; ++t) {
        const double* ar = A + 3 * t;
        const double* ae = Ae + 3 * t;
        const double bx[3] = {ar[1] * pf[2] - ar[2] * pf[1], ar[2] * pf[0] - ar[0] * pf[2], ar[0] * pf[1] - ar[1] * pf[0]};       // A [pf]x
        const double ex[3] = {ae[1] * pi_[2] - ae[2] * pi_[1], ae[2] * pi_[0] - ae[0] * pi_[2], ae[0] * pi_[1] - ae[1] * pi_[0]};   // Ae [pf_i]x
        for (int j = 0; j < 3; ++j) {
          h[t][j] = bx[j]; h[t][3 + j] = -ar[j]; h[t][6 + j] = 0.0;       // SE23
          h[t][9 + j] = ex[j]; h[t][12 + j] = -ae[j];                     // extrinsics
          h[t][15 + j] = -bx[j]; h[t][18 + j] = 0.0;                      // anchor clone
          h[t][21 + j] = ar[j];                                           // landmark
        }
        r2[t] = g.uv[((size_t)b * nl + l) * 2 + t] - (t == 0 ? pcl[0] : pcl[1]) * iz;
      }
      for (int j = 0; j < 9; ++j) col[j] = j;
      for (int j = 0; j < 6; ++j) { col[9 + j] = 15 + j; col[15 + j] = L.idx_clone[a] + j; }
      for (int j = 0; j < 3; ++j) col[21 + j] = L.idx_lm[l] + j;
      // S = H P_s H^T + sigma^2 I (2 x 2), gamma = r^T S^-1 r  (Update.cpp:36-56, dof = res.rows() = 2)
      double s00 = g.noise2, s01 = 0.0, s11 = g.noise2;
      for (int p = 0; p < 24; ++p) {
        double a0 = 0.0, a1 = 0.0;
        for (int q = 0; q < 24; ++q) {
          const double pv = Pb[col[p] + (size_t)col[q] * ld];
          a0 = fma(pv, h[0][q], a0);
          a1 = fma(pv, h[1][q], a1);
        }
        s00 = fma(h[0][p], a0, s00); s01 = fma(h[0][p], a1, s01); s11 = fma(h[1][p], a1, s11);
      }
      const double det = s00 * s11 - s01 * s01;
      gam = (s11 * r2[0] * r2[0] - 2.0 * s01 * r2[0] * r2[1] + s00 * r2[1] * r2[1]) / det;
      const bool acc = chi2_n >= 2 && gam < chi2[1];
      if (acc) {
        // columns of the stacked H: SE23 0..8, extrinsics 9..14, clone s at 15 + 6 s, landmark l at 15 + 6 ncl + 3 l
        for (int t = 0; t < 2; ++t) {
          const int row = 2 * l + t;
          for (int j = 0; j < 15; ++j) H[row + (size_t)j * g.ldh] = h[t][j];
          for (int j = 0; j < 3; ++j) H[row + (size_t)(15 + 6 * a + j) * g.ldh] = h[t][15 + j];
          for (int j = 0; j < 3; ++j) H[row + (size_t)(15 + 6 * ncl + 3 * l + j) * g.ldh] = h[t][21 + j];
          rs[row] = r2[t];
        }
        atomicAdd(&s_cnt, 1);
      }
    }
    if (g.gamma) g.gamma[(size_t)b * nl + l] = gam;
  }
  __syncthreads();
  if (threadIdx.x == 0) g.n_acc[b] = s_cnt;
}

// H (3 x 15, col-major) of the anchor change on (old anchor 6, new anchor 6, landmark 3): [-[pf]x 0 | [pf]x 0 | I]
__global__ void k_lm_anchor_H(const double* X, int xsize, int lm_off, int lm_slot, double* H, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* pf = X + (size_t)b * xsize + lm_off + 3 * lm_slot;
  double* Hb = H + (size_t)b * 45;
  for (int t = 0; t < 45; ++t) Hb[t] = 0.0;
  const double sk[9] = {0.0, -pf[2], pf[1], pf[2], 0.0, -pf[0], -pf[1], pf[0], 0.0};   // [pf]x row-major
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Hb[i + 3 * j] = -sk[3 * i + j];
      Hb[i + 3 * (6 + j)] = sk[3 * i + j];
    }
  for (int i = 0; i < 3; ++i) Hb[i + 3 * (12 + i)] = 1.0;
}

__global__ void k_set_lm_value(double* X, int xsize, int lm_off, int lm_slot, const double* pf, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < 3; ++i) X[(size_t)b * xsize + lm_off + 3 * lm_slot + i] = pf ? pf[(size_t)b * 3 + i] : 0.0;
}

}  // namespace

void igv_launch_lm_init_rows(igv_batch* h, const IgvLmInitLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_lm_init_rows<<<h->B, 64, 0, h->stream>>>(h->Xc(), h->xsize, h->layout(), l);
  h->launches++;
}
void igv_launch_lm_update_rows(igv_batch* h, const IgvLmUpdateLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_lm_update_rows<<<h->B, 32, 0, h->stream>>>(h->Pc(), h->ld, h->Xc(), h->xsize, h->layout(), l, h->chi2, h->chi2_n);
  h->launches++;
}
void igv_launch_lm_anchor_H(igv_batch* h, int lm_slot, double* H) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_lm_anchor_H<<<(h->B + 63) / 64, 64, 0, h->stream>>>(h->Xc(), h->xsize, h->layout().lm_off, lm_slot, H, h->B);
  h->launches++;
}
void igv_launch_set_lm_value(igv_batch* h, int lm_slot, const double* pf_dev) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_set_lm_value<<<(h->B + 63) / 64, 64, 0, h->stream>>>(h->Xc(), h->xsize, h->layout().lm_off, lm_slot, pf_dev, h->B);
  h->launches++;
}
