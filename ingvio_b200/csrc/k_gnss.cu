// GNSS pseudo-range / Doppler measurement rows.  One CTA per sequence, one thread per satellite.
//
// Reference: GnssUpdate::updateTrackedSys (GnssUpdate.cpp:124-284): per satellite a pseudo-range row
//   [u^T R_w2e [p]x , -u^T R_w2e , 0 | yof | 1 on its clock bias]   res = -res_pos   (:161-169,:194-208)
// and a Doppler row
//   [u^T R_w2e [v]x , 0 , -u^T R_w2e | yof | 1 on the clock drift]  res = -res_vel   (:236-244,:263-269)
// with R_w2e = R_enu2ecef * Rz(yaw offset) (:136, GnssManager.cpp:60-63), optional per-row chi^2
// (:190,:259).  Column order here is fixed as [SE23(9), YOF, clock biases present in state, FS];
// the reference orders clock-bias columns by first appearance (:200-206), which only permutes
// columns.  Rejected / untracked rows are left as zero rows with unit noise: they change neither the
// posterior nor the joint gate statistic, and `cnt` carries the true row count for that gate.
#include "igv_device.cuh"

using namespace igv;

namespace {

struct GnssArgs {
  const double* P; int ld;
  const double* X; int xsize;
  int S;
  const double* unit; const double* res_pos; const double* res_vel; const double* sig_psr; const double* sig_dopp;
  const int* sys; const double* Renu;
  int adjust_yof, chi2_test;
  int idx_yof; int idx_gnss[6]; int col_of_gnss[6]; int ncols;
  double* H; int ldh; double* res; double* Rd; int* cnt;
  const double* chi2; int chi2_n;
};

__global__ void k_gnss_rows(GnssArgs a) {
  const int b = blockIdx.x, S = a.S, rows = 2 * S;
  const double* Pb = a.P + (size_t)b * a.ld * a.ld;
  const double* Xb = a.X + (size_t)b * a.xsize;
  double* Hb = a.H + (size_t)b * a.ldh * 16;
  double* rb = a.res + (size_t)b * a.ldh;
  double* Rb = a.Rd + (size_t)b * a.ldh;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  for (int t = threadIdx.x; t < a.ldh * a.ncols; t += blockDim.x) Hb[t] = 0.0;
  __syncthreads();
  const double yof = Xb[33 + IGV_GNSS_YOF];
  double sy, cy;
  sincos(yof, &sy, &cy);
  const double* Re = a.Renu + (size_t)b * 9;
  const double Rz[9] = {cy, -sy, 0.0, sy, cy, 0.0, 0.0, 0.0, 1.0};
  const double dRz[9] = {-sy, -cy, 0.0, cy, -sy, 0.0, 0.0, 0.0, 0.0};
  double Rw[9], dRw[9];
  mat3_mul(Re, Rz, Rw);
  mat3_mul(Re, dRz, dRw);
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const bool is_psr = r < S;
    const int i = is_psr ? r : r - S;
    const size_t bi = (size_t)b * S + i;
    const int g = a.sys[bi];
    double h[11];
    for (int k = 0; k < 11; ++k) h[k] = 0.0;
    int col_last = -1, idx_last = -1;
    bool ok = (g >= 0 && g < 4 && a.idx_gnss[g] >= 0);
    double resv = 0.0, sig = 1.0;
    if (ok) {
      const double* u = a.unit + bi * 3;
      double uR[3], udR[3];
      mat3T_vec(Rw, u, uR);    // u^T R_w2e
      mat3T_vec(dRw, u, udR);
      const double* x = is_psr ? (Xb + 9) : (Xb + 12);  // p or v
      // (u^T R [x]x)_j :  a^T [x]x = (a2 x3 - a3 x2, a3 x1 - a1 x3, a1 x2 - a2 x1)
      h[0] = uR[1] * x[2] - uR[2] * x[1];
      h[1] = uR[2] * x[0] - uR[0] * x[2];
      h[2] = uR[0] * x[1] - uR[1] * x[0];
      const int o = is_psr ? 3 : 6;
      h[o] = -uR[0]; h[o + 1] = -uR[1]; h[o + 2] = -uR[2];
      if (a.adjust_yof) h[9] = -(udR[0] * x[0] + udR[1] * x[1] + udR[2] * x[2]);
      h[10] = 1.0;
      const int gg = is_psr ? g : IGV_GNSS_FS;
      col_last = a.col_of_gnss[gg];
      idx_last = a.idx_gnss[gg];
      resv = -(is_psr ? a.res_pos[bi] : a.res_vel[bi]);
      sig = is_psr ? a.sig_psr[bi] : a.sig_dopp[bi];
      if (a.chi2_test) {  // Update.cpp:81-102 with dof = 1 on [SE23, YOF, (cb|fs)]
        int idx[11];
        for (int k = 0; k < 9; ++k) idx[k] = k;
        idx[9] = a.idx_yof; idx[10] = idx_last;
        double s = sig * sig;
        for (int p = 0; p < 11; ++p) {
          double acc = 0.0;
          for (int q = 0; q < 11; ++q) acc = fma(Pb[idx[p] + (size_t)idx[q] * a.ld], h[q], acc);
          s = fma(h[p], acc, s);
        }
        const double gam = resv * resv / s;
        if (!(a.chi2_n >= 1 && gam < a.chi2[0])) ok = false;
      }
    }
    if (ok) {
      for (int k = 0; k < 10; ++k) Hb[r + (size_t)k * a.ldh] = h[k];
      Hb[r + (size_t)col_last * a.ldh] = 1.0;
      rb[r] = resv;
      Rb[r] = sig * sig;
      atomicAdd(&s_cnt, 1);
    } else {
      rb[r] = 0.0;
      Rb[r] = 1.0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) a.cnt[b] = s_cnt;
}

// Rows of GnssUpdate::addNewTrackedSys for one new system (GnssUpdate.cpp:372-473): the satellites of that
// constellation (clock bias) or every satellite (clock drift FS), compacted to the front in satellite order; the
// remaining rows are zero padding. H_x columns: SE23 (9) then the yaw offset.  The yaw-offset column uses
// getRecef2enu() * dotRw2enu (:403, :464) where updateTrackedSys uses getRenu2ecef() (:166): kept as the reference has it.
__global__ void k_gnss_new_rows(const double* X, int xsize, IgvGnssNewRowsLaunch g) {
  const int b = blockIdx.x, S = g.S;
  const double* Xb = X + (size_t)b * xsize;
  double* Hx = g.Hx + (size_t)b * S * 10;
  double* Hf = g.Hf + (size_t)b * S;
  double* rs = g.res + (size_t)b * S;
  __shared__ int s_row[128];
  __shared__ int s_n;
  __shared__ double s_noise;
  const bool fs = (g.gtype == IGV_GNSS_FS);
  if (threadIdx.x == 0) {
    int n = 0;
    double acc = 0.0;
    for (int i = 0; i < S; ++i) {
      const size_t bi = (size_t)b * S + i;
      const bool mine = fs || g.sys[bi] == g.gtype;
      s_row[i] = mine ? n : -1;
      if (mine) {
        const double sg = fs ? g.sig_dopp[bi] : g.sig_psr[bi];
        acc += sg * sg;
        ++n;
      }
    }
    s_n = n;
    s_noise = n > 0 ? acc / n : 1.0;     // avg_noise^2 = mean of sigma_i^2 (the amplitude is inside sigma_i)
    g.count[b] = n;
    g.noise2[b] = s_noise;
  }
  for (int t = threadIdx.x; t < S * 10; t += blockDim.x) Hx[t] = 0.0;
  for (int t = threadIdx.x; t < S; t += blockDim.x) { Hf[t] = 0.0; rs[t] = 0.0; }
  __syncthreads();
  const double yof = Xb[33 + IGV_GNSS_YOF];
  double sy, cy;
  sincos(yof, &sy, &cy);
  const double* Re = g.R_enu2ecef + (size_t)b * 9;
  const double Rz[9] = {cy, -sy, 0.0, sy, cy, 0.0, 0.0, 0.0, 1.0};
  const double dRz[9] = {-sy, -cy, 0.0, cy, -sy, 0.0, 0.0, 0.0, 0.0};
  double Rw[9], dRq[9], Rq[9];
  mat3_mul(Re, Rz, Rw);
  if (g.R_ecef2enu) {
    for (int k = 0; k < 9; ++k) Rq[k] = g.R_ecef2enu[(size_t)b * 9 + k];
  } else {   // the aligner's R_ecef2enu is the transpose of its R_enu2ecef
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rq[3 * i + j] = Re[3 * j + i];
  }
  mat3_mul(Rq, dRz, dRq);
  const double* x = fs ? (Xb + 12) : (Xb + 9);   // v or p
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const int r = s_row[i];
    if (r < 0) continue;
    const size_t bi = (size_t)b * S + i;
    const double* u = g.unit + bi * 3;
    double uR[3], udR[3];
    mat3T_vec(Rw, u, uR);
    mat3T_vec(dRq, u, udR);
    Hx[r + 0 * S] = uR[1] * x[2] - uR[2] * x[1];
    Hx[r + 1 * S] = uR[2] * x[0] - uR[0] * x[2];
    Hx[r + 2 * S] = uR[0] * x[1] - uR[1] * x[0];
    const int o = fs ? 6 : 3;
    Hx[r + (o + 0) * S] = -uR[0]; Hx[r + (o + 1) * S] = -uR[1]; Hx[r + (o + 2) * S] = -uR[2];
    if (g.adjust_yof) Hx[r + 9 * S] = -(udR[0] * x[0] + udR[1] * x[1] + udR[2] * x[2]);
    Hf[r] = 1.0;
    rs[r] = -(fs ? g.res_vel[bi] : g.res_pos[bi]);
  }
}

}  // namespace

void igv_launch_gnss_new_rows(igv_batch* h, const IgvGnssNewRowsLaunch& g) {
  IgvProfScope prof_scope_(h, IGV_K_GNSS_ROWS);
  k_gnss_new_rows<<<h->B, 64, 0, h->stream>>>(h->Xc(), h->xsize, g);
  h->launches++;
}

void igv_launch_gnss_rows(igv_batch* h, const IgvGnssLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_GNSS_ROWS);
  GnssArgs a;
  a.P = h->Pc(); a.ld = h->ld; a.X = h->Xc(); a.xsize = h->xsize;
  a.S = l.S; a.unit = l.unit; a.res_pos = l.res_pos; a.res_vel = l.res_vel; a.sig_psr = l.sig_psr;
  a.sig_dopp = l.sig_dopp; a.sys = l.sys; a.Renu = l.R_enu2ecef;
  a.adjust_yof = l.adjust_yof; a.chi2_test = l.chi2_test;
  IgvLayout L = h->layout();
  a.idx_yof = L.idx_gnss[IGV_GNSS_YOF];
  for (int i = 0; i < 6; ++i) { a.idx_gnss[i] = L.idx_gnss[i]; a.col_of_gnss[i] = l.col_of_gnss[i]; }
  a.ncols = l.blk.n;
  a.H = h->Hg; a.ldh = 2 * h->cfg.max_sats; a.res = h->rg; a.Rd = h->Rg; a.cnt = h->cnt_g;
  a.chi2 = h->chi2; a.chi2_n = h->chi2_n;
  k_gnss_rows<<<h->B, 64, 0, h->stream>>>(a);
  h->launches++;
}
