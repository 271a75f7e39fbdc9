// QR compression of the stacked MSCKF Jacobian: [R | Q^T r] of the m x (n+1) stack, keeping n rows.
//
// Reference call sites: RemoveLostUpdate.cpp:139-155, SwMargUpdate.cpp:161-176,
// KeyframeUpdate.cpp:557-572 -- `H.sparseView()` -> Eigen::SPQR (SuiteSparse, natural ordering) ->
// dense Q^T H / Q^T r -> topRows.  Only Q^T[H r] is consumed, so any Householder QR is equivalent.
//
// Kernel: a running upper-triangular factor R (packed, in shared memory) is updated with successive
// row chunks of the stack ("triangle-on-top-of-rectangle" Householder, as LAPACK tpqrt). A chunk of
// W*ROWS rows lives entirely in REGISTERS: thread (warp w, lane l) holds rows [w*ROWS,(w+1)*ROWS) of
// columns {l, l+32, l+64, ...}. A reflector for column j therefore needs no cross-lane reduction for
// the rank-1 update: the owner lane publishes v through SMEM, every thread forms the partial dot
// products of its own columns over its own rows, partials are combined across warps through SMEM.
// Flops are the minimum 2 m n^2 (no TSQR tree inflation) and HBM traffic is one coalesced read of
// the stack + one write of the n x (n+1) result: 8 m (n+1) + 8 n (n+1) bytes per sequence.
// For few sequences the rows are split over `split` CTAs and a second launch folds the partial
// triangles (TSQR with a flat tree).
#include <cstdlib>

#include "igv_device.cuh"

using namespace igv;

namespace {

struct QrArgs {
  // source 0: feature blocks
  const double* Hs; int F; int qmax; int ldo; const int* f_rows; int max_valid;
  // source 1: dense rows (partial triangles)
  const double* dense; long dense_stride; int dense_rows;
  int src_mode;
  int n;                    // columns to eliminate; ncols1 = n + 1 (last column = residual)
  double* out; long out_stride;   // [b][part] n x (n+1) row-major
  int* n_acc;               // accepted-feature count per sequence (written by part 0, source 0)
  int F_alloc;              // leading dimension of f_rows per sequence
  size_t hs_seq_stride;     // elements per sequence in Hs
};

template <int NSLOT, int ROWS, int W>
__global__ void __launch_bounds__(W * 32) k_qr_compress(QrArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, nc1 = n + 1, ldo = a.ldo;
  const int npk = n * (n + 3) / 2;              // packed upper rows 0..n-1, cols j..n
  double* Rp = sm;                              // [npk]
  double* vbuf = Rp + npk;                      // [W*ROWS]
  double* s_part = vbuf + W * ROWS;             // [W][32*NSLOT]
  double* s_norm = s_part + W * 32 * NSLOT;     // [2][W]
  int* rowstart = reinterpret_cast<int*>(s_norm + 2 * W + 2);  // [F_range + 1]
  __shared__ int s_total, s_f0;
  for (int t = tid; t < npk; t += blockDim.x) Rp[t] = 0.0;
  // ---- row enumeration -------------------------------------------------------------------------
  const double* src_base;
  if (a.src_mode == 0) {
    const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
    if (tid == 0) {
      // accepted features before f0 count towards the max_valid cap (RemoveLostUpdate.cpp:120-122)
      const int* fr = a.f_rows + (size_t)b * a.F_alloc;
      int acc = 0;
      for (int f = 0; f < f0; ++f) acc += (fr[f] > 0);
      int rows = 0;
      for (int f = f0; f < f1; ++f) {
        rowstart[f - f0] = rows;
        const bool on = fr[f] > 0 && (a.max_valid <= 0 || acc < a.max_valid);
        if (fr[f] > 0) ++acc;
        if (on) rows += fr[f];
      }
      rowstart[f1 - f0] = rows;
      s_total = rows;
      s_f0 = f0;
      if (part == nparts - 1 && a.n_acc) a.n_acc[b] = (a.max_valid > 0) ? min(acc, a.max_valid) : acc;
    }
    src_base = a.Hs + (size_t)b * a.hs_seq_stride;
  } else {
    if (tid == 0) { s_total = a.dense_rows; s_f0 = 0; }
    src_base = a.dense + (size_t)b * a.dense_stride;
  }
  __syncthreads();
  const int total = s_total;
  const int nfr = (a.src_mode == 0) ? ((int)((long)a.F * (part + 1) / nparts) - s_f0) : 0;

  auto off = [&](int j) { return j * nc1 - (j * (j - 1)) / 2; };  // packed offset of R[j][j]

  double tile[ROWS][NSLOT];
  for (int base = 0; base < total; base += W * ROWS) {
    // ---- load chunk: lane r < ROWS resolves the physical row of virtual row base + warp*ROWS + r ----
    long phys = -1;
    if (lane < ROWS) {
      const int v = base + warp * ROWS + lane;
      if (v < total) {
        if (a.src_mode == 0) {
          int lo = 0, hi = nfr;  // largest f with rowstart[f] <= v
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (rowstart[mid] <= v) lo = mid; else hi = mid;
          }
          phys = ((long)(s_f0 + lo) * a.qmax + (v - rowstart[lo])) * ldo;
        } else {
          phys = (long)v * ldo;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long o = __shfl_sync(0xffffffffu, phys, r);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int c = lane + 32 * s;
        tile[r][s] = (o >= 0 && c < nc1) ? __ldg(src_base + o + c) : 0.0;
      }
    }
    // ---- eliminate columns 0..n-1 of the chunk against R -----------------------------------------
#pragma unroll
    for (int sj = 0; sj < NSLOT; ++sj) {
      for (int lj = 0; lj < 32; ++lj) {
        const int j = sj * 32 + lj;
        if (j >= n) break;
        double* nb = s_norm + (j & 1) * W;
        if (lane == lj) {
          double ss = 0.0;
#pragma unroll
          for (int r = 0; r < ROWS; ++r) ss = fma(tile[r][sj], tile[r][sj], ss);
          nb[warp] = ss;
        }
        __syncthreads();  // A
        double sigma = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) sigma += nb[w];
        if (sigma == 0.0) continue;  // nothing below the diagonal in this chunk (uniform branch)
        const double alpha = Rp[off(j)];
        const double beta = -copysign(sqrt(fma(alpha, alpha, sigma)), alpha);
        const double tau = (beta - alpha) / beta;
        const double scale = 1.0 / (alpha - beta);
        if (lane == lj) {
#pragma unroll
          for (int r = 0; r < ROWS; ++r) vbuf[warp * ROWS + r] = tile[r][sj] * scale;
        }
        double rjk[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          rjk[s] = (k > j && k < nc1) ? Rp[off(j) + (k - j)] : 0.0;
        }
        __syncthreads();  // B
        const double* vw = vbuf + warp * ROWS;
        double d[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) d[s] = 0.0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const double v = vw[r];
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s >= sj) d[s] = fma(v, tile[r][s], d[s]);
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          if (s >= sj && k > j && k < nc1) s_part[warp * 32 * NSLOT + k] = d[s];
        }
        __syncthreads();  // C
        double wk[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          wk[s] = 0.0;
          if (s >= sj && k > j && k < nc1) {
            double dot = rjk[s];
#pragma unroll
            for (int w = 0; w < W; ++w) dot += s_part[w * 32 * NSLOT + k];
            wk[s] = tau * dot;
            if (warp == 0) Rp[off(j) + (k - j)] = rjk[s] - wk[s];
          }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const double v = vw[r];
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s >= sj) tile[r][s] = fma(-wk[s], v, tile[r][s]);
        }
        if (warp == 0 && lane == lj) Rp[off(j)] = beta;
      }
    }
    __syncthreads();
  }
  // ---- write [R | Q^T r] as n x (n+1) row-major, zeros below the diagonal -------------------------
  double* out = a.out + (size_t)b * a.out_stride + (size_t)part * n * nc1;
  for (int t = tid; t < n * nc1; t += blockDim.x) {
    const int j = t / nc1, k = t % nc1;
    out[t] = (k >= j) ? Rp[off(j) + (k - j)] : 0.0;
  }
}

template <int NSLOT, int ROWS, int W>
void launch_one(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st) {
  const int n = a.n;
  size_t smem = sizeof(double) * ((size_t)n * (n + 3) / 2 + W * ROWS + (size_t)W * 32 * NSLOT + 2 * W + 2) +
                sizeof(int) * (max_frange + 2);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_qr_compress<NSLOT, ROWS, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr_set = true;
  }
  dim3 grid(split, B);
  k_qr_compress<NSLOT, ROWS, W><<<grid, W * 32, smem, st>>>(a);
}

void launch_qr(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st) {
  const int nslot = (a.n + 1 + 31) / 32;
  if (nslot <= 1) launch_one<1, 32, 4>(a, split, B, max_frange, st);
  else if (nslot <= 2) launch_one<2, 32, 4>(a, split, B, max_frange, st);
  else if (nslot <= 3) launch_one<3, 32, 4>(a, split, B, max_frange, st);
  else if (nslot <= 4) launch_one<4, 24, 4>(a, split, B, max_frange, st);
  else if (nslot <= 6) launch_one<6, 16, 4>(a, split, B, max_frange, st);
  else if (nslot <= 8) launch_one<8, 12, 4>(a, split, B, max_frange, st);
  else launch_one<13, 7, 4>(a, split, B, max_frange, st);
}

}  // namespace

void igv_launch_qr_compress(igv_batch* h, int F, int max_valid) {
  IgvProfScope prof_scope_(h, IGV_K_QR);
  IgvLayout L = h->layout();
  const int n = 6 * L.n_clones;
  // enough CTAs to cover the chip: split the rows of each sequence when the batch is small
  int split = 1;
  const int target = 2 * 148;
  if (h->B < target) split = min(h->qr_split_cap, max(1, min((target + h->B - 1) / h->B, (F + 7) / 8)));
  if (const char* env = getenv("IGV_QR_SPLIT")) {  // test knob: force the row split
    const int s = atoi(env);
    if (s >= 1) split = min(h->qr_split_cap, min(s, max(1, F)));
  }
  QrArgs a;
  a.Hs = h->Hs; a.F = F; a.qmax = h->qmax; a.ldo = n + 1; a.f_rows = h->f_rows; a.max_valid = max_valid;
  a.dense = nullptr; a.dense_stride = 0; a.dense_rows = 0; a.src_mode = 0;
  a.n = n;
  a.n_acc = h->n_acc;
  a.F_alloc = h->cfg.max_feats;
  a.hs_seq_stride = (size_t)h->cfg.max_feats * h->qmax * (h->ncols_max + 1);
  if (split == 1) {
    a.out = h->Hc; a.out_stride = (long)h->ncols_max * (h->ncols_max + 1);
    launch_qr(a, 1, h->B, F, h->stream);
    h->launches++;
  } else {
    a.out = h->Rpart; a.out_stride = (long)h->qr_split_cap * h->ncols_max * (h->ncols_max + 1);
    launch_qr(a, split, h->B, (F + split - 1) / split + 1, h->stream);
    QrArgs c = a;
    c.src_mode = 1; c.dense = h->Rpart; c.dense_stride = a.out_stride; c.dense_rows = split * n;
    c.n_acc = nullptr;
    c.out = h->Hc; c.out_stride = (long)h->ncols_max * (h->ncols_max + 1);
    launch_qr(c, 1, h->B, 1, h->stream);
    h->launches += 2;
  }
}
