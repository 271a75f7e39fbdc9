// Compression of the stacked MSCKF Jacobian through its Gram matrix ("Cholesky QR") on the FP64 tensor pipe.
//
// Reference call sites: RemoveLostUpdate.cpp:139-155, SwMargUpdate.cpp:161-176, KeyframeUpdate.cpp:557-572
// (Eigen::SPQR, natural ordering, then Q^T H and Q^T r, topRows). The update only consumes the n x (n+1)
// block [R | Q^T r]. With the stack W = [H r] (m x (n+1)),
//     G = W^T W = [ H^T H   H^T r ]          R^T R = H^T H   (R upper triangular = the QR factor up to row signs)
//                 [ r^T H   r^T r ]          R^T y = H^T r   (y = Q^T r on the rows of R)
// so one symmetric rank-m accumulation followed by an (n+1)-column Cholesky elimination of G yields the same
// block the Householder kernels (k_qr.cu) produce, with m n^2 instead of 2 m n^2 flops and no serial
// reflector chain: the accumulation is a pure stream of DMMA.8x8x4 (mma.sync.m8n8k4.f64).
//
// Numerics. The stack is rank deficient by construction (gauge directions), so G is only semi-definite: the
// elimination is the outer-product Cholesky in its LDL^T form with a relative pivot threshold; a pivot at
// rounding level marks a dependent column, whose row of R (and entry of y) is set to zero -- exactly what a QR
// returns there up to rounding. Because R enters the update only through R^T R and R^T y, the backward error
// is that of G itself, eps * ||H||^2, relative to which S = R P R^T + sigma^2 I is as well conditioned as with a
// Householder R (no squaring of a condition number happens: G is never inverted).
//
// k_gram_accum: grid (split, B). The CTA streams its share of the accepted rows through shared memory in
// chunks of KC rows (row stride = 4 mod 8 doubles: conflict-free fragment loads) and every warp owns whole
// 24 x 24 super-blocks (3 x 3 DMMA tiles) of the upper triangle of G: 6 fragment loads feed 9 DMMAs.
// k_gram_factor: grid B. Sums the partial Gram matrices in a fixed order (bitwise reproducible), eliminates,
// and writes [R | y] in the layout k_ekf_update consumes.
//
// Algorithmic bytes per sequence: one read of the stack 8 m (n+1) + one write of the result 8 n (n+1) (plus
// 2 x 8 (n+1)^2 for the Gram round trip between the two kernels).
#include <cstdlib>

#include "igv_device.cuh"
#include "k_gram_tc.cuh"

using namespace igv;

namespace {

constexpr int KC = 32;   // rows per staged chunk

struct GramArgs {
  const double* Hs; int F; int qmax; int ldo; const int* f_rows; int max_valid;
  int n;                      // columns of H; the stack has n + 1 columns
  int F_alloc; size_t hs_seq_stride;
  int hs_f32;                    // the stack holds floats (IGV_PREC_FP32_STACK)
  double* G; long g_seq_stride; int n1p;     // partial Gram matrices [b][part][n1p x n1p], row-major
  int lds;                    // shared-memory row stride of a chunk
  int nsb;                    // 24-column super-blocks per side
  int* n_acc;
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

template <int SPW>   // super-blocks per warp
__global__ void __launch_bounds__(SPW == 1 ? 256 : 288) k_gram_accum(GramArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int n1 = a.n + 1, lds = a.lds, nsb = a.nsb;
  double* tile = sm;                                              // [KC][lds]
  int* rowstart = reinterpret_cast<int*>(tile + KC * lds);        // [F_range + 1]
  __shared__ int s_total;
  const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
  const int nfr = f1 - f0;
  if (tid == 0) {
    // accepted tracks before f0 count towards the max_valid cap (RemoveLostUpdate.cpp:120-122)
    const int* fr = a.f_rows + (size_t)b * a.F_alloc;
    int acc = 0, rows = 0;
    for (int f = 0; f < f0; ++f) acc += (fr[f] > 0);
    for (int f = f0; f < f1; ++f) {
      rowstart[f - f0] = rows;
      const bool on = fr[f] > 0 && (a.max_valid <= 0 || acc < a.max_valid);
      if (fr[f] > 0) ++acc;
      if (on) rows += fr[f];
    }
    rowstart[nfr] = rows;
    s_total = rows;
    if (part == nparts - 1 && a.n_acc) a.n_acc[b] = (a.max_valid > 0) ? min(acc, a.max_valid) : acc;
  }
  for (int t = tid; t < KC * lds; t += blockDim.x) tile[t] = 0.0;   // pad columns stay zero for good
  __syncthreads();
  const int total = s_total;
  const double* src = a.Hs + (size_t)b * a.hs_seq_stride;

  // this warp's super-blocks: s = warp, warp + nw, ... in row-major order of the upper triangle
  int bi[SPW], bj[SPW];
  bool own[SPW];
  const int nsbt = nsb * (nsb + 1) / 2;
#pragma unroll
  for (int s = 0; s < SPW; ++s) {
    int idx = warp + s * nw;
    own[s] = idx < nsbt;
    if (!own[s]) idx = 0;
    int r = 0;
    while (idx >= nsb - r) { idx -= nsb - r; ++r; }
    bi[s] = r;
    bj[s] = r + idx;
  }
  double acc[SPW][3][3][2];
#pragma unroll
  for (int s = 0; s < SPW; ++s)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[s][i][j][0] = acc[s][i][j][1] = 0.0;

  const int fk = lane & 3, fc = lane >> 2;   // fragment element: row k0 + fk of the chunk, column 8 t + fc
  for (int base = 0; base < total; base += KC) {
    if (base > 0) __syncthreads();           // everybody is done with the previous chunk
    for (int r = warp; r < KC; r += nw) {
      const int v = base + r;
      double* trow = tile + r * lds;
      if (v < total) {
        int lo = 0, hi = nfr;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (rowstart[mid] <= v) lo = mid; else hi = mid;
        }
        const size_t off = ((size_t)(f0 + lo) * a.qmax + (v - rowstart[lo])) * a.ldo;
        if (a.hs_f32) {
          const float* srow = reinterpret_cast<const float*>(src) + off;
          for (int c = lane; c < n1; c += 32) trow[c] = (double)__ldg(srow + c);
        } else {
          const double* srow = src + off;
          for (int c = lane; c < n1; c += 32) trow[c] = __ldg(srow + c);
        }
      } else {
        for (int c = lane; c < n1; c += 32) trow[c] = 0.0;
      }
    }
    __syncthreads();
    const int kend = min(KC, (total - base + 3) & ~3);
    for (int k0 = 0; k0 < kend; k0 += 4) {
      const double* frow = tile + (k0 + fk) * lds + fc;
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        if (!own[s]) continue;
        double fa[3], fb[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) fa[t] = frow[8 * (3 * bi[s] + t)];
        if (bi[s] == bj[s]) {
#pragma unroll
          for (int t = 0; t < 3; ++t) fb[t] = fa[t];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) dmma(acc[s][i][j][0], acc[s][i][j][1], fa[i], fb[j]);
        } else {
#pragma unroll
          for (int t = 0; t < 3; ++t) fb[t] = frow[8 * (3 * bj[s] + t)];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) dmma(acc[s][i][j][0], acc[s][i][j][1], fa[i], fb[j]);
        }
      }
    }
  }
  // accumulator fragment: thread holds G[8 ti + lane/4][8 tj + 2 (lane%4) + {0,1}]
  double* G = a.G + (size_t)b * a.g_seq_stride + (size_t)part * a.n1p * a.n1p;
#pragma unroll
  for (int s = 0; s < SPW; ++s) {
    if (!own[s]) continue;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (bi[s] == bj[s] && j < i) continue;
        const int row = 8 * (3 * bi[s] + i) + fc, col = 8 * (3 * bj[s] + j) + 2 * fk;
        *reinterpret_cast<double2*>(G + (size_t)row * a.n1p + col) = make_double2(acc[s][i][j][0], acc[s][i][j][1]);
      }
  }
}

// ---- narrow stacks (n + 1 <= 72 columns): every warp owns the WHOLE upper triangle of G --------------------
// The rows are dealt to the CTA's four warps in chunks of 8; each warp streams its chunks through a private
// double-buffered staging area with cp.async (no block barrier in the main loop), keeps all NT (NT+1)/2
// accumulator tiles in registers (NT fragment loads feed NT (NT+1)/2 DMMAs) and the four partial sums are
// folded through shared memory in a fixed order at the end.
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int MINB>
__global__ void __launch_bounds__(128, MINB) k_gram_stream(GramArgs a) {
  constexpr int NTT = NT * (NT + 1) / 2, LDS = 8 * NT + 4, RC = 8, NWARP = 4;
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n1 = a.n + 1;
  constexpr int STAGE = NWARP * 2 * RC * LDS, FOLD = 2 * NTT * 64;
  double* stage = sm + warp * (2 * RC * LDS);
  int* rowstart = reinterpret_cast<int*>(sm + (STAGE > FOLD ? STAGE : FOLD));
  __shared__ int s_total;
  const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
  const int nfr = f1 - f0;
  if (tid == 0) {
    const int* fr = a.f_rows + (size_t)b * a.F_alloc;
    int acc = 0, rows = 0;
    for (int f = 0; f < f0; ++f) acc += (fr[f] > 0);
    for (int f = f0; f < f1; ++f) {
      rowstart[f - f0] = rows;
      const bool on = fr[f] > 0 && (a.max_valid <= 0 || acc < a.max_valid);
      if (fr[f] > 0) ++acc;
      if (on) rows += fr[f];
    }
    rowstart[nfr] = rows;
    s_total = rows;
    if (part == nparts - 1 && a.n_acc) a.n_acc[b] = (a.max_valid > 0) ? min(acc, a.max_valid) : acc;
  }
  for (int t = lane; t < 2 * RC * LDS; t += 32) stage[t] = 0.0;   // pad columns stay zero
  __syncthreads();
  const int total = s_total;
  const double* src = a.Hs + (size_t)b * a.hs_seq_stride;
  const int nchunks = (total + RC - 1) / RC;

  auto issue = [&](int c, int buf) {
    // lane r (< 8) resolves row 8 c + r of the accepted stack
    long so = -1;
    {
      const int v = c * RC + (lane & 7);
      if (v < total) {
        int lo = 0, hi = nfr;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (rowstart[mid] <= v) lo = mid; else hi = mid;
        }
        so = ((long)(f0 + lo) * a.qmax + (v - rowstart[lo])) * a.ldo;
      }
    }
    double* dst = stage + buf * (RC * LDS);
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      const long o = __shfl_sync(0xffffffffu, so, r);
      if (o >= 0) {
        for (int c2 = lane; c2 < n1; c2 += 32) cp_async8(dst + r * LDS + c2, src + o + c2);
      } else {
        for (int c2 = lane; c2 < n1; c2 += 32) dst[r * LDS + c2] = 0.0;
      }
    }
    cp_async_commit();
  };

  double acc[NTT][2];
#pragma unroll
  for (int t = 0; t < NTT; ++t) acc[t][0] = acc[t][1] = 0.0;
  const int fk = lane & 3, fc = lane >> 2;
  int buf = 0;
  if (warp < nchunks) issue(warp, 0);
  for (int c = warp; c < nchunks; c += NWARP) {
    const bool more = c + NWARP < nchunks;
    if (more) issue(c + NWARP, buf ^ 1);
    if (more) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncwarp();
    const double* base = stage + buf * (RC * LDS) + fk * LDS + fc;
#pragma unroll
    for (int ks = 0; ks < RC / 4; ++ks) {
      double f[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) f[t] = base[ks * 4 * LDS + 8 * t];
      int idx = 0;
#pragma unroll
      for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = i; j < NT; ++j, ++idx) dmma(acc[idx][0], acc[idx][1], f[i], f[j]);
    }
    __syncwarp();
    buf ^= 1;
  }
  // ---- fold the four partial sums: (0 += 1, 2 += 3), then 0 += 2, fixed order ------------------------------
  __syncthreads();
  double2* fold = reinterpret_cast<double2*>(sm);
  if (warp & 1) {
    double2* dst = fold + (warp >> 1) * (NTT * 32);
#pragma unroll
    for (int t = 0; t < NTT; ++t) dst[t * 32 + lane] = make_double2(acc[t][0], acc[t][1]);
  }
  __syncthreads();
  if (!(warp & 1)) {
    const double2* s2 = fold + (warp >> 1) * (NTT * 32);
#pragma unroll
    for (int t = 0; t < NTT; ++t) { const double2 v = s2[t * 32 + lane]; acc[t][0] += v.x; acc[t][1] += v.y; }
  }
  __syncthreads();
  if (warp == 2) {
#pragma unroll
    for (int t = 0; t < NTT; ++t) fold[t * 32 + lane] = make_double2(acc[t][0], acc[t][1]);
  }
  __syncthreads();
  if (warp == 0) {
    double* G = a.G + (size_t)b * a.g_seq_stride + (size_t)part * a.n1p * a.n1p;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int j = i; j < NT; ++j, ++idx) {
        const double2 v = fold[idx * 32 + lane];
        const int row = 8 * i + fc, col = 8 * j + 2 * fk;
        *reinterpret_cast<double2*>(G + (size_t)row * a.n1p + col) = make_double2(acc[idx][0] + v.x, acc[idx][1] + v.y);
      }
  }
}

template <int NT, int MINB>
void launch_stream_gram(const GramArgs& a, int split, int B, int frange, cudaStream_t st) {
  constexpr int NTT = NT * (NT + 1) / 2, LDS = 8 * NT + 4;
  constexpr size_t STAGE = 4 * 2 * 8 * LDS, FOLD = 2 * NTT * 64;
  const size_t smem = sizeof(double) * (STAGE > FOLD ? STAGE : FOLD) + sizeof(int) * (frange + 2);
  IGV_SMEM_OPTIN((k_gram_stream<NT, MINB>), 200 * 1024);
  dim3 grid(split, B);
  k_gram_stream<NT, MINB><<<grid, 128, smem, st>>>(a);
}

struct FactorArgs {
  const double* G; long g_seq_stride; int n1p; int nparts;
  int n;
  double* out; long out_stride;    // n x (n+1) row-major [R | y]
  double tol;                      // relative pivot threshold
  int* flags;                      // [B] IGV_FLAG_WEAK_PIVOT
};

__global__ void __launch_bounds__(256) k_gram_factor(FactorArgs a) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int n = a.n, n1 = n + 1;
  const int npk = n * (n + 3) / 2;
  double* U = sm;                 // packed upper triangle, rows 0..n-1, columns i..n (column n = right-hand side)
  double* dref = U + npk;         // [n] original diagonal = ||column j of H||^2
  double* dsc = dref + n;         // [n] row scale 1/sqrt(pivot) (0: dependent column)
  auto off = [&](int i) { return i * n1 - (i * (i - 1)) / 2 - i; };   // U[off(i) + k] = entry (i, k), k >= i
  const double* Gb = a.G + (size_t)b * a.g_seq_stride;
  const size_t pstride = (size_t)a.n1p * a.n1p;
  for (int i = warp; i < n; i += nw) {
    const int oi = off(i);
    for (int k = i + lane; k < n1; k += 32) {
      double s = 0.0;
      for (int p = 0; p < a.nparts; ++p) s += Gb[p * pstride + (size_t)i * a.n1p + k];   // fixed order
      U[oi + k] = s;
      if (k == i) dref[i] = s;
    }
  }
  __syncthreads();
  // outer-product elimination; rows stay unscaled (U = D R), the update uses 1/pivot, so there is one barrier
  // per column and the square roots leave the serial chain
  for (int j = 0; j < n; ++j) {
    const int oj = off(j);
    const double d = U[oj + j];
    const bool live = d > a.tol * dref[j] && d > 1e-280;   // (the second test keeps the fast reciprocal in range)
    if (live && tid == 0 && d < kWeakPivot * dref[j]) atomicOr(&a.flags[b], IGV_FLAG_WEAK_PIVOT);
    if (live) {
      const double inv = rcp_nobranch(d);   // d > 0 and normal here; every thread needs it, so no slow-path division
      for (int i = j + 1 + warp; i < n; i += nw) {
        const int oi = off(i);
        const double lij = U[oj + i] * inv;
        for (int k = i + lane; k < n1; k += 32) U[oi + k] = fma(-lij, U[oj + k], U[oi + k]);
      }
    }
    if (tid == 0) dsc[j] = live ? rsqrt_nobranch(d) : 0.0;
    __syncthreads();
  }
  double* out = a.out + (size_t)b * a.out_stride;
  for (int i = warp; i < n; i += nw) {
    const double sc = dsc[i];
    const int oi = off(i);
    for (int k = lane; k < n1; k += 32) out[(size_t)i * n1 + k] = (k >= i) ? U[oi + k] * sc : 0.0;
  }
}

// Blocked version for n <= 160: G (lower, column-major) and the right-hand side sit in shared memory and the
// factorisation is cta_chol_solve_fused<4> in its semi-definite mode -- 4-column panels, the trailing update on DMMA,
// the right-hand side riding along -- instead of one barrier and one rank-1 update per column.
template <int NB>
__global__ void __launch_bounds__(256) k_gram_factor_blocked(FactorArgs a) {
  extern __shared__ double sm[];
  __shared__ int s_ok;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int n = a.n, n1 = n + 1;
  const int lds = n + ((4 - n % 8) + 8) % 8;     // 4 (mod 8): conflict-free DMMA fragments
  double* S = sm;                 // [n][lds] lower triangle, column-major: S[i + j*lds] = G[j][i], i >= j
  double* Zr = S + (size_t)n * lds;   // [lds] right-hand side H^T r
  double* dref = Zr + lds;        // [n] original diagonal
  const double* Gb = a.G + (size_t)b * a.g_seq_stride;
  const size_t pstride = (size_t)a.n1p * a.n1p;
  for (int j = warp; j < n; j += nw)
    for (int i = j + lane; i < n1; i += 32) {
      double v = 0.0;
      for (int p = 0; p < a.nparts; ++p) v += Gb[p * pstride + (size_t)j * a.n1p + i];   // fixed order
      if (i < n) S[i + (size_t)j * lds] = v; else Zr[j] = v;
      if (i == j) dref[j] = v;
    }
  __syncthreads();
  // panel width: with ONE right-hand side the serial diagonal block dominates; measured at c2: 4-column panels 0.137 -> 0.117 ms
  // when the batch fills the chip, but 71 -> 81 us for a single sequence, where 8 stays
  cta_chol_solve_fused<NB>(S, n, lds, Zr, lds, 0, 1, &s_ok, dref, a.tol);
  if (tid == 0 && (s_ok & 2)) atomicOr(&a.flags[b], IGV_FLAG_WEAK_PIVOT);
  double* out = a.out + (size_t)b * a.out_stride;
  for (int i = warp; i < n; i += nw)
    for (int k = lane; k < n1; k += 32)
      out[(size_t)i * n1 + k] = (k < i) ? 0.0 : ((k < n) ? S[k + (size_t)i * lds] : Zr[i]);
}

template <int SPW>
void launch_accum(const GramArgs& a, int split, int B, int warps, size_t smem, cudaStream_t st) {
  IGV_SMEM_OPTIN((k_gram_accum<SPW>), 200 * 1024);
  dim3 grid(split, B);
  k_gram_accum<SPW><<<grid, warps * 32, smem, st>>>(a);
}

}  // namespace

// row stride of a partial Gram matrix: whole 24-column super-blocks, plus 8 so that the widest tile row of the
// stream kernel's template (8 NT columns) always fits
int igv_gram_n1p(int ncols_max) { return 24 * ((ncols_max + 1 + 23) / 24) + 8; }

bool igv_gram_supported(int n) { return (n + 1 + 23) / 24 <= 9; }   // up to 216 columns (SW <= 35)

void igv_launch_gram_compress(igv_batch* h, int F, int max_valid, int split) {
  igv_commit_copies(h);
  IgvLayout L = h->layout();
  const int n = 6 * L.n_clones;
  GramArgs a;
  a.Hs = h->Hs; a.F = F; a.qmax = h->qmax; a.ldo = n + 1; a.f_rows = h->f_rows; a.max_valid = max_valid;
  a.n = n; a.F_alloc = h->cfg.max_feats; a.hs_f32 = h->stack_f32;
  a.hs_seq_stride = (size_t)h->cfg.max_feats * h->qmax * (h->ncols_max + 1);
  a.nsb = (n + 1 + 23) / 24;
  a.n1p = 24 * a.nsb + 8;
  a.G = h->Gws; a.g_seq_stride = (long)h->qr_split_cap * h->gram_n1p * h->gram_n1p;
  a.lds = a.n1p + 4;             // = 4 (mod 8): the 4 rows x 8 columns of a fragment hit 32 distinct banks pairs
  a.n_acc = h->n_acc;
  const int nsbt = a.nsb * (a.nsb + 1) / 2;
  const int spw = (nsbt <= 8) ? 1 : ((nsbt <= 16) ? 2 : ((nsbt <= 24) ? 3 : ((nsbt <= 36) ? 4 : 5)));
  const int warps = (nsbt + spw - 1) / spw;
  const size_t smem = sizeof(double) * (size_t)KC * a.lds + sizeof(int) * ((F + split - 1) / split + 2);
  const int nt = (n + 1 + 7) / 8;
  const int frange = (F + split - 1) / split;
  const int gram_cfg = h->knobs.gram_cfg;            // test knob: 1 forces the super-block kernel
  h->last_gram_tc = 0;
#ifndef IGV_EMULATE
  // (narrow stacks stay on the FP64 accumulation of the float stack: at n + 1 <= 128 the DMMA kernels are as fast)
  if (h->gram_tc && h->stack_f32 && n + 1 <= 192 && (n + 1 > 128 || h->knobs.gram_cfg == 3)) {
    // IGV_PREC_TF32_GRAM: the float stack through tcgen05 (k_gram_tc.cuh); one CTA per (part, sequence), one CTA per SM
    igv_tc::GramTcArgs t;
    t.Hs = reinterpret_cast<const float*>(h->Hs); t.hs_seq_stride = 2 * a.hs_seq_stride;   // floats (the buffer is sized in doubles)
    t.F = F; t.F_alloc = a.F_alloc; t.qmax = a.qmax; t.ldo = a.ldo; t.f_rows = a.f_rows; t.max_valid = max_valid;
    t.n1 = n + 1; t.NC = (n + 1 + 31) / 32;
    t.G = a.G; t.g_seq_stride = a.g_seq_stride; t.n1p = a.n1p; t.n_acc = a.n_acc; t.drain_stages = h->knobs.tc_drain;
    // one CTA per SM (it owns the whole TMEM): small batches get as many parts as fill the chip in ONE wave, batches that
    // fill it anyway two (measured at c5: B = 148, 1 / 2 parts: 2.02 / 1.68 ms incl. the factorisation; B = 32, 4 / 10 parts:
    // 0.93 / 0.95 ms)
    split = max(1, min(min(split, h->B >= 148 ? 2 : 148 / max(1, h->B)), F));
    const size_t tsmem = igv_tc::gram_tc_smem_bytes(t.NC, (F + split - 1) / split);
    IGV_SMEM_OPTIN((igv_tc::k_gram_tc), 226 * 1024);
    dim3 tgrid(split, h->B);
    igv_tc::k_gram_tc<<<tgrid, igv_tc::kThreads, tsmem, h->stream>>>(t);
    h->last_gram_tc = 1;
    igv_launch_gram_factor(h, split);
    h->launches += 1;
    return;
  }
#endif
  const bool wide = gram_cfg == 1 || nt > 9 || h->stack_f32;   // the stream kernel copies doubles asynchronously
  if (!wide) {
    if (nt <= 4) launch_stream_gram<4, 6>(a, split, h->B, frange, h->stream);
    else if (nt <= 6) launch_stream_gram<6, 4>(a, split, h->B, frange, h->stream);
    else if (gram_cfg == 2) launch_stream_gram<9, 3>(a, split, h->B, frange, h->stream);   // A/B: 3 CTAs/SM, spills
    else launch_stream_gram<9, 2>(a, split, h->B, frange, h->stream);
  } else
  switch (spw) {
    case 1: launch_accum<1>(a, split, h->B, warps, smem, h->stream); break;
    case 2: launch_accum<2>(a, split, h->B, warps, smem, h->stream); break;
    case 3: launch_accum<3>(a, split, h->B, warps, smem, h->stream); break;
    case 4: launch_accum<4>(a, split, h->B, warps, smem, h->stream); break;
    default: launch_accum<5>(a, split, h->B, warps, smem, h->stream); break;
  }
  igv_launch_gram_factor(h, split);
  h->launches += 1;
}

void igv_launch_gram_factor(igv_batch* h, int nparts) {
  igv_commit_copies(h);
  IgvLayout L = h->layout();
  const int n = 6 * L.n_clones;
  FactorArgs f;
  f.G = h->Gws; f.g_seq_stride = (long)h->qr_split_cap * h->gram_n1p * h->gram_n1p;
  f.n1p = 24 * ((n + 1 + 23) / 24) + 8; f.nparts = nparts;
  f.n = n; f.out = h->Hc; f.out_stride = (long)h->ncols_max * (h->ncols_max + 1);
  // (measured on c5: thresholds between 1e-13 and 1e-5 give identical results behind the tensor-core Gram matrix, 1e-4 and
  // above start dropping observable directions; the knob stays for A/B runs)
  f.tol = h->last_gram_tc ? h->knobs.tc_pivot_tol : 1e-13; f.flags = h->flags;
  IGV_SMEM_OPTIN((k_gram_factor), 220 * 1024);
  IGV_SMEM_OPTIN((k_gram_factor_blocked<4>), 220 * 1024);
  IGV_SMEM_OPTIN((k_gram_factor_blocked<8>), 220 * 1024);
  const int factor_cfg = h->knobs.factor_cfg;     // test knob: 1 forces the column-by-column kernel
  const int lds = n + ((4 - n % 8) + 8) % 8;
  const size_t bsmem = sizeof(double) * ((size_t)n * lds + lds + n);
  if (bsmem <= 200 * 1024 && factor_cfg != 1) {
    if (h->B >= 296) k_gram_factor_blocked<4><<<h->B, 256, bsmem, h->stream>>>(f);
    else k_gram_factor_blocked<8><<<h->B, 256, bsmem, h->stream>>>(f);
  } else {
    const size_t fsmem = sizeof(double) * ((size_t)n * (n + 3) / 2 + 2 * n);
    k_gram_factor<<<h->B, 256, fsmem, h->stream>>>(f);
  }
  h->launches += 1;
}
