// QR compression of the stacked MSCKF Jacobian: [R | Q^T r] of the m x (n+1) stack, keeping n rows.
//
// Reference call sites: RemoveLostUpdate.cpp:139-155, SwMargUpdate.cpp:161-176,
// KeyframeUpdate.cpp:557-572 -- `H.sparseView()` -> Eigen::SPQR (SuiteSparse, natural ordering) ->
// dense Q^T H / Q^T r -> topRows.  Only Q^T[H r] is consumed, so any Householder QR is equivalent.
//
// Kernel: a running upper-triangular factor R (packed, in shared memory) is updated with successive
// row chunks of the stack ("triangle-on-top-of-rectangle" Householder, as LAPACK tpqrt). A chunk of
// W*ROWS rows lives entirely in REGISTERS: thread (warp w, lane l) holds rows [w*ROWS,(w+1)*ROWS) of
// NSLOT columns. A reflector for column j needs no cross-lane reduction for the rank-1 update: the
// owner lane publishes its raw column through SMEM, every thread forms the partial dot products of its
// own columns over its own rows, partials are combined across warps through SMEM (2 barriers/column).
//
// Column -> (slot, lane) map: columns are dealt to slots from the TOP slot down, in elimination
// order, so that the slot holding the columns eliminated first is the partially filled one and the
// residual column (never eliminated, updated by every reflector) shares slot 0 with the columns
// eliminated last. At step j only slots <= slot(j) are live; for n+1 = 67 columns this cuts the
// executed slot-steps from 162 (cyclic map) to 104 (ideal 69).
//
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
  int* n_acc;               // accepted-feature count per sequence (written by the last part, source 0)
  int F_alloc;              // leading dimension of f_rows per sequence
  size_t hs_seq_stride;     // elements per sequence in Hs
};

template <int NSLOT, int ROWS, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) k_qr_compress(QrArgs a) {
  static_assert(ROWS % 2 == 0 && ROWS <= 32, "ROWS must be even and <= 32");
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, nc1 = n + 1, ldo = a.ldo;
  const int r0 = nc1 - 32 * (NSLOT - 1);        // columns in the top slot (1..32)
  const int npk = n * (n + 3) / 2;              // packed upper rows 0..n-1, cols j..n
  double* Rp = sm;                              // [npk]
  double* vbuf = Rp + npk + (npk & 1);          // [2][W*ROWS], 16-byte aligned
  double* s_part = vbuf + 2 * W * ROWS;         // [W][NSLOT][32]
  int* rowstart = reinterpret_cast<int*>(s_part + W * 32 * NSLOT);  // [F_range + 1]
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
  // natural column held by (slot s, lane) ; -1 if none
  int mycol[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    if (s == NSLOT - 1) mycol[s] = (lane < r0) ? lane : -1;
    else mycol[s] = r0 + 32 * (NSLOT - 2 - s) + lane;
  }

  double tile[ROWS][NSLOT];
  unsigned it = 0;  // reflector counter (parity selects the vbuf half)
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
      for (int s = 0; s < NSLOT; ++s)
        tile[r][s] = (o >= 0 && mycol[s] >= 0) ? __ldg(src_base + o + mycol[s]) : 0.0;
    }
    // ---- eliminate columns 0..n-1 of the chunk against R, top slot first ---------------------------
#pragma unroll
    for (int sj = NSLOT - 1; sj >= 0; --sj) {
      const int nl = (sj == NSLOT - 1) ? r0 : 32;
      for (int lj = 0; lj < nl; ++lj) {
        const int j = (sj == NSLOT - 1) ? lj : r0 + 32 * (NSLOT - 2 - sj) + lj;
        if (j >= n) break;  // the residual column is never eliminated
        double* vb = vbuf + (it & 1u) * (W * ROWS);
        ++it;
        double sigma;
        if constexpr (W == 1) {
          // single-warp stream: the owner lane holds the whole column of this chunk -> local norm, no barrier
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          if (lane == lj) {
#pragma unroll
            for (int r = 0; r < ROWS; r += 2) {
              *reinterpret_cast<double2*>(vb + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
              if (r % 4 == 0) { s0 = fma(tile[r][sj], tile[r][sj], s0); s1 = fma(tile[r + 1][sj], tile[r + 1][sj], s1); }
              else { s2 = fma(tile[r][sj], tile[r][sj], s2); s3 = fma(tile[r + 1][sj], tile[r + 1][sj], s3); }
            }
          }
          sigma = __shfl_sync(0xffffffffu, (s0 + s1) + (s2 + s3), lj);
          __syncwarp();
        } else {
          if (lane == lj) {
#pragma unroll
            for (int r = 0; r < ROWS; r += 2)
              *reinterpret_cast<double2*>(vb + warp * ROWS + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
          }
          __syncthreads();  // A: raw column visible
          double ss = 0.0;
          for (int i = lane; i < W * ROWS; i += 32) ss = fma(vb[i], vb[i], ss);
          sigma = warp_sum(ss);
        }
        if (sigma == 0.0) continue;  // nothing below the diagonal in this chunk (uniform branch)
        // Householder with the UN-normalised vector v' = [alpha - beta ; raw column]:
        //   H = I - tau' v' v'^T,  tau' = 1 / (beta^2 - alpha beta) = 1 / (nrm2 + |alpha| |beta|)
        // (one rsqrt + one reciprocal on the critical path instead of sqrt + two divisions)
        const double alpha = Rp[off(j)];
        const double nrm2 = fma(alpha, alpha, sigma);
        const double absb = nrm2 * rsqrt(nrm2);
        const double beta = -copysign(absb, alpha);
        const double amb = alpha - beta;
        const double taup = __drcp_rn(fma(fabs(alpha), absb, nrm2));
        bool act[NSLOT];
        double rjk[NSLOT];
        constexpr int ACC = (ROWS % 4 == 0) ? 4 : 2;   // independent accumulator chains per slot (DFMA latency)
        double d[NSLOT][ACC];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          act[s] = (s < sj) || (s == sj && lane > lj && mycol[s] >= 0);
          rjk[s] = (s <= sj && act[s]) ? Rp[off(j) + (mycol[s] - j)] : 0.0;
#pragma unroll
          for (int q = 0; q < ACC; ++q) d[s][q] = 0.0;
        }
        const double* vw = vb + warp * ROWS;
#pragma unroll
        for (int r = 0; r < ROWS; r += 2) {
          const double2 v = *reinterpret_cast<const double2*>(vw + r);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s <= sj) {
              d[s][r % ACC] = fma(v.x, tile[r][s], d[s][r % ACC]);
              d[s][(r + 1) % ACC] = fma(v.y, tile[r + 1][s], d[s][(r + 1) % ACC]);
            }
        }
        double dsum[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          double t = d[s][0] + d[s][1];
          if (ACC == 4) t += d[s][2] + d[s][3];
          dsum[s] = t;
        }
        if constexpr (W > 1) {
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s <= sj && act[s]) s_part[(warp * NSLOT + s) * 32 + lane] = dsum[s];
          __syncthreads();  // C: partial dots visible
        }
        double wks[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          wks[s] = 0.0;
          if (s <= sj && act[s]) {
            double acc = 0.0;
            if constexpr (W > 1) {
#pragma unroll
              for (int w = 0; w < W; ++w) acc += s_part[(w * NSLOT + s) * 32 + lane];
            } else {
              acc = dsum[s];
            }
            const double wk = taup * fma(amb, rjk[s], acc);
            if (warp == 0) Rp[off(j) + (mycol[s] - j)] = fma(-wk, amb, rjk[s]);
            wks[s] = wk;
          }
        }
#pragma unroll
        for (int r = 0; r < ROWS; r += 2) {
          const double2 v = *reinterpret_cast<const double2*>(vw + r);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s <= sj) { tile[r][s] = fma(-wks[s], v.x, tile[r][s]); tile[r + 1][s] = fma(-wks[s], v.y, tile[r + 1][s]); }
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

// ---- single-warp stream kernel (large batches) -------------------------------------------------
// Same algorithm, one warp per CTA owning one row range of one sequence, restructured around the
// per-reflector latency chain (ncu r01c: 43% of the cycles were fixed-latency waits, the FP64 pipe
// was 43% busy with 2 warps per scheduler):
//   * every lane keeps the squared norm of its own column of the live slot up to date INSIDE the
//     rank-1 update loop (extra FMAs interleave with the update, nothing waits on them), so the
//     norm of the next reflector is one shuffle away when the step starts;
//   * the next column is published to the other half of vbuf at the end of the step (one
//     __syncwarp per step, no norm pass on the critical path);
//   * beta/tau come from branch-free MUFU seeds + Newton steps so the scalar chain sits in the same
//     basic block as the dot-product loop and is scheduled underneath it.
#include "k_qr_mma.cuh"

template <int NSLOT, int ROWS, int MINB, bool BF>
__global__ void __launch_bounds__(32, MINB) k_qr_stream(QrArgs a) {
  static_assert(ROWS % 4 == 0 && ROWS <= 32, "ROWS must be a multiple of 4, <= 32");
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int lane = threadIdx.x;
  const int n = a.n, nc1 = n + 1, ldo = a.ldo;
  const int r0 = nc1 - 32 * (NSLOT - 1);
  const int npk = n * (n + 3) / 2;
  double* Rp = sm;
  double* vbuf = Rp + npk + 2 - (npk & 1);      // [2][ROWS] (+ [ROWS] dump area for the branch-free publish); Rp[npk] is padding
  double* dump = vbuf + 2 * ROWS;
  int* rowstart = reinterpret_cast<int*>(vbuf + 3 * ROWS);
  __shared__ int s_total, s_f0;
  for (int t = lane; t < npk; t += 32) Rp[t] = 0.0;
  const double* src_base;
  if (a.src_mode == 0) {
    const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
    if (lane == 0) {
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
    if (lane == 0) { s_total = a.dense_rows; s_f0 = 0; }
    src_base = a.dense + (size_t)b * a.dense_stride;
  }
  __syncwarp();
  const int total = s_total;
  const int nfr = (a.src_mode == 0) ? ((int)((long)a.F * (part + 1) / nparts) - s_f0) : 0;
  auto off = [&](int j) { return j * nc1 - (j * (j - 1)) / 2; };
  int mycol[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    if (s == NSLOT - 1) mycol[s] = (lane < r0) ? lane : -1;
    else mycol[s] = r0 + 32 * (NSLOT - 2 - s) + lane;
  }
  auto resolve = [&](int v) -> long {   // physical offset of virtual row v (-1: none)
    if (v >= total) return -1;
    if (a.src_mode != 0) return (long)v * ldo;
    int lo = 0, hi = nfr;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rowstart[mid] <= v) lo = mid; else hi = mid;
    }
    return ((long)(s_f0 + lo) * a.qmax + (v - rowstart[lo])) * ldo;
  };

  double tile[ROWS][NSLOT];
  unsigned cur = 0;
  for (int base = 0; base < total; base += ROWS) {
    const long phys = (lane < ROWS) ? resolve(base + lane) : -1;
    {  // pull the next chunk's rows towards L2 while this one is being eliminated
      const long nx = (lane < ROWS) ? resolve(base + ROWS + lane) : -1;
      if (nx >= 0) {
        const char* p = reinterpret_cast<const char*>(src_base + nx);
        for (int o = 0; o < nc1 * 8; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long o = __shfl_sync(0xffffffffu, phys, r);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        tile[r][s] = (o >= 0 && mycol[s] >= 0) ? __ldg(src_base + o + mycol[s]) : 0.0;
    }
#pragma unroll
    for (int sj = NSLOT - 1; sj >= 0; --sj) {
      const int nl = (sj == NSLOT - 1) ? r0 : 32;
      const int jbase = (sj == NSLOT - 1) ? 0 : r0 + 32 * (NSLOT - 2 - sj);
      if (jbase >= n) continue;
      // slot prologue: own-column norms of the slot, lane 0 publishes the first reflector
      double ss;
      {
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
        for (int r = 0; r < ROWS; r += 4) {
          q0 = fma(tile[r][sj], tile[r][sj], q0);
          q1 = fma(tile[r + 1][sj], tile[r + 1][sj], q1);
          q2 = fma(tile[r + 2][sj], tile[r + 2][sj], q2);
          q3 = fma(tile[r + 3][sj], tile[r + 3][sj], q3);
        }
        ss = (q0 + q1) + (q2 + q3);
        if (lane == 0) {
          double* vb = vbuf + cur * ROWS;
#pragma unroll
          for (int r = 0; r < ROWS; r += 2)
            *reinterpret_cast<double2*>(vb + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
        }
      }
      for (int lj = 0; lj < nl; ++lj) {
        const int j = jbase + lj;
        if (j >= n) break;  // the residual column is never eliminated
        const double* vb = vbuf + cur * ROWS;
        double* vnext = vbuf + (cur ^ 1u) * ROWS;
        cur ^= 1u;
        const bool pub = (lane == lj + 1) && (lj + 1 < nl);
        const double sigma = __shfl_sync(0xffffffffu, ss, lj);
        __syncwarp();
        if (sigma < 1e-290) {
          // nothing below the diagonal in this chunk (uniform branch): tile untouched, norms still valid
          if (pub) {
#pragma unroll
            for (int r = 0; r < ROWS; r += 2)
              *reinterpret_cast<double2*>(vnext + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
          }
          continue;
        }
        const int oj = off(j);
        const double alpha = Rp[oj];
        const double nrm2 = fma(alpha, alpha, sigma);
        const double absb = nrm2 * rsqrt_nobranch(nrm2);
        const double beta = -copysign(absb, alpha);
        const double amb = alpha - beta;
        const double taup = rcp_nobranch(fma(fabs(alpha), absb, nrm2));
        bool act[NSLOT];
        double rjk[NSLOT];
        constexpr int ACC = 4;
        double d[NSLOT][ACC];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          act[s] = (s < sj) || (s == sj && lane > lj && mycol[s] >= 0);
          rjk[s] = (s <= sj && act[s]) ? Rp[oj + (mycol[s] - j)] : 0.0;
#pragma unroll
          for (int q = 0; q < ACC; ++q) d[s][q] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < ROWS; r += 2) {
          const double2 v = *reinterpret_cast<const double2*>(vb + r);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s <= sj) {
              d[s][r % ACC] = fma(v.x, tile[r][s], d[s][r % ACC]);
              d[s][(r + 1) % ACC] = fma(v.y, tile[r + 1][s], d[s][(r + 1) % ACC]);
            }
        }
        double wks[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          wks[s] = 0.0;
          if (s <= sj && act[s]) {
            const double acc = (d[s][0] + d[s][1]) + (d[s][2] + d[s][3]);
            const double wk = taup * fma(amb, rjk[s], acc);
            Rp[oj + (mycol[s] - j)] = fma(-wk, amb, rjk[s]);
            wks[s] = wk;
          }
        }
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
        for (int r = 0; r < ROWS; r += 4) {
          const double2 va = *reinterpret_cast<const double2*>(vb + r);
          const double2 vc = *reinterpret_cast<const double2*>(vb + r + 2);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s)
            if (s <= sj) {
              tile[r][s] = fma(-wks[s], va.x, tile[r][s]);
              tile[r + 1][s] = fma(-wks[s], va.y, tile[r + 1][s]);
              tile[r + 2][s] = fma(-wks[s], vc.x, tile[r + 2][s]);
              tile[r + 3][s] = fma(-wks[s], vc.y, tile[r + 3][s]);
            }
          q0 = fma(tile[r][sj], tile[r][sj], q0);
          q1 = fma(tile[r + 1][sj], tile[r + 1][sj], q1);
          q2 = fma(tile[r + 2][sj], tile[r + 2][sj], q2);
          q3 = fma(tile[r + 3][sj], tile[r + 3][sj], q3);
        }
        ss = (q0 + q1) + (q2 + q3);
        // Every lane read alpha = Rp[oj] at the top of this step and lane lj overwrites it below: without this barrier the
        // order is only guaranteed by the lanes running in lockstep (found by the CPU execution model of tests/emul, where
        // they do not).
        __syncwarp();
        if constexpr (BF) {
          // every lane stores (non-owners into the dump area): no divergent region at the end of the step
          double* dst = pub ? vnext : dump;
#pragma unroll
          for (int r = 0; r < ROWS; r += 2)
            *reinterpret_cast<double2*>(dst + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
          Rp[(lane == lj) ? oj : npk] = beta;
        } else {
          if (pub) {
#pragma unroll
            for (int r = 0; r < ROWS; r += 2)
              *reinterpret_cast<double2*>(vnext + r) = make_double2(tile[r][sj], tile[r + 1][sj]);
          }
          if (lane == lj) Rp[oj] = beta;
        }
      }
    }
    __syncwarp();
  }
  double* out = a.out + (size_t)b * a.out_stride + (size_t)part * n * nc1;
  for (int j = 0; j < n; ++j) {
    const int oj = off(j);
    for (int k = lane; k < nc1; k += 32) out[(size_t)j * nc1 + k] = (k >= j) ? Rp[oj + (k - j)] : 0.0;
  }
}

template <int NSLOT, int ROWS, int MINB, bool BF>
void launch_stream(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st) {
  const int n = a.n;
  const size_t npk = (size_t)n * (n + 3) / 2;
  size_t smem = sizeof(double) * (npk + 2 + 3 * ROWS) + sizeof(int) * (max_frange + 2);
  IGV_SMEM_OPTIN((k_qr_stream<NSLOT, ROWS, MINB, BF>), 220 * 1024);
  dim3 grid(split, B);
  k_qr_stream<NSLOT, ROWS, MINB, BF><<<grid, 32, smem, st>>>(a);
}

template <int NSLOT, int ROWS, int W, int MINB>
void launch_one(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st) {
  const int n = a.n;
  const size_t npk = (size_t)n * (n + 3) / 2;
  size_t smem = sizeof(double) * (npk + (npk & 1) + 2 * W * ROWS + (size_t)W * 32 * NSLOT) + sizeof(int) * (max_frange + 2);
  IGV_SMEM_OPTIN((k_qr_compress<NSLOT, ROWS, W, MINB>), 220 * 1024);
  dim3 grid(split, B);
  k_qr_compress<NSLOT, ROWS, W, MINB><<<grid, W * 32, smem, st>>>(a);
}

void launch_qr(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st, int cfg) {
  const int nslot = (a.n + 1 + 31) / 32;   // 32-column register slots (DFMA kernels)
  const int nct = (a.n + 1 + 7) / 8;       // 8-column tiles (DMMA kernel)
  // IGV_QR_CFG: test/tuning knob. 0 = automatic; 20 = force the DMMA panel kernel, 8 = force the single-warp
  // DFMA stream kernel, 9 = force the multi-warp DFMA kernel (5, 1, 3: legacy variants for A/B timing).
  // one independent single-warp stream per CTA when there are enough CTAs to fill the chip
  const bool many = a.src_mode == 0 && (long)B * split >= 1000;
  if (nct <= 9 && (cfg == 20 || (cfg == 0 && many))) {
    if (nct <= 4) launch_mma<4, 12>(a, split, B, max_frange, st);
    else if (nct <= 6) launch_mma<6, 10>(a, split, B, max_frange, st);
    else launch_mma<9, 8>(a, split, B, max_frange, st);
    return;
  }
  if (nslot <= 3 && (cfg == 8 || (cfg == 0 && many))) {
    if (nslot <= 1) launch_stream<1, 32, 12, false>(a, split, B, max_frange, st);
    else if (nslot <= 2) launch_stream<2, 32, 9, false>(a, split, B, max_frange, st);
    else launch_stream<3, 32, 8, false>(a, split, B, max_frange, st);
    return;
  }
  if (nslot <= 1) launch_one<1, 32, 4, 2>(a, split, B, max_frange, st);
  else if (nslot <= 2) launch_one<2, 32, 4, 2>(a, split, B, max_frange, st);
  else if (nslot <= 3) {
    if (cfg == 5) launch_one<3, 32, 1, 8>(a, split, B, max_frange, st);
    else if (cfg == 1) launch_one<3, 16, 4, 3>(a, split, B, max_frange, st);
    else if (cfg == 3) launch_one<3, 32, 8, 1>(a, split, B, max_frange, st);
    else launch_one<3, 32, 4, 2>(a, split, B, max_frange, st);
  }
  else if (nslot <= 4) launch_one<4, 24, 4, 2>(a, split, B, max_frange, st);
  else if (nslot <= 6) launch_one<6, 16, 4, 2>(a, split, B, max_frange, st);
  else if (nslot <= 8) launch_one<8, 12, 4, 2>(a, split, B, max_frange, st);
  else launch_one<13, 6, 4, 2>(a, split, B, max_frange, st);
}

}  // namespace

void igv_launch_qr_compress(igv_batch* h, int F, int max_valid) {
  IgvProfScope prof_scope_(h, IGV_K_QR);
  IgvLayout L = h->layout();
  const int n = 6 * L.n_clones;
  // enough CTAs to cover the chip: split the rows of each sequence when the batch is small
  int split = 1;
  // (wide stacks -- more than 72 columns: the super-block Gram kernel, 9 warps per CTA -- want 8 CTAs per SM, measured on
  // c5 / B = 148: 9.2 / 7.6 / 7.1 / 6.9 ms with 2 / 4 / 6 / 8 parts; the narrow kernels are served by 2)
  const bool gram_path = !(((h->knobs.qr_cfg > 0 && h->knobs.qr_cfg != 30) || h->compress == IGV_COMPRESS_HOUSEHOLDER) && !h->stack_f32) &&
                         igv_gram_supported(n);
  const int target = ((n + 1 > 72 && !h->feat_fused && gram_path) ? 8 : 2) * 148;
  if (h->B < target) split = min(h->qr_split_cap, max(1, min((target + h->B - 1) / h->B, (F + 7) / 8)));
  if (h->knobs.qr_split >= 1)   // test knob: force the row split
    split = min(h->qr_split_cap, min(h->knobs.qr_split, max(1, F)));
  if (h->feat_fused) {   // k_msckf_features<FUSE> already left the Gram matrix of the accepted stack in Gws
    h->feat_fused = false;
    h->last_visual_path = 2;
    igv_launch_gram_factor(h, 1);
    return;
  }
  {
    // IGV_QR_CFG (test knob) >= 1 forces one of the Householder kernels, 30 forces the Gram path
    const int cfg = h->knobs.qr_cfg;
    const bool forced_hh = ((cfg > 0 && cfg != 30) || h->compress == IGV_COMPRESS_HOUSEHOLDER) && !h->stack_f32;
    if (!forced_hh && igv_gram_supported(n)) {
      h->last_visual_path = 1;
      igv_launch_gram_compress(h, F, max_valid, split);
      return;
    }
  }
  h->last_visual_path = 0;
  QrArgs a;
  a.Hs = h->Hs; a.F = F; a.qmax = h->qmax; a.ldo = n + 1; a.f_rows = h->f_rows; a.max_valid = max_valid;
  a.dense = nullptr; a.dense_stride = 0; a.dense_rows = 0; a.src_mode = 0;
  a.n = n;
  a.n_acc = h->n_acc;
  a.F_alloc = h->cfg.max_feats;
  a.hs_seq_stride = (size_t)h->cfg.max_feats * h->qmax * (h->ncols_max + 1);
  if (split == 1) {
    a.out = h->Hc; a.out_stride = (long)h->ncols_max * (h->ncols_max + 1);
    launch_qr(a, 1, h->B, F, h->stream, h->knobs.qr_cfg);
    h->launches++;
  } else {
    a.out = h->Rpart; a.out_stride = (long)h->qr_split_cap * h->ncols_max * (h->ncols_max + 1);
    launch_qr(a, split, h->B, (F + split - 1) / split + 1, h->stream, h->knobs.qr_cfg);
    QrArgs c = a;
    c.src_mode = 1; c.dense = h->Rpart; c.dense_stride = a.out_stride; c.dense_rows = split * n;
    c.n_acc = nullptr;
    c.out = h->Hc; c.out_stride = (long)h->ncols_max * (h->ncols_max + 1);
    launch_qr(c, 1, h->B, 1, h->stream, h->knobs.qr_cfg);
    h->launches += 2;
  }
}
