// Gram matrix of the single-precision MSCKF stack on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM) -- the reduced-precision compression of IGV_PREC_TF32_GRAM.  Reference call sites of the stage it serves:
// RemoveLostUpdate.cpp:139-155, SwMargUpdate.cpp:161-176, KeyframeUpdate.cpp:557-572 (Eigen::SPQR of the stacked Jacobian).
//
//   G = W^T W,  W = [H | r]  (m rows x n1 <= 192 columns, float, rows of the accepted tracks in track order)
//
// Arithmetic (profiles/r02_tcgen05_eval.md): every float is split x = hi + lo with hi, lo representable in TF32
// (cvt.rna.tf32.f32), and  W^T W ~= hi^T hi + hi^T lo + lo^T hi  runs as three UMMAs per 8 rows with FP32 accumulation in
// TMEM; after every 128 rows the accumulator is drained into FP64 sums held in shared memory (so the FP32 accumulation
// never spans more than 16 K-steps), and the FP64 upper triangle goes to the same partial-Gram buffer the FP64 kernels of
// k_gram.cu write -- k_gram_factor and the EKF update behind it stay FP64.
//
// Operand layout.  Both operands of W^T W contract over the ROW index of the row-major stack, i.e. they are "MN-major"
// in UMMA terms (the M / N index is the contiguous one), which kind::tf32 accepts (instruction descriptor bits 15 / 16).
// For 32-bit MN-major operands the only shared-memory layout the unit accepts is SWIZZLE_128B_BASE32B (layout type 1): atoms of
// 4 stack rows x 128 bytes (32 columns), the four 32-byte chunks of a row XOR-ed with the row index. Four stack rows are
// stored as NCs such atoms: atom stride (leading byte offset) 512 B, stride between 4-row groups (stride byte offset)
// NCs * 512 B; one UMMA (K = 8) spans two groups. One descriptor addresses 4 atoms (M = 128) or N / 32 atoms from its start
// address, so the SAME bytes serve as A (= W^T, rows of G) and as B (= W, columns of G):
//     tile 0:  G[0:128, 0:32 NC]           A = atoms 0..3,        B = atoms 0..NC-1   (M = 128, N = 32 NC)
//     tile 1:  G[32(NC-4) : 32 NC, 128 : 32 NC]   A = atoms NC-4..NC-1,  B = atoms 4..NC-1   (M = 128, N = 32 (NC-4)), NC > 4 only
// (tile 1 recomputes a few rows of tile 0 instead of needing an M = 64 shape; only the upper triangle is kept).
//
// Roles (13 warps): warps 0-3 drain TMEM (warp w may only touch TMEM lanes 32 w .. 32 w + 31 = rows of the tile),
// warp 4 allocates TMEM and issues the MMAs from one thread, warps 5-12 stream the stack: global float -> registers ->
// hi / lo split -> swizzled shared memory, three 16-row stages, full / empty mbarriers; tcgen05.commit releases a stage
// and publishes a finished 128-row accumulator, which is double buffered (2 x 256 TMEM columns).
#pragma once
#include <cstddef>
#include <cstdint>

namespace igv_tc {

constexpr int kStageRows = 16;       // two K-groups of 8 rows
constexpr int kStages = 3;
constexpr int kDrainStages = 8;      // 128 rows per FP32 accumulation
constexpr int kThreads = 13 * 32;
constexpr int kProducerWarp0 = 5, kProducerWarps = 8, kMmaWarp = 4;

struct GramTcArgs {
  const float* Hs; size_t hs_seq_stride;   // floats per sequence
  int F, F_alloc, qmax, ldo;
  const int* f_rows; int max_valid;
  int n1;                   // columns of the stack (n + 1) <= 192
  int NC;                   // ceil(n1 / 32)
  double* G; long g_seq_stride; int n1p;   // partial Gram matrices [b][part][n1p x n1p], row-major, upper triangle
  int* n_acc;
  int drain_stages;         // 16-row stages per FP32 accumulation in TMEM (8 = 128 rows); 0 = default
};

__host__ __device__ inline int tc_ncs(int NC) { return NC < 4 ? 4 : NC; }
__host__ __device__ inline int tc_stage_bytes(int NC) { return 2 * (kStageRows / 8) * tc_ncs(NC) * 1024; }   // hi + lo
// FP64 sums of the upper triangle, column-major packed: column j holds rows 0..min(j, 127); behind them tile 1's rows
// 128.. (columns 128..), also packed by column
__host__ __device__ inline int tc_off0(int j) { return j < 128 ? (j * (j + 1)) / 2 : 8256 + (j - 128) * 128; }
__host__ __device__ inline int tc_nacc(int NC) {
  const int c = 32 * NC;
  return tc_off0(c) + (NC > 4 ? ((c - 128) * (c - 127)) / 2 : 0);
}
__host__ inline size_t gram_tc_smem_bytes(int NC, int frange) {
  return 1024 /* alignment slack */ + (size_t)kStages * tc_stage_bytes(NC) + sizeof(double) * tc_nacc(NC) +
         sizeof(int) * (frange + 2) + 256;
}

#ifndef IGV_EMULATE
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error becomes a trap (launch failure) instead of a hung GPU
__device__ __forceinline__ void tc_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, MN-major (cute::UMMA::SmemDescriptor: start address [0,14), leading byte offset
// [16,30), stride byte offset [32,46), version 1 at [46,48), layout type SWIZZLE_128B_BASE32B = 1 at [61,64);
// all offsets in units of 16 bytes)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (1ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both MN-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  // load and wait in ONE statement: the registers are defined only after tcgen05.wait::ld
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
}
__device__ __forceinline__ uint32_t tc_to_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return y;
}

// grid (parts, B), kThreads threads, gram_tc_smem_bytes() of dynamic shared memory (one CTA per SM: it owns all of TMEM)
__global__ void __launch_bounds__(kThreads, 1) k_gram_tc(GramTcArgs a) {
  extern __shared__ unsigned char tc_sm_raw[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NC = a.NC, NCs = tc_ncs(NC), n1 = a.n1;
  const uint32_t raw = tc_smem_u32(tc_sm_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;            // the swizzle acts on absolute address bits 4..9
  unsigned char* sm = tc_sm_raw + pad;
  const int stage_bytes = tc_stage_bytes(NC), half_bytes = stage_bytes / 2, kg_bytes = NCs * 1024;
  unsigned char* stages = sm;
  double* acc = reinterpret_cast<double*>(sm + (size_t)kStages * stage_bytes);
  const int nacc = tc_nacc(NC);
  int* rowstart = reinterpret_cast<int*>(acc + nacc);
  const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
  const int nfr = f1 - f0;
  __shared__ __align__(8) unsigned long long bars[2 * kStages + 4];
  __shared__ uint32_t s_tmem;
  __shared__ int s_total;
  const uint32_t bar_full = tc_smem_u32(&bars[0]), bar_empty = tc_smem_u32(&bars[kStages]),
                 bar_tfull = tc_smem_u32(&bars[2 * kStages]), bar_tempty = tc_smem_u32(&bars[2 * kStages + 2]);

  if (tid == 0) {
    // accepted tracks before f0 count towards the max_valid cap (RemoveLostUpdate.cpp:120-122)
    const int* fr = a.f_rows + (size_t)b * a.F_alloc;
    int cnt = 0, rows = 0;
    for (int f = 0; f < f0; ++f) cnt += (fr[f] > 0);
    for (int f = f0; f < f1; ++f) {
      rowstart[f - f0] = rows;
      const bool on = fr[f] > 0 && (a.max_valid <= 0 || cnt < a.max_valid);
      if (fr[f] > 0) ++cnt;
      if (on) rows += fr[f];
    }
    rowstart[nfr] = rows;
    s_total = rows;
    if (part == nparts - 1 && a.n_acc) a.n_acc[b] = (a.max_valid > 0) ? min(cnt, a.max_valid) : cnt;
    for (int s = 0; s < kStages; ++s) {
      tc_bar_init(bar_full + 8 * s, kProducerWarps);
      tc_bar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc_bar_init(bar_tfull + 8 * s, 1);
      tc_bar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = tid; t < kStages * stage_bytes / 16; t += kThreads) reinterpret_cast<uint4*>(stages)[t] = make_uint4(0, 0, 0, 0);
  for (int t = tid; t < nacc; t += kThreads) acc[t] = 0.0;
  tc_fence_async_smem();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int total = s_total;
  const int nst = (total + kStageRows - 1) / kStageRows;
  const int dstages = a.drain_stages > 0 ? a.drain_stages : kDrainStages;
  const int ndrain = (nst + dstages - 1) / dstages;

  if (warp >= kProducerWarp0) {
    // ===== producers: two rows of every stage per warp, lane = column within a 32-column atom ===========================
    const int pw = warp - kProducerWarp0;
    const float* src = a.Hs + (size_t)b * a.hs_seq_stride;
    float x[3][2][6];     // [register buffer][row][atom]: loads run two stages ahead of the stores
    int cur[2] = {0, 0};  // track of the row each slot fetched last (rows only move forward)
    auto fetch = [&](int st, float (&dst)[2][6]) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int v = st * kStageRows + 2 * pw + rr;
        const float* srow = nullptr;
        if (v < total) {
          int c = cur[rr];
          while (rowstart[c + 1] <= v) ++c;          // v < total = rowstart[nfr]: stops at the track that holds row v
          cur[rr] = c;
          srow = src + ((size_t)(f0 + c) * a.qmax + (v - rowstart[c])) * a.ldo;
        }
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const int c = lane + 32 * t;
          dst[rr][t] = (srow != nullptr && t < NC && c < n1) ? __ldg(srow + c) : 0.f;
        }
      }
    };
    auto publish = [&](int st, const float (&val)[2][6]) {
      const int stage = st % kStages;
      tc_bar_wait(bar_empty + 8 * stage, ((st / kStages) & 1) ^ 1);
      unsigned char* hi_base = stages + (size_t)stage * stage_bytes;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        // K-group of 4 rows (one swizzle atom in K), row r4 inside it; 32-byte chunk index XOR r4 (SWIZZLE_128B_BASE32B)
        const int r = 2 * pw + rr, g4 = r >> 2, r4 = r & 3;
        const int inner = g4 * (kg_bytes >> 1) + r4 * 128 + ((((lane >> 3) ^ r4) & 3) << 5) + ((lane & 7) << 2);
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          if (t < NC) {
            const float xv = val[rr][t];
            // hi = x rounded to TF32's 10 mantissa bits (round half away on the bit pattern: two integer ops instead of
            // the quarter-rate cvt.rna.tf32), lo = x - hi (exact, <= 13 significant bits) rounded the same way
            const uint32_t h = (__float_as_uint(xv) + 0x1000u) & 0xFFFFE000u;
            const uint32_t l = (__float_as_uint(xv - __uint_as_float(h)) + 0x1000u) & 0xFFFFE000u;   // rounded too: the unit would truncate
            *reinterpret_cast<uint32_t*>(hi_base + inner + t * 512) = h;
            *reinterpret_cast<uint32_t*>(hi_base + half_bytes + inner + t * 512) = l;
          }
        }
      }
      tc_fence_async_smem();
      __syncwarp();
      if (lane == 0) tc_bar_arrive(bar_full + 8 * stage);
    };
    if (nst > 0) fetch(0, x[0]);
    if (nst > 1) fetch(1, x[1]);
    for (int st = 0; st < nst; st += 3) {
      if (st + 2 < nst) fetch(st + 2, x[2]);
      publish(st, x[0]);
      if (st + 1 < nst) {
        if (st + 3 < nst) fetch(st + 3, x[0]);
        publish(st + 1, x[1]);
      }
      if (st + 2 < nst) {
        if (st + 4 < nst) fetch(st + 4, x[1]);
        publish(st + 2, x[2]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (one thread) ======================================================================================
    if (lane == 0) {
      const uint32_t idesc0 = tc_idesc(128, 32 * NC);
      const uint32_t idesc1 = tc_idesc(128, NC > 4 ? 32 * (NC - 4) : 32);
      const uint32_t stage0 = tc_smem_u32(stages);
      for (int st = 0; st < nst; ++st) {
        const int stage = st % kStages, dchunk = st / dstages, buf = dchunk & 1;
        const bool first = (st % dstages) == 0;
        if (first) {
          tc_bar_wait(bar_tempty + 8 * buf, ((dchunk >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        tc_bar_wait(bar_full + 8 * stage, (st / kStages) & 1);
        tc_fence_after();
        const uint32_t d0 = tmem + buf * 256, d1 = d0 + 32 * NC;
#pragma unroll
        for (int kg = 0; kg < kStageRows / 8; ++kg) {
          const uint32_t hi = stage0 + stage * stage_bytes + kg * kg_bytes, lo = hi + half_bytes;
          const uint32_t accf = (first && kg == 0) ? 0u : 1u;
          const uint32_t lbo = 512, sbo = kg_bytes >> 1;     // atom stride along MN, stride between 4-row K atoms
          const uint64_t ah = tc_desc(hi, lbo, sbo), al = tc_desc(lo, lbo, sbo);
          tc_mma_tf32(d0, ah, ah, idesc0, accf);
          tc_mma_tf32(d0, ah, al, idesc0, 1u);
          tc_mma_tf32(d0, al, ah, idesc0, 1u);
          if (NC > 4) {
            const uint64_t a1h = tc_desc(hi + (NC - 4) * 512, lbo, sbo), a1l = tc_desc(lo + (NC - 4) * 512, lbo, sbo);
            const uint64_t b1h = tc_desc(hi + 4 * 512, lbo, sbo), b1l = tc_desc(lo + 4 * 512, lbo, sbo);
            tc_mma_tf32(d1, a1h, b1h, idesc1, accf);
            tc_mma_tf32(d1, a1h, b1l, idesc1, 1u);
            tc_mma_tf32(d1, a1l, b1h, idesc1, 1u);
          }
        }
        tc_commit(bar_empty + 8 * stage);                                   // the stage may be refilled once these MMAs are done
        if ((st % dstages) == dstages - 1 || st == nst - 1) tc_commit(bar_tfull + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // ===== drain: TMEM (FP32) -> FP64 sums in shared memory; warp w owns TMEM lanes 32 w .. 32 w + 31 ===================
    const int i = 32 * warp + lane;                 // TMEM lane = row of tile 0
    const int ncols = 32 * NC;
    const int lmin = 128 - 32 * (NC - 4);           // tile 1: TMEM lane l holds row 32 (NC - 4) + l of G; rows >= 128 are kept
    const int base1 = tc_off0(ncols);
    float v[32];
    for (int d = 0; d < ndrain; ++d) {
      const int buf = d & 1;
      tc_bar_wait(bar_tfull + 8 * buf, (d >> 1) & 1);
      tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(32 * warp) << 16) + buf * 256;
      for (int c = warp; c < NC; ++c) {             // columns 32 c .. 32 c + 31 hold entries j >= i only from chunk `warp` on
        tc_ld32(t0 + 32 * c, v);
        if (i < ncols) {
          // all loads, then all adds, then all stores: the entries are distinct, which the compiler cannot see
          const int jb = 32 * c;
          double* col0 = acc + tc_off0(jb) + i;               // entry (i, jb); column j + 1 starts min(j, 127) + 1 further on
          double s8[32];
          int off = 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            s8[e] = (jb + e >= i) ? col0[off] : 0.0;
            off += min(jb + e, 127) + 1;
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) s8[e] += (double)v[e];
          off = 0;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (jb + e >= i) col0[off] = s8[e];
            off += min(jb + e, 127) + 1;
          }
        }
      }
      if (NC > 4 && 32 * warp + 31 >= lmin) {
        const int i1 = i - lmin;                    // row 128 + i1
        for (int c = 0; c < NC - 4; ++c) {
          tc_ld32(t0 + ncols + 32 * c, v);
          if (i1 >= 0) {
            const int jb = 32 * c;
            double* col0 = acc + base1 + (jb * (jb + 1)) / 2 + i1;
            double s8[32];
            int off = 0;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              s8[e] = (jb + e >= i1) ? col0[off] : 0.0;
              off += jb + e + 1;
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) s8[e] += (double)v[e];
            off = 0;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              if (jb + e >= i1) col0[off] = s8[e];
              off += jb + e + 1;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_bar_arrive(bar_tempty + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
  // upper triangle to the partial Gram matrix (row-major, columns >= row), the layout k_gram_factor sums
  double* G = a.G + (size_t)b * a.g_seq_stride + (size_t)part * a.n1p * a.n1p;
  for (int row = warp; row < n1; row += kThreads / 32) {
    for (int col = row + lane; col < n1; col += 32) {
      double val;
      if (row < 128) val = acc[tc_off0(col) + row];
      else val = acc[tc_off0(32 * NC) + ((col - 128) * (col - 127)) / 2 + (row - 128)];
      G[(size_t)row * a.n1p + col] = val;
    }
  }
}
#endif   // IGV_EMULATE

}  // namespace igv_tc
