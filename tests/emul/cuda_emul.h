// TEST INFRASTRUCTURE: a minimal CPU execution model for CUDA kernels that only use thread/block indices, shared
// memory, __syncthreads, warp ballots / shuffles, popcount and integer atomics (the track-table kernels,
// ingvio_b200/csrc/k_tracks.cu). One CTA runs at a time; its threads are real std::threads meeting at a
// std::barrier, so barrier and warp-level semantics are exercised for real. It lets `pytest -m "not gpu"` run the
// UNMODIFIED kernel source against the oracle on a machine without a GPU. It is not a product path: nothing under
// ingvio_b200/ includes or links it.
#pragma once
#include <cuda_runtime.h>   // host-side types only (dim3, uint3, cudaStream_t)

#include <barrier>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __shared__
#undef __launch_bounds__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static            /* one CTA at a time: a function-local static is shared by its threads */
#define __launch_bounds__(...)
#define __restrict__

namespace emul {
inline thread_local uint3 t_threadIdx{0, 0, 0};
inline thread_local uint3 t_blockIdx{0, 0, 0};
inline dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);
inline std::vector<char> g_dyn_smem;
struct WarpBox {
  std::unique_ptr<std::barrier<>> bar;
  unsigned long long vals[32];
  double fa[32], fb[32];          // DMMA fragments
};
inline std::vector<WarpBox> g_warps;
inline std::unique_ptr<std::barrier<>> g_cta_bar;

inline WarpBox& my_warp() { return g_warps[t_threadIdx.x >> 5]; }
inline unsigned long long exchange(unsigned long long v, int src_lane) {
  WarpBox& w = my_warp();
  w.vals[t_threadIdx.x & 31] = v;
  w.bar->arrive_and_wait();
  const unsigned long long r = w.vals[src_lane & 31];
  w.bar->arrive_and_wait();
  return r;
}

// Run kernel(args...) over grid x block with `smem` bytes of dynamic shared memory. blockDim.x must be a multiple of
// 32 and every thread of a warp that calls a warp intrinsic must reach it (true for the kernels tested here).
template <class F>
void launch(dim3 grid3, unsigned block, size_t smem, F&& body) {
  g_blockDim = dim3(block, 1, 1);
  g_gridDim = grid3;
  const unsigned grid = grid3.x * grid3.y * grid3.z;
  for (unsigned b = 0; b < grid; ++b) {
    g_dyn_smem.assign(smem + 16, 0);
    g_cta_bar = std::make_unique<std::barrier<>>(block);
    g_warps.clear();
    g_warps.resize((block + 31) / 32);
    for (auto& w : g_warps) w.bar = std::make_unique<std::barrier<>>(32);
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, b, t] {
        t_threadIdx = uint3{t, 0, 0};
        t_blockIdx = uint3{b % grid3.x, (b / grid3.x) % grid3.y, b / (grid3.x * grid3.y)};
        body();
        g_cta_bar->arrive_and_drop();   // a finished thread no longer takes part in __syncthreads
      });
    for (auto& x : th) x.join();
  }
}
// mma.sync.aligned.m8n8k4.row.col.f64: lane l supplies A(l/4, l%4) and B(l%4, l/4) and owns C(l/4, 2(l%4) + {0,1}).
inline void dmma884(double& d0, double& d1, double a, double b) {
  WarpBox& w = my_warp();
  const int lane = t_threadIdx.x & 31;
  w.fa[lane] = a;
  w.fb[lane] = b;
  w.bar->arrive_and_wait();
  const int row = lane >> 2, c0 = 2 * (lane & 3);
  for (int k = 0; k < 4; ++k) {
    d0 = std::fma(w.fa[row * 4 + k], w.fb[c0 * 4 + k], d0);
    d1 = std::fma(w.fa[row * 4 + k], w.fb[(c0 + 1) * 4 + k], d1);
  }
  w.bar->arrive_and_wait();
}
// kernel<<<grid, block, smem, stream>>>(args) as rewritten by tests/emul/preprocess.py (synchronous; 1-D blocks)
template <class F>
void launch_k(F&& body, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr) {
  launch(grid, block.x, smem, body);
}
}  // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)
#define IGV_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(emul::g_dyn_smem.data()) + 15) & ~uintptr_t(15))

inline void __syncthreads() { emul::g_cta_bar->arrive_and_wait(); }
inline void __syncwarp() { emul::exchange(0, 0); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  emul::WarpBox& w = emul::my_warp();
  w.vals[threadIdx.x & 31] = pred ? 1 : 0;
  w.bar->arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (w.vals[l] ? 1u : 0u) << l;
  w.bar->arrive_and_wait();
  return m;
}
inline int __shfl_down_sync(unsigned, int v, int delta) {
  const int lane = threadIdx.x & 31;
  const int src = lane + delta < 32 ? lane + delta : lane;
  return (int)emul::exchange((unsigned long long)(unsigned)v, src);
}
#include <algorithm>
using std::min;
using std::max;
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  raw = emul::exchange(raw, src_lane);
  std::memcpy(&v, &raw, sizeof(T));
  return v;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, (int)(threadIdx.x & 31) ^ lane_mask); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __drcp_rn(double x) { return 1.0 / x; }
inline size_t __cvta_generic_to_shared(const void*) { return 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
