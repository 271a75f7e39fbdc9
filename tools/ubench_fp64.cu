// Micro-benchmarks that size the QR stream kernel's latency model on the box at hand:
// dependent-issue latency of DFMA / MUFU.RSQ64H / LDS / SHFL and DFMA issue rate vs independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_fp64.cu -o ingvio_b200/lib/ubench_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dfma(double* out, long long* cyc, double a, double b, int iters) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// register-file pressure: every DFMA reads 3 distinct 64-bit registers (MODE 0), or shares one multiplicand
// across the chains (MODE 1: the register-tiled rank-1 update pattern, operand-reuse cache can help)
template <int CH, int MODE>
__global__ void k_dfma3(double* out, long long* cyc, const double* in, int iters) {
  double x[CH], y[CH], z[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { x[c] = in[(c)*256 + threadIdx.x]; y[c] = in[(CH + c)*256 + threadIdx.x]; z[c] = in[(2 * CH + c)*256 + threadIdx.x]; }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c) x[c] = fma(y[c], MODE == 0 ? z[c] : z[u % 4], x[c]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH, int MODE>
void run_dfma3(double* out, long long* cyc, const double* in, int warps) {
  const int iters = 512;
  k_dfma3<CH, MODE><<<1, 32 * warps>>>(out, cyc, in, iters);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma3 mode=%d chains=%d warps/CTA=%d: %.2f cycles per DFMA per warp\n", MODE, CH, warps, (double)c / (iters * 8.0 * CH));
}

__global__ void k_rsq(double* out, long long* cyc, double a, int iters) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_lds(double* out, long long* cyc, int iters) {
  __shared__ int idx[64];
  if (threadIdx.x < 64) idx[threadIdx.x] = (threadIdx.x + 1) & 63;
  __syncthreads();
  int p = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = idx[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_shfl(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH>
void run_dfma(double* out, long long* cyc, int warps) {
  const int iters = 256;
  k_dfma<CH><<<1, 32 * warps>>>(out, cyc, 0.999, 1e-3, iters);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dfma chains=%d warps/CTA=%d: %.2f cycles per DFMA per warp, %.2f cycles per dependent step\n", CH, warps,
         (double)c / (iters * 16.0 * CH), (double)c / (iters * 16.0));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 16);
  for (int rep = 0; rep < 2; ++rep) {
    run_dfma<1>(out, cyc, 1); run_dfma<2>(out, cyc, 1); run_dfma<4>(out, cyc, 1); run_dfma<8>(out, cyc, 1);
    run_dfma<12>(out, cyc, 1); run_dfma<1>(out, cyc, 4); run_dfma<4>(out, cyc, 4); run_dfma<4>(out, cyc, 8); run_dfma<8>(out, cyc, 8);
  }
  double* in; cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
  run_dfma3<8, 0>(out, cyc, in, 1); run_dfma3<16, 0>(out, cyc, in, 1); run_dfma3<16, 0>(out, cyc, in, 4); run_dfma3<16, 0>(out, cyc, in, 8);
  run_dfma3<8, 1>(out, cyc, in, 1); run_dfma3<16, 1>(out, cyc, in, 1); run_dfma3<16, 1>(out, cyc, in, 4); run_dfma3<16, 1>(out, cyc, in, 8);
  long long c;
  k_rsq<<<1, 32>>>(out, cyc, 1.0, 256); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.RSQ64H dependent: %.2f cycles\n", (double)c / 4096.0);
  k_lds<<<1, 32>>>(out, cyc, 256); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS.32 dependent: %.2f cycles\n", (double)c / 4096.0);
  k_shfl<<<1, 32>>>(out, cyc, 256); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64 dependent: %.2f cycles\n", (double)c / 4096.0);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
