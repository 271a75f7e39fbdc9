// FP64 tensor-core (DMMA, mma.sync m8n8k4 / m16n8k8 f64) issue rate and whole-chip throughput on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_dmma.cu -o ingvio_b200/lib/ubench_dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int CH>
__global__ void k_dmma(double* out, long long* cyc, const double* in, int iters) {
  double c0[CH], c1[CH], a[CH], b[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    c0[c] = in[c * 256 + threadIdx.x]; c1[c] = in[(CH + c) * 256 + threadIdx.x];
    a[c] = in[(2 * CH + c) * 256 + threadIdx.x]; b[c] = in[(3 * CH + c) * 256 + threadIdx.x];
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c) dmma884(c0[c], c1[c], a[c], b[(c + u) % CH]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += c0[c] + c1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
__global__ void k_dmma16(double* out, long long* cyc, const double* in, int iters) {
  double d[CH][4], a[CH][4], b[CH][2];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { d[c][q] = in[(c * 10 + q) * 256 + threadIdx.x]; a[c][q] = in[(c * 10 + 4 + q) * 256 + threadIdx.x]; }
    b[c][0] = in[(c * 10 + 8) * 256 + threadIdx.x]; b[c][1] = in[(c * 10 + 9) * 256 + threadIdx.x];
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < CH; ++c) dmma1688(d[c], a[c], b[(c + u) % CH]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
void run(double* out, long long* cyc, const double* in, int warps, int blocks) {
  const int iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma<CH><<<blocks, 32 * warps>>>(out, cyc, in, iters);
  cudaEventRecord(e0);
  k_dmma<CH><<<blocks, 32 * warps>>>(out, cyc, in, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * CH;
  printf("m8n8k4  chains=%d warps/CTA=%2d CTAs=%4d: %.2f cycles per DMMA per warp; %.2f TFLOP/s\n", CH, warps, blocks, c / n,
         n * 512.0 * warps * blocks / (ms * 1e-3) / 1e12);
}
template <int CH>
void run16(double* out, long long* cyc, const double* in, int warps, int blocks) {
  const int iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma16<CH><<<blocks, 32 * warps>>>(out, cyc, in, iters);
  cudaEventRecord(e0);
  k_dmma16<CH><<<blocks, 32 * warps>>>(out, cyc, in, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * CH;
  printf("m16n8k8 chains=%d warps/CTA=%2d CTAs=%4d: %.2f cycles per DMMA per warp; %.2f TFLOP/s\n", CH, warps, blocks, c / n,
         n * 2048.0 * warps * blocks / (ms * 1e-3) / 1e12);
}

int main() {
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
  run<1>(out, cyc, in, 1, 1); run<2>(out, cyc, in, 1, 1); run<4>(out, cyc, in, 1, 1); run<8>(out, cyc, in, 1, 1);
  run<4>(out, cyc, in, 4, 1); run<4>(out, cyc, in, 8, 1); run<4>(out, cyc, in, 16, 1);
  run<4>(out, cyc, in, 8, 148); run<4>(out, cyc, in, 16, 148); run<8>(out, cyc, in, 8, 296);
  run16<1>(out, cyc, in, 1, 1); run16<2>(out, cyc, in, 1, 1); run16<4>(out, cyc, in, 1, 1);
  run16<4>(out, cyc, in, 4, 1); run16<4>(out, cyc, in, 8, 1); run16<4>(out, cyc, in, 8, 148); run16<4>(out, cyc, in, 16, 148);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
