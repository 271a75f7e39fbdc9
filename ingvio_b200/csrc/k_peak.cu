// FP64 FMA peak probe: every SM runs independent DFMA chains; used as the FP64-pipe roofline
// denominator in bench.py (MEASURED_PEAKS.json only carries HBM copy and bf16 GEMM peaks).
#include <cuda_runtime.h>

#include "../../include/ingvio_b200.h"

namespace {
__global__ void __launch_bounds__(512) k_dfma_probe(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" igv_status igv_measure_fp64_peak(int device, double* tflops_out) {
  if (!tflops_out) return IGV_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return IGV_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return IGV_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 4000;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return IGV_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dfma_probe<<<blocks, threads>>>(buf, iters, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return IGV_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops_out = best;
  return IGV_OK;
}
