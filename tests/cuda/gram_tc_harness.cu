// TEST INFRASTRUCTURE: stand-alone check of k_gram_tc (ingvio_b200/csrc/k_gram_tc.cuh) against a double-precision Gram
// matrix of the same float stack, plus a timing at a c5-sized stack.  Built and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/gram_tc_harness tests/cuda/gram_tc_harness.cu && /tmp/gram_tc_harness
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../ingvio_b200/csrc/k_gram_tc.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static int run_case(int B, int F, int qmax, int n1, int parts, int max_valid, bool timing) {
  const int ldo = n1, NC = (n1 + 31) / 32, n1p = 24 * ((n1 + 23) / 24) + 8;
  const size_t seq = (size_t)F * qmax * ldo;
  std::mt19937 rng(1234 + n1 + F);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int Bh = timing ? 1 : B;     // timing: one host sequence, replicated on the device
  std::vector<float> H(Bh * seq);
  for (auto& v : H) v = nd(rng) * std::exp(2.f * nd(rng));          // a few decades of dynamic range
  std::vector<int> fr(B * F);
  for (int b = 0; b < B; ++b)
    for (int f = 0; f < F; ++f) {
      const int u = rng() % 10;
      fr[b * F + f] = timing ? qmax : (u == 0 ? 0 : (u == 1 ? 1 + rng() % qmax : qmax));
    }
  // stale garbage in the rows that are not part of the stack
  for (int b = 0; b < Bh; ++b)
    for (int f = 0; f < F; ++f)
      for (int i = fr[b * F + f]; i < qmax; ++i)
        for (int c = 0; c < ldo; ++c) H[b * seq + ((size_t)f * qmax + i) * ldo + c] = NAN;
  float* dH; int* dfr; double* dG; int* dnacc;
  CK(cudaMalloc(&dH, (size_t)B * seq * 4)); CK(cudaMalloc(&dfr, fr.size() * 4));
  CK(cudaMalloc(&dG, (size_t)B * parts * n1p * n1p * 8)); CK(cudaMalloc(&dnacc, B * 4));
  CK(cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice));
  for (int b = Bh; b < B; ++b) CK(cudaMemcpy(dH + (size_t)b * seq, dH, seq * 4, cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(dfr, fr.data(), fr.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dG, 0xff, (size_t)B * parts * n1p * n1p * 8));
  igv_tc::GramTcArgs a;
  a.Hs = dH; a.hs_seq_stride = seq; a.F = F; a.F_alloc = F; a.qmax = qmax; a.ldo = ldo; a.f_rows = dfr; a.max_valid = max_valid;
  a.drain_stages = getenv("TC_DRAIN") ? atoi(getenv("TC_DRAIN")) : 0;
  a.n1 = n1; a.NC = NC; a.G = dG; a.g_seq_stride = (long)parts * n1p * n1p; a.n1p = n1p; a.n_acc = dnacc;
  const size_t smem = igv_tc::gram_tc_smem_bytes(NC, (F + parts - 1) / parts);
  CK(cudaFuncSetAttribute(igv_tc::k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  dim3 grid(parts, B);
  igv_tc::k_gram_tc<<<grid, igv_tc::kThreads, smem, 0>>>(a);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (timing) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) igv_tc::k_gram_tc<<<grid, igv_tc::kThreads, smem, 0>>>(a);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)B * F * qmax;
    printf("timing B=%d F=%d qmax=%d n1=%d parts=%d: %.3f ms per launch, %.1f GB/s of stack, %.1f TFLOP/s (3 x TF32 MMAs counted once: m n1^2)\n",
           B, F, qmax, n1, parts, ms / 10, rows * n1 * 4 / (ms / 10 * 1e-3) / 1e9, rows * n1 * n1 / (ms / 10 * 1e-3) / 1e12);
  }
  std::vector<double> G((size_t)B * parts * n1p * n1p);
  std::vector<int> nacc(B);
  CK(cudaMemcpy(G.data(), dG, G.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nacc.data(), dnacc, B * 4, cudaMemcpyDeviceToHost));
  double worst = 0.0; int bad = 0;
  const int bcheck = timing ? 1 : B;
  for (int b = 0; b < bcheck; ++b) {
    std::vector<double> ref((size_t)n1 * n1, 0.0);
    int cnt = 0;
    for (int f = 0; f < F; ++f) {
      const int q = fr[b * F + f];
      const bool on = q > 0 && (max_valid <= 0 || cnt < max_valid);
      if (q > 0) ++cnt;
      if (!on) continue;
      for (int i = 0; i < q; ++i) {
        const float* row = &H[b * seq + ((size_t)f * qmax + i) * ldo];
        for (int r = 0; r < n1; ++r)
          for (int c = r; c < n1; ++c) ref[(size_t)r * n1 + c] += (double)row[r] * (double)row[c];
      }
    }
    const int want = max_valid > 0 ? std::min(cnt, max_valid) : cnt;
    if (nacc[b] != want) { printf("  n_acc[%d] = %d, expected %d\n", b, nacc[b], want); ++bad; }
    for (int r = 0; r < n1; ++r)
      for (int c = r; c < n1; ++c) {
        double g = 0.0;
        for (int p = 0; p < parts; ++p) g += G[((size_t)b * parts + p) * n1p * n1p + (size_t)r * n1p + c];
        const double scale = std::sqrt(ref[(size_t)r * n1 + r] * ref[(size_t)c * n1 + c]) + 1e-300;
        const double err = std::fabs(g - ref[(size_t)r * n1 + c]) / scale;
        if (!(err <= worst)) { worst = err; }
        if (!(err < 1e-5)) { if (bad < 8) printf("  b=%d G[%d][%d] = %.9e, expected %.9e (rel %.2e)\n", b, r, c, g, ref[(size_t)r * n1 + c], err); ++bad; }
      }
  }
  printf("case B=%d F=%d qmax=%d n1=%d parts=%d max_valid=%d: worst |dG| / sqrt(G_rr G_cc) = %.3e, %d bad entries -> %s\n", B, F, qmax, n1,
         parts, max_valid, worst, bad, bad ? "FAIL" : "ok");
  cudaFree(dH); cudaFree(dfr); cudaFree(dG); cudaFree(dnacc);
  return bad ? 1 : 0;
}

int main() {
  int rc = 0;
  rc |= run_case(2, 6, 9, 31, 1, 0, false);        // c1-sized stack: one atom
  if (rc) { printf("HARNESS FAIL\n"); return 1; }
  rc |= run_case(2, 20, 19, 67, 1, 0, false);      // c2 / c3 width: three atoms, M = 128 reads a fourth (zero) atom
  rc |= run_case(3, 30, 41, 67, 3, 12, false);     // stereo rows, three parts, max_valid cap
  rc |= run_case(2, 12, 57, 128, 1, 0, false);     // exactly four atoms
  rc |= run_case(2, 12, 57, 150, 2, 0, false);     // five atoms: tile 1 with N = 32
  rc |= run_case(2, 40, 57, 181, 1, 0, false);     // c5: six atoms
  rc |= run_case(2, 40, 57, 181, 4, 25, false);
  if (!rc) rc |= run_case(148, 400, 57, 181, 1, 0, true);   // c5 at B = 148
  if (!rc) rc |= run_case(1184, 110, 41, 67, 1, 0, true);   // c3 at B = 1184
  printf(rc ? "HARNESS FAIL\n" : "HARNESS OK\n");
  return rc;
}
