// State / covariance lifecycle kernels: init, get/set, marginal blocks, add variable, marginalise,
// clone augmentation, box-plus, trace.  Reference: StateManager.cpp:121-296, State.cpp:60-167.
#include "igv_device.cuh"

using namespace igv;

namespace {

__global__ void k_state_init(double* P, int ld, double* X, int xsize, const double* R, const double* p,
                             const double* v, const double* bg, const double* ba, const double* Rext,
                             const double* pext, const double* diag21) {
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  for (int t = threadIdx.x; t < 21 * 21; t += blockDim.x) {
    const int i = t % 21, j = t / 21;
    Pb[i + (size_t)j * ld] = (i == j) ? diag21[i] : 0.0;
  }
  double* Xb = X + (size_t)b * xsize;
  for (int t = threadIdx.x; t < xsize; t += blockDim.x) {
    double val = 0.0;
    if (t < 9) val = R[b * 9 + t];
    else if (t < 12) val = p[b * 3 + t - 9];
    else if (t < 15) val = v[b * 3 + t - 12];
    else if (t < 18) val = bg[b * 3 + t - 15];
    else if (t < 21) val = ba[b * 3 + t - 18];
    else if (t < 30) val = Rext[b * 9 + t - 21];
    else if (t < 33) val = pext[b * 3 + t - 30];
    Xb[t] = val;
  }
}

__global__ void k_cov_copy(double* P, int ld, int N, double* user, int ldu, int to_user) {
  const int b = blockIdx.y;
  double* Pb = P + (size_t)b * ld * ld;
  double* Ub = user + (size_t)b * ldu * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x) {
    const int i = t % N, j = t / N;
    if (to_user) Ub[i + (size_t)j * ldu] = Pb[i + (size_t)j * ld];
    else Pb[i + (size_t)j * ld] = Ub[i + (size_t)j * ldu];
  }
}

__global__ void k_cov_blocks(const double* P, int ld, IgvBlocks blk, double* dst) {
  __shared__ int cols[6 * IGV_MAX_BLOCKS];
  const int b = blockIdx.x, n = blk.n;
  for (int q = threadIdx.x; q < blk.n_blocks; q += blockDim.x) {
    int off = 0;
    for (int r = 0; r < q; ++r) off += blk.size[r];
    for (int k = 0; k < blk.size[q]; ++k) cols[off + k] = blk.idx[q] + k;
  }
  __syncthreads();
  const double* Pb = P + (size_t)b * ld * ld;
  double* D = dst + (size_t)b * n * n;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t % n, j = t / n;
    D[t] = Pb[cols[i] + (size_t)cols[j] * ld];
  }
}

// append a variable of `size` with covariance block (col-major size x size), zero cross terms
__global__ void k_add_variable(double* P, int ld, int N, int size, const double* blk) {
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  const int Nn = N + size;
  for (int t = threadIdx.x; t < Nn * size; t += blockDim.x) {
    const int i = t % Nn, k = t / Nn;  // column N+k, row i  and mirrored
    double val = 0.0;
    if (i >= N) val = blk[(i - N) + k * size];
    Pb[i + (size_t)(N + k) * ld] = val;
    Pb[(N + k) + (size_t)i * ld] = (i >= N) ? blk[k + (i - N) * size] : 0.0;
  }
}

// out-of-place deletion of rows/cols [s, s+sz) (StateManager.cpp:167-177) + mean compaction
__global__ void k_marginalize(const double* Pin, double* Pout, int ld, int N, int s, int sz, const double* Xin,
                              double* Xout, int xsize, int clone_slot, int n_clones, int lm_slot, int lm_off, int n_lm) {
  const int b = blockIdx.y;
  const double* A = Pin + (size_t)b * ld * ld;
  double* O = Pout + (size_t)b * ld * ld;
  const int Nn = N - sz;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Nn * Nn; t += gridDim.x * blockDim.x) {
    const int i = t % Nn, j = t / Nn;
    const int si = i + (i >= s ? sz : 0), sj = j + (j >= s ? sz : 0);
    O[i + (size_t)j * ld] = A[si + (size_t)sj * ld];
  }
  if (blockIdx.x == 0) {
    const double* xi = Xin + (size_t)b * xsize;
    double* xo = Xout + (size_t)b * xsize;
    for (int t = threadIdx.x; t < xsize; t += blockDim.x) {
      double val = xi[t];
      if (clone_slot >= 0 && t >= IGV_X_CORE + 12 * clone_slot && (lm_off <= 0 || t < lm_off)) {
        const int src = t + 12;
        val = (src < IGV_X_CORE + 12 * n_clones && src < xsize) ? xi[src] : 0.0;
      }
      if (lm_slot >= 0 && t >= lm_off + 3 * lm_slot) {   // landmark values close the gap
        const int src = t + 3;
        val = (src < lm_off + 3 * n_lm && src < xsize) ? xi[src] : 0.0;
      }
      xo[t] = val;
    }
  }
}

__global__ void k_set_gnss_value(double* X, int xsize, int gtype, const double* value, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) X[(size_t)b * xsize + 33 + gtype] = value ? value[b] : 0.0;
}

// StateManager::augmentSlidingWindowPose (StateManager.cpp:253-296): P grows by 6.
__global__ void k_augment(double* P, int ld, int N, double* X, int xsize, int slot, const double* R_user,
                          const double* cR_user, const double* cp_user) {
  extern __shared__ double sm[];
  double* row = sm;  // 6 x (N+6), row-major
  const int b = blockIdx.x;
  double* Pb = P + (size_t)b * ld * ld;
  double* Xb = X + (size_t)b * xsize;
  const double* R = R_user ? R_user + (size_t)b * 9 : Xb;  // R_i2w row-major
  const int W = N + 6;
  // new rows = J * P[0:21, :]  with J = [I6 | 0 | blockdiag(R,R)]
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    double a[6], e[6];
    for (int k = 0; k < 6; ++k) {
      a[k] = Pb[k + (size_t)j * ld];
      e[k] = Pb[15 + k + (size_t)j * ld];
    }
    for (int r = 0; r < 3; ++r) {
      row[r * W + j] = a[r] + R[3 * r] * e[0] + R[3 * r + 1] * e[1] + R[3 * r + 2] * e[2];
      row[(3 + r) * W + j] = a[3 + r] + R[3 * r] * e[3] + R[3 * r + 1] * e[4] + R[3 * r + 2] * e[5];
    }
  }
  __syncthreads();
  // corner = rows[:, 0:21] J^T
  if (threadIdx.x < 36) {
    const int r = threadIdx.x / 6, c = threadIdx.x % 6;
    const double* rr = row + r * W;
    const int cc = c % 3, base = (c < 3) ? 15 : 18;
    row[r * W + N + c] = rr[c] + rr[base] * R[3 * cc] + rr[base + 1] * R[3 * cc + 1] + rr[base + 2] * R[3 * cc + 2];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 6 * N; t += blockDim.x) {
    const int r = t / N, j = t % N;
    const double val = row[r * W + j];
    Pb[(N + r) + (size_t)j * ld] = val;
    Pb[j + (size_t)(N + r) * ld] = val;
  }
  if (threadIdx.x < 36) {
    const int r = threadIdx.x / 6, c = threadIdx.x % 6;
    Pb[(N + r) + (size_t)(N + c) * ld] = 0.5 * (row[r * W + N + c] + row[c * W + N + r]);
  }
  // clone mean: T_i2w * T_cl2i (StateManager.cpp:263-272)
  if (threadIdx.x == 0) {
    double* c = Xb + IGV_X_CORE + 12 * slot;
    if (cR_user) {
      for (int i = 0; i < 9; ++i) c[i] = cR_user[(size_t)b * 9 + i];
      for (int i = 0; i < 3; ++i) c[9 + i] = cp_user[(size_t)b * 3 + i];
    } else {
      double Rc[9], t[3];
      mat3_mul(Xb, Xb + 21, Rc);
      mat3_vec(Xb, Xb + 30, t);
      for (int i = 0; i < 9; ++i) c[i] = Rc[i];
      for (int i = 0; i < 3; ++i) c[9 + i] = t[i] + Xb[9 + i];
    }
  }
}

__global__ void k_boxplus(double* X, int xsize, const double* dx, int N, IgvLayout L) {
  const int b = blockIdx.x;
  boxplus_all(X + (size_t)b * xsize, dx + (size_t)b * N, L);
}

__global__ void k_trace(const double* P, int ld, int N, double* out) {
  const int b = blockIdx.x;
  const double* Pb = P + (size_t)b * ld * ld;
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += 32) s += Pb[i + (size_t)i * ld];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[b] = s;
}

}  // namespace

void igv_launch_state_init(igv_batch* h, const double* R, const double* p, const double* v, const double* bg,
                           const double* ba, const double* Rext, const double* pext, const double* diag21) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_state_init<<<h->B, 128, 0, h->stream>>>(h->Pc(), h->ld, h->Xc(), h->xsize, R, p, v, bg, ba, Rext, pext, diag21);
  h->launches++;
}
void igv_launch_cov_copy(igv_batch* h, double* user, int ld_user, bool to_user) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  dim3 grid(max(1, min(64, (h->N * h->N + 255) / 256)), h->B);
  k_cov_copy<<<grid, 256, 0, h->stream>>>(h->Pc(), h->ld, h->N, user, ld_user, to_user ? 1 : 0);
  h->launches++;
}
void igv_launch_cov_blocks(igv_batch* h, const IgvBlocks& blk, double* dst) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_cov_blocks<<<h->B, 256, 0, h->stream>>>(h->Pc(), h->ld, blk, dst);
  h->launches++;
}
void igv_launch_add_variable(igv_batch* h, int size, const double* cov_block_dev) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_add_variable<<<h->B, 128, 0, h->stream>>>(h->Pc(), h->ld, h->N, size, cov_block_dev);
  h->launches++;
}
void igv_launch_marginalize(igv_batch* h, int start, int size, int clone_slot, int lm_slot) {
  IgvProfScope prof_scope_(h, IGV_K_MARG);
  const int Nn = h->N - size;
  dim3 grid(max(1, min(32, (Nn * Nn + 255) / 256)), h->B);
  IgvLayout L = h->layout();
  k_marginalize<<<grid, 256, 0, h->stream>>>(h->P[h->cur], h->P[h->cur ^ 1], h->ld, h->N, start, size,
                                             h->X[h->xcur], h->X[h->xcur ^ 1], h->xsize, clone_slot, L.n_clones,
                                             lm_slot, L.lm_off, L.n_lm);
  h->cur ^= 1;
  h->xcur ^= 1;
  h->launches++;
}
void igv_launch_set_gnss_value(igv_batch* h, int gtype, const double* value_dev) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_set_gnss_value<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->Xc(), h->xsize, gtype, value_dev, h->B);
  h->launches++;
}
void igv_launch_augment(igv_batch* h, const double* R_i2w, const double* clone_R, const double* clone_p) {
  IgvProfScope prof_scope_(h, IGV_K_AUGMENT);
  IgvLayout L = h->layout();
  const size_t smem = sizeof(double) * 6 * (h->N + 6);
  k_augment<<<h->B, 128, smem, h->stream>>>(h->Pc(), h->ld, h->N, h->Xc(), h->xsize, L.n_clones, R_i2w, clone_R,
                                            clone_p);
  h->launches++;
}
void igv_launch_boxplus(igv_batch* h, const double* dx) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_boxplus<<<h->B, 64, 0, h->stream>>>(h->Xc(), h->xsize, dx, h->N, h->layout());
  h->launches++;
}
void igv_launch_trace(igv_batch* h, double* out) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trace<<<h->B, 32, 0, h->stream>>>(h->Pc(), h->ld, h->N, out);
  h->launches++;
}
