// QR compression on the FP64 tensor pipe (DMMA.8x8x4 = mma.sync.m8n8k4.f64), included by k_qr.cu.
//
// Same "triangle on top of a rectangle" Householder elimination as k_qr_compress / k_qr_stream and the
// same reflectors in exact arithmetic, but BLOCKED: the chunk's columns are processed in panels of 8.
//   1. panel factorisation -- 8 sequential reflectors on the 32 x 8 panel, each column spread over the 4
//      threads of a "quad" (DFMA + two quad shuffles per dot product); the same dot instruction gives, for
//      the already eliminated quads, v_i . v_k, i.e. the Gram matrix the compact-WY factor needs;
//   2. T (8 x 8, LAPACK dlarft forward/columnwise recurrence), one row per lane;
//   3. trailing update of every column tile to the right with three small GEMMs on DMMA:
//        Z = D R_panel,k + V^T A_k        W = T^T Z        R_panel,k -= D W ,  A_k -= V W .
// Measured on B200 (tools/ubench_*.cu): a DFMA with three distinct register operands issues every 3.07
// cycles per scheduler (2.2 with operand reuse), i.e. the register-tiled rank-1 kernels cap at ~2/3 of the
// 37 TFLOP/s DFMA peak before any overhead, while DMMA.8x8x4 sustains 16 cycles per 256 FMAs (= the
// full 37 TFLOP/s) with one quarter of the register-file traffic and 1/8 of the issue slots.
//
// Register layout: the chunk is held TRANSPOSED in the accumulator-fragment layout, thread (g = lane/4,
// q = lane%4) owns A[row 8 rt + 2 q + {0,1}][col 8 ct + g]. With it
//   * Z^T = A^T V       : A^T is directly the A-operand (the k index, chunk rows, runs over (q, parity) in a
//                         permuted order that the B-operand -- V -- simply follows),
//   * W^T = Z^T T       : the accumulator fragment of Z^T is again a valid A-operand (k = reflector index),
//   * A^T += (-W^T) V^T : accumulates in place.
// No fragment ever has to be transposed or exchanged between lanes.
#pragma once

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  return x;
}

// squared norm of the 8 column entries a thread holds, summed over the quad that shares the column
__device__ __forceinline__ double quad_norm8(const double (&P)[4][2]) {
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int rt = 0; rt < 4; ++rt) {
    s0 = fma(P[rt][0], P[rt][0], s0);
    s1 = fma(P[rt][1], P[rt][1], s1);
  }
  return quad_sum(s0 + s1);
}

template <int NCT, int MINB>
__global__ void __launch_bounds__(32, MINB) k_qr_mma(QrArgs a) {
  constexpr int NRT = 4, ROWS = 8 * NRT, VS = 34;   // VS: row stride of Vt (bank spread for the fragment loads)
  extern __shared__ double sm[];
  const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  const int n = a.n, nc1 = n + 1, ldo = a.ldo;
  const int npk = n * (n + 3) / 2;
  double* Rp = sm;                            // packed [R | Q^T r], rows 0..n-1, cols j..n
  double* Vt = Rp + npk + 2 - (npk & 1);      // [8][VS]  reflector vectors of the current panel, transposed
  double* Gs = Vt + 8 * VS;                   // [8][8]   V^T V (strict upper part used)
  double* Ts = Gs + 64;                       // [8][8]   -T
  double* tau_s = Ts + 64;                    // [8]
  double* amb_s = tau_s + 8;                  // [8]      alpha - beta of each reflector (its entry in the R row)
  int* rowstart = reinterpret_cast<int*>(amb_s + 10);   // amb_s[8] is the scratch slot of the panel steps
  __shared__ int s_total, s_f0;
  for (int t = lane; t < npk; t += 32) Rp[t] = 0.0;
  const double* src_base;
  if (a.src_mode == 0) {
    const int f0 = (int)((long)a.F * part / nparts), f1 = (int)((long)a.F * (part + 1) / nparts);
    if (lane == 0) {
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
    if (lane == 0) { s_total = a.dense_rows; s_f0 = 0; }
    src_base = a.dense + (size_t)b * a.dense_stride;
  }
  __syncwarp();
  const int total = s_total;
  const int nfr = (a.src_mode == 0) ? ((int)((long)a.F * (part + 1) / nparts) - s_f0) : 0;
  auto off = [&](int j) { return j * nc1 - (j * (j - 1)) / 2; };  // packed offset of R[j][j]
  auto resolve = [&](int v) -> long {   // element offset of virtual row v (-1: none)
    if (v >= total) return -1;
    if (a.src_mode != 0) return (long)v * ldo;
    int lo = 0, hi = nfr;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rowstart[mid] <= v) lo = mid; else hi = mid;
    }
    return ((long)(s_f0 + lo) * a.qmax + (v - rowstart[lo])) * ldo;
  };

  double At[NCT][NRT][2];
  for (int base = 0; base < total; base += ROWS) {
    // ---- load the chunk into the fragment layout ------------------------------------------------------
    {
      const long phys = resolve(base + lane);   // lane r resolves chunk row r
      const long nx = resolve(base + ROWS + lane);
      if (nx >= 0) {  // pull the next chunk towards L2 while this one is being eliminated
        const char* p = reinterpret_cast<const char*>(src_base + nx);
        for (int o = 0; o < nc1 * 8; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
      }
      long po[NRT][2];
#pragma unroll
      for (int rt = 0; rt < NRT; ++rt)
#pragma unroll
        for (int par = 0; par < 2; ++par) po[rt][par] = __shfl_sync(0xffffffffu, phys, 8 * rt + 2 * q + par);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) {
        const int col = 8 * ct + g;
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt)
#pragma unroll
          for (int par = 0; par < 2; ++par)
            At[ct][rt][par] = (po[rt][par] >= 0 && col < nc1) ? __ldg(src_base + po[rt][par] + col) : 0.0;
      }
    }
#pragma unroll
    for (int p = 0; p < NCT; ++p) {
      if (8 * p >= n) break;
      // ================= panel factorisation: columns 8p .. 8p+7, quad g owns column 8p+g ==============
      const int colp = 8 * p + g;
      const int nj = min(8, n - 8 * p);   // reflectors in this panel (the residual column is never eliminated)
      if (lane < 8) { tau_s[lane] = 0.0; amb_s[lane] = 0.0; }
      Gs[lane] = 0.0;
      Gs[lane + 32] = 0.0;
      if (g >= nj) {
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) *reinterpret_cast<double2*>(Vt + g * VS + 8 * rt + 2 * q) = make_double2(0.0, 0.0);
      }
      // squared norm of the own column below the diagonal, kept per quad. It is computed exactly here and
      // whenever it is flagged stale (encoded as a negative value); in between it is DOWNDATED from
      // quantities the step has anyway (||a - w v||^2 = ||a||^2 - w (2 v.a - w ||v||^2)), which takes two
      // quad-shuffle rounds and eight FMAs off the serial chain of every reflector. `big` bounds the
      // magnitudes that entered the downdates since the last exact value: absolute error <= ~30 ulp(big),
      // so a result below 1e-3 big (cancellation) is recomputed exactly before it is used (cf. LAPACK
      // dgeqp3's safeguard). The step body is branch-free apart from that rare uniform recompute: a
      // column with nothing below the diagonal gets tau = 0 and alpha - beta = 0 (H = I) by selects.
      double ss = quad_norm8(At[p]), big = ss;
      const int dump = (int)(amb_s - sm) + 8;        // scratch slot for lanes with nothing to store
      int ivrow = (int)(Vt - sm) + 2 * q;            // Vt[jj][2q]
      int irj = off(8 * p);                          // R[j][j]
      int igc = (int)(Gs - sm) + g * 8;              // G[g][jj]
#pragma unroll 1
      for (int jj = 0; jj < nj; ++jj) {
        const int j = 8 * p + jj;
        const double alpha = sm[irj];
        double sigma = __shfl_sync(0xffffffffu, ss, 4 * jj);
        if (sigma < 0.0) {   // stale (uniform, rare)
          ss = quad_norm8(At[p]);
          big = ss;
          sigma = __shfl_sync(0xffffffffu, ss, 4 * jj);
        }
        const bool live = sigma >= 1e-290;
        const bool trail = (g > jj) && (colp < nc1);
        const int irjk = irj + (g - jj);
        const double rjk = trail ? sm[irjk] : 0.0;   // read by the whole quad before the shuffles below
        // Householder with the un-normalised vector u = [alpha - beta ; v] (see k_qr_compress)
        const double nrm2 = live ? fma(alpha, alpha, sigma) : 1.0;
        const double absb = nrm2 * rsqrt_nobranch(nrm2);
        const double beta = live ? -copysign(absb, alpha) : alpha;
        const double amb = alpha - beta;
        const double taup = (live ? 1.0 : 0.0) * rcp_short(fma(fabs(alpha), absb, nrm2));   // (no branch)
        if (g == jj) {
#pragma unroll
          for (int rt = 0; rt < NRT; ++rt)
            *reinterpret_cast<double2*>(sm + ivrow + 8 * rt) = make_double2(At[p][rt][0], At[p][rt][1]);
        }
        __syncwarp();
        double v[NRT][2];
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          const double2 t = *reinterpret_cast<const double2*>(sm + ivrow + 8 * rt);
          v[rt][0] = t.x;
          v[rt][1] = t.y;
        }
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          d0 = fma(v[rt][0], At[p][rt][0], d0);
          d1 = fma(v[rt][1], At[p][rt][1], d1);
        }
        const double d = quad_sum(d0 + d1);      // quads g > jj: v . a_g ; quads g < jj: v_g . v_jj
        const double wk = trail ? taup * fma(amb, rjk, d) : 0.0;
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          At[p][rt][0] = fma(-wk, v[rt][0], At[p][rt][0]);
          At[p][rt][1] = fma(-wk, v[rt][1], At[p][rt][1]);
        }
        {
          big = fmax(big, fma(wk * wk, sigma, ss));
          const double sn = fma(-wk, fma(-wk, sigma, d + d), ss);
          ss = (ss < 0.0 || sn < 1e-3 * big) ? ((big > 0.0) ? -1.0 : 0.0) : sn;   // negative = stale, sticky
        }
        // exactly one store per lane: q == 0 lanes write R[j][col] (g > jj) or G[g][jj] (g < jj); lanes
        // 1, 5, 9 write beta, tau and alpha - beta; everybody else hits the scratch slot
        {
          int ist = dump;
          double val = 0.0;
          if (q == 0 && g < jj) { ist = igc; val = d; }
          if (q == 0 && trail) { ist = irjk; val = fma(-wk, amb, rjk); }
          if (lane == 1) { ist = irj; val = beta; }
          if (lane == 5) { ist = (int)(tau_s - sm) + jj; val = taup; }
          if (lane == 9) { ist = (int)(amb_s - sm) + jj; val = amb; }
          sm[ist] = val;
        }
        ivrow += VS;
        irj += nc1 - j;
        igc += 1;
      }
      __syncwarp();
      if (p + 1 < NCT) {
        // ================= T: H_0 ... H_7 = I - U T U^T, lane i (mod 8) builds row i ==================
        {
          const int ti = lane & 7;
          double Trow[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < k; ++l) acc = fma(Trow[l], Gs[l * 8 + k], acc);
            const double tk = tau_s[k];
            Trow[k] = (k == ti) ? tk : ((k > ti) ? -tk * acc : 0.0);
          }
          if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(Ts + ti * 8 + k) = make_double2(-Trow[k], -Trow[k + 1]);
          }
        }
        __syncwarp();
        double TBn[2], ambq[2], VB1[NRT][2], VB2[NRT][2];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          TBn[par] = Ts[(2 * q + par) * 8 + g];
          ambq[par] = amb_s[2 * q + par];
        }
#pragma unroll
        for (int rt = 0; rt < NRT; ++rt) {
          const double2 t = *reinterpret_cast<const double2*>(Vt + g * VS + 8 * rt + 2 * q);
          VB1[rt][0] = t.x;                                   // V[8rt + 2q + par][g]
          VB1[rt][1] = t.y;
          VB2[rt][0] = Vt[(2 * q) * VS + 8 * rt + g];         // V[8rt + g][2q + par]
          VB2[rt][1] = Vt[(2 * q + 1) * VS + 8 * rt + g];
        }
        // ================= trailing update of the column tiles to the right ==========================
#pragma unroll
        for (int ct = p + 1; ct < NCT; ++ct) {
          const int col = 8 * ct + g;
          double rold[2], Za[2], Zb[2] = {0.0, 0.0};
          int idx[2];
          bool valid[2];
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            const int rowR = 8 * p + 2 * q + par;
            valid[par] = (col < nc1) && (rowR < n);
            idx[par] = off(rowR) + (col - rowR);
            rold[par] = valid[par] ? Rp[idx[par]] : 0.0;
            Za[par] = ambq[par] * rold[par];
          }
#pragma unroll
          for (int rt = 0; rt < NRT; rt += 2) {
            dmma884(Za[0], Za[1], At[ct][rt][0], VB1[rt][0]);
            dmma884(Zb[0], Zb[1], At[ct][rt + 1][0], VB1[rt + 1][0]);
            dmma884(Za[0], Za[1], At[ct][rt][1], VB1[rt][1]);
            dmma884(Zb[0], Zb[1], At[ct][rt + 1][1], VB1[rt + 1][1]);
          }
          const double z0 = Za[0] + Zb[0], z1 = Za[1] + Zb[1];
          double w0 = 0.0, w1 = 0.0;                          // -(Z^T T)[col g][reflector 2q + {0,1}]
          dmma884(w0, w1, z0, TBn[0]);
          dmma884(w0, w1, z1, TBn[1]);
          if (valid[0]) Rp[idx[0]] = fma(ambq[0], w0, rold[0]);
          if (valid[1]) Rp[idx[1]] = fma(ambq[1], w1, rold[1]);
#pragma unroll
          for (int rt = 0; rt < NRT; ++rt) {
            dmma884(At[ct][rt][0], At[ct][rt][1], w0, VB2[rt][0]);
            dmma884(At[ct][rt][0], At[ct][rt][1], w1, VB2[rt][1]);
          }
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // ---- write [R | Q^T r] as n x (n+1) row-major, zeros below the diagonal ---------------------------
  double* out = a.out + (size_t)b * a.out_stride + (size_t)part * n * nc1;
  for (int j = 0; j < n; ++j) {
    const int oj = off(j);
    for (int k = lane; k < nc1; k += 32) out[(size_t)j * nc1 + k] = (k >= j) ? Rp[oj + (k - j)] : 0.0;
  }
}

template <int NCT, int MINB>
void launch_mma(const QrArgs& a, int split, int B, int max_frange, cudaStream_t st) {
  const int n = a.n;
  const size_t npk = (size_t)n * (n + 3) / 2;
  size_t smem = sizeof(double) * (npk + 2 + 8 * 34 + 64 + 64 + 18) + sizeof(int) * (max_frange + 2);
  IGV_SMEM_OPTIN((k_qr_mma<NCT, MINB>), 220 * 1024);
  dim3 grid(split, B);
  k_qr_mma<NCT, MINB><<<grid, 32, smem, st>>>(a);
}
