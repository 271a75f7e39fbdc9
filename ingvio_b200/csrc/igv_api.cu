// C-ABI entry points (include/ingvio_b200.h): handle lifetime, variable layout bookkeeping
// (the host-side mirror of State::_err_variables, State.h:129-135), staging of host arguments and
// kernel sequencing.  No torch types, no exit(): every failure is an igv_status + igv_last_error.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "igv_internal.h"

namespace {

#define IGV_CUDA(h, call)                                                            \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
      return IGV_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

igv_status fail(igv_batch* h, igv_status s, const char* msg) {
  if (h) h->err = msg;
  return s;
}

igv_status check_launch(igv_batch* h) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string("kernel launch: ") + cudaGetErrorString(e);
    return IGV_ERR_CUDA;
  }
  return IGV_OK;
}

// Reserve `bytes` in the current staging slot. On growth the old block is RETIRED, not freed: pointers handed
// out earlier in the same API call (and copies / kernels already enqueued on them) stay valid.
igv_status arena_reserve(igv_batch* h, size_t bytes, char** out) {
  igv_batch::Slot& s = h->slots[h->slot];
  bytes = (bytes + 255) & ~size_t(255);
  if (s.off + bytes > s.cap) {
    // every slot converges to the same capacity (the largest call seen so far), so a steady-state frame loop stops
    // allocating after one trip around the ring
    const size_t ncap = std::max(std::max(s.cap * 2, bytes + (size_t(4) << 20)), h->slot_cap);
    h->slot_cap = ncap;
    char* n = nullptr;
    IGV_CUDA(h, cudaMalloc(&n, ncap));
    if (s.mem) h->retired.push_back(s.mem);   // kept until igv_destroy: cudaFree would synchronise the device
    s.mem = n;
    s.cap = ncap;
    s.off = 0;
    // bring every other slot to the new capacity NOW (their old blocks are retired, so work in flight on them
    // stays valid): all allocation happens inside the first large call instead of trickling into later frames
    for (auto& o : h->slots) {
      if (&o == &s || o.cap >= ncap) continue;
      char* m = nullptr;
      if (cudaMalloc(&m, ncap) != cudaSuccess) { cudaGetLastError(); continue; }
      if (o.mem) h->retired.push_back(o.mem);
      o.mem = m;
      o.cap = ncap;
      // `off` is untouched: a slot is only ever bumped by the call that owns it, after arena_reset zeroed it
    }
  }
  *out = s.mem + s.off;
  s.off += bytes;
  s.used = true;
  return IGV_OK;
}

// Start of an API call: every kernel of the previous call has been enqueued, so its slot is marked consumed once
// they finish; the call gets the next slot of the ring, whose earlier readers the copy stream waits for.
void arena_reset(igv_batch* h) {
  igv_commit_copies(h);
  igv_batch::Slot& prev = h->slots[h->slot];
  // (while igv_frame_step captures a graph no event of the staging ring may be touched: an event recorded inside a
  // capture belongs to the graph; igv_arena_quiesce recorded it just before the capture began)
  if (prev.used && prev.consumed && !h->capturing) cudaEventRecord(prev.consumed, h->stream);
  h->slot = (h->slot + 1) % igv_batch::kSlots;
  igv_batch::Slot& s = h->slots[h->slot];
  if (s.used && s.consumed && h->copy_stream && !h->capturing) cudaStreamWaitEvent(h->copy_stream, s.consumed, 0);
  s.off = 0;
  s.used = false;
  if (s.cap < h->slot_cap) {   // converge to the common capacity right away (the old block is retired, never freed here)
    char* n = nullptr;
    if (cudaMalloc(&n, h->slot_cap) == cudaSuccess) {
      if (s.mem) h->retired.push_back(s.mem);
      s.mem = n;
      s.cap = h->slot_cap;
    } else {
      cudaGetLastError();
    }
  }
}

// Device view of a bulk argument: the pointer itself in DEVICE mode, else a copy in the staging slot, made on the
// copy stream (the compute stream is ordered after all of a call's copies by ONE event, igv_commit_copies).
template <class T>
igv_status stage(igv_batch* h, const T* src, size_t count, const T** out) {
  if (!src) { *out = nullptr; return IGV_OK; }
  if (h->ptr_mode == IGV_PTR_DEVICE) { *out = src; return IGV_OK; }
  char* mem = nullptr;
  igv_status s = arena_reserve(h, count * sizeof(T), &mem);
  if (s != IGV_OK) return s;
  T* dst = reinterpret_cast<T*>(mem);
  if (h->copy_stream) {
    IGV_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, h->copy_stream));
    h->copies_pending = true;   // igv_commit_copies orders the compute stream after it before the first kernel
  } else {
    IGV_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  }
  *out = dst;
  return IGV_OK;
}

// Device buffer for an output argument; fetch() copies it back in HOST mode.
template <class T>
igv_status out_buf(igv_batch* h, T* user, size_t count, T** dev) {
  if (!user) { *dev = nullptr; return IGV_OK; }
  if (h->ptr_mode == IGV_PTR_DEVICE) { *dev = user; return IGV_OK; }
  char* mem = nullptr;
  igv_status s = arena_reserve(h, count * sizeof(T), &mem);
  if (s != IGV_OK) return s;
  *dev = reinterpret_cast<T*>(mem);
  return IGV_OK;
}
template <class T>
igv_status fetch(igv_batch* h, T* user, const T* dev, size_t count) {
  if (!user || h->ptr_mode == IGV_PTR_DEVICE) return IGV_OK;
  IGV_CUDA(h, cudaMemcpyAsync(user, dev, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

#define IGV_TRY(expr)                 \
  do {                                \
    igv_status s_ = (expr);           \
    if (s_ != IGV_OK) return s_;      \
  } while (0)

igv_status make_blocks(igv_batch* h, int n_blocks, const int* idx, const int* size, IgvBlocks* out) {
  if (n_blocks <= 0 || n_blocks > IGV_MAX_BLOCKS || !idx || !size) return fail(h, IGV_ERR_INVALID, "bad block list");
  out->n_blocks = n_blocks;
  out->n = 0;
  for (int i = 0; i < n_blocks; ++i) {
    // checkSubOrder (StateManager.cpp:428-445): every block must be a variable of the state
    bool found = false;
    for (const auto& v : h->vars) if (v.idx == idx[i] && v.size == size[i]) found = true;
    if (!found) return fail(h, IGV_ERR_STATE, "var_order block is not a variable of the state");
    out->idx[i] = idx[i];
    out->size[i] = size[i];
    out->n += size[i];
  }
  if (out->n > 6 * IGV_MAX_BLOCKS) return fail(h, IGV_ERR_CAPACITY, "too many measured columns");
  return IGV_OK;
}

void reindex(igv_batch* h) {
  int idx = 0;
  for (auto& v : h->vars) { v.idx = idx; idx += v.size; }
  h->N = idx;
}

// ---- track-table column bookkeeping (igv_tracks_*): window slot -> physical column, following the clone list ----
void trk_on_augment(igv_batch* h) {
  if (h->trk.T == 0) return;
  igv_trk_col_alloc(h->trk);   // never -1: the window holds at most cfg.max_clones = C clones
}
void trk_on_marg_clone(igv_batch* h, int slot) {
  if (h->trk.T == 0 || slot < 0 || slot >= (int)h->trk.col_of_slot.size()) return;
  // Observations at a clone that left the window cannot be used by anything; with the reference's call order
  // (clean*ObsAtMargTime -> changeMSCKFAnchor -> margSwPose) the column is already empty and this is a no-op.
  igv_launch_trk_clean(h, 1ull << h->trk.col_of_slot[slot], 0);
  igv_trk_col_release(h->trk, slot);
}
igv_status trk_ready(igv_batch* h) {
  if (!h) return IGV_ERR_INVALID;
  if (h->trk.T == 0) return fail(h, IGV_ERR_STATE, "track table not created (igv_tracks_create)");
  return IGV_OK;
}
igv_status trk_slot_bits(igv_batch* h, int n, const int* slots, unsigned long long* bits) {
  *bits = 0ull;
  if (n < 0 || (n > 0 && !slots)) return IGV_ERR_INVALID;
  if (!igv_trk_slot_bits(h->trk, n, slots, bits)) return fail(h, IGV_ERR_STATE, "clone slot not in the sliding window");
  return IGV_OK;
}

// Regularised lower incomplete gamma P(a, x): series below a + 1, Lentz continued fraction above.
double reg_gamma_p(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  const double pre = std::exp(a * std::log(x) - x - std::lgamma(a));
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term;
    for (int k = 1; k < 2000; ++k) {
      term *= x / (a + k);
      sum += term;
      if (term < sum * 1e-17) break;
    }
    return pre * sum;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, f = d;
  for (int k = 1; k < 2000; ++k) {
    const double an = -k * (k - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double delta = c * d;
    f *= delta;
    if (std::fabs(delta - 1.0) < 1e-16) break;
  }
  return 1.0 - pre * f;
}

}  // namespace

// Everything enqueued so far that reads the current staging slot is marked consumed NOW (eagerly, outside any capture).
void igv_arena_quiesce(igv_batch* h) {
  igv_commit_copies(h);
  igv_batch::Slot& prev = h->slots[h->slot];
  if (prev.used && prev.consumed) cudaEventRecord(prev.consumed, h->stream);
}

IgvLayout igv_batch::layout() const {
  IgvLayout L;
  L.N = N; L.ld = ld; L.xsize = xsize; L.n_clones = 0;
  L.n_lm = 0; L.lm_off = IGV_X_CORE + 12 * cfg.max_clones;
  for (int i = 0; i < 6; ++i) L.idx_gnss[i] = -1;
  for (int i = 0; i < IGV_MAX_CLONES; ++i) L.idx_clone[i] = -1;
  for (int i = 0; i < IGV_MAX_LM; ++i) { L.idx_lm[i] = -1; L.lm_anchor[i] = -1; }
  for (const auto& v : vars) {
    if (v.kind == VK_GNSS) L.idx_gnss[v.tag] = v.idx;
    if (v.kind == VK_CLONE && L.n_clones < IGV_MAX_CLONES) L.idx_clone[L.n_clones++] = v.idx;
    if (v.kind == VK_LANDMARK && L.n_lm < IGV_MAX_LM) { L.idx_lm[L.n_lm] = v.idx; L.lm_anchor[L.n_lm] = v.tag; ++L.n_lm; }
  }
  return L;
}

extern "C" {

igv_status igv_create(const igv_config* cfg, igv_batch** out) {
  if (!cfg || !out) return IGV_ERR_INVALID;
  *out = nullptr;
  if (cfg->batch < 1 || cfg->max_dim < 21 || cfg->max_dim > 512 || cfg->max_clones < 0 ||
      cfg->max_clones > IGV_MAX_CLONES || cfg->max_feats < 0 || cfg->max_sats < 0 || cfg->max_sats > 64 ||
      cfg->max_landmarks < 0 || cfg->max_landmarks > IGV_MAX_LM)
    return IGV_ERR_INVALID;
  igv_batch* h = new igv_batch();
  h->cfg = *cfg;
  h->B = cfg->batch;
  h->ld = (cfg->max_dim + 1) & ~1;
  h->xsize = IGV_X_CORE + 12 * cfg->max_clones + 3 * cfg->max_landmarks;
  h->rho = cfg->stereo ? 4 : 2;
  h->ncols_max = 6 * std::max(1, cfg->max_clones);
  h->qmax = std::max(1, h->rho * cfg->max_clones - 3);
  h->max_rows = std::max(std::max(h->ncols_max, 2 * cfg->max_sats), 32);
  auto bail = [&](const char* what, cudaError_t e) {
    std::fprintf(stderr, "igv_create: %s: %s\n", what, cudaGetErrorString(e));
    igv_destroy(h);
    return IGV_ERR_CUDA;
  };
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return bail("cudaSetDevice", e);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore_{prev_dev == cfg->device ? -1 : prev_dev};
  {
    auto knob = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
    h->knobs.fuse = knob("IGV_FUSE", -1);
    h->knobs.qr_cfg = knob("IGV_QR_CFG", 0);
    h->knobs.qr_split = knob("IGV_QR_SPLIT", 0);
    h->knobs.gram_cfg = knob("IGV_GRAM_CFG", 0);
    h->knobs.factor_cfg = knob("IGV_FACTOR_CFG", 0);
    h->knobs.tri_cfg = knob("IGV_TRI_CFG", 0);
    h->knobs.fuse_minw = knob("IGV_FUSE_MINW", 12);
    h->knobs.tri_minb = knob("IGV_TRI_MINB", 4);
    h->knobs.graph = knob("IGV_GRAPH", -1);
    h->knobs.feat_const = knob("IGV_FEAT_CONST", 1);
    h->knobs.prop_tma = knob("IGV_PROP_TMA", 1);
    h->knobs.feat_warps = knob("IGV_FEAT_WARPS", 0);
    h->knobs.feat_ps = knob("IGV_FEAT_PS", 1);
    h->knobs.tc_drain = knob("IGV_TC_DRAIN", 0);
    if (const char* e = std::getenv("IGV_TC_PIVOT_TOL")) h->knobs.tc_pivot_tol = std::atof(e);
    h->knobs.ekf_t_small = knob("IGV_EKF_T_SMALL", 0);
    h->knobs.ekf_t_big = knob("IGV_EKF_T_BIG", 0);
  }
  if (cfg->stream) {
    h->stream = static_cast<cudaStream_t>(cfg->stream);
  } else {
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail("cudaStreamCreate", e);
    h->own_stream = true;
  }
  const size_t B = h->B;
#define IGV_ALLOC(ptr, count)                                                              \
  do {                                                                                     \
    e = cudaMalloc(reinterpret_cast<void**>(&(ptr)), sizeof(*(ptr)) * (size_t)(count));    \
    if (e != cudaSuccess) return bail(#ptr, e);                                            \
    e = cudaMemsetAsync((ptr), 0, sizeof(*(ptr)) * (size_t)(count), h->stream);            \
    if (e != cudaSuccess) return bail(#ptr, e);                                            \
  } while (0)
  e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return bail("copy stream", e);
  e = cudaEventCreateWithFlags(&h->copied, cudaEventDisableTiming);
  if (e != cudaSuccess) return bail("event", e);
  for (auto& sl : h->slots) {
    e = cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming);
    if (e != cudaSuccess) return bail("event", e);
  }
  for (auto& f : h->fences) {
    e = cudaEventCreateWithFlags(&f, cudaEventDisableTiming);
    if (e != cudaSuccess) return bail("event", e);
  }
  IGV_ALLOC(h->P[0], B * h->ld * h->ld);
  IGV_ALLOC(h->P[1], B * h->ld * h->ld);
  IGV_ALLOC(h->X[0], B * h->xsize);
  IGV_ALLOC(h->X[1], B * h->xsize);
  IGV_ALLOC(h->flags, B);
  IGV_ALLOC(h->chi2, 1024);
  IGV_ALLOC(h->chi2_095, 128);
  {
    double tab[128];
    for (int d = 1; d <= 128; ++d) tab[d - 1] = igv_chi2_quantile(0.95, d);
    e = cudaMemcpyAsync(h->chi2_095, tab, sizeof(tab), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);   // `tab` is a stack array
    if (e != cudaSuccess) return bail("chi2_095", e);
  }
  const size_t F = std::max(1, cfg->max_feats);
  IGV_ALLOC(h->Hs, B * F * h->qmax * (h->ncols_max + 1));
  IGV_ALLOC(h->f_rows, B * F);
  IGV_ALLOC(h->f_gamma, B * F);
  IGV_ALLOC(h->n_acc, B);
  IGV_ALLOC(h->Hc, B * h->ncols_max * (h->ncols_max + 1));
  h->qr_split_cap = (h->B < 296) ? 16 : 1;
  if (h->qr_split_cap > 1) IGV_ALLOC(h->Rpart, B * h->qr_split_cap * h->ncols_max * (h->ncols_max + 1));
  h->gram_n1p = igv_gram_n1p(h->ncols_max);
  IGV_ALLOC(h->Gws, B * h->qr_split_cap * h->gram_n1p * h->gram_n1p);
  IGV_ALLOC(h->Zws, B * h->max_rows * (h->ld + 1));
  IGV_ALLOC(h->Sws, B * h->max_rows * h->max_rows);
  IGV_ALLOC(h->dxws, B * h->ld);
  const size_t S2 = std::max(2, 2 * cfg->max_sats);
  IGV_ALLOC(h->Hg, B * S2 * 16);
  IGV_ALLOC(h->rg, B * S2);
  IGV_ALLOC(h->Rg, B * S2);
  IGV_ALLOC(h->cnt_g, B);
  IGV_ALLOC(h->gam_ws, B);
  if (cfg->max_landmarks > 0) {   // landmark initialisation measures every clone: rows rho * clones, columns 6 * clones
    h->dws_rows = std::max(128, h->rho * cfg->max_clones);
    h->dws_cols = std::max(16, 6 * cfg->max_clones);
  }
  h->dws_stride = (size_t)h->dws_rows * (h->dws_cols + 2) + 16;
  IGV_ALLOC(h->Dws, B * h->dws_stride);
  if (cfg->max_landmarks > 0)
    IGV_ALLOC(h->Lws, B * (size_t)(2 * cfg->max_landmarks) * (15 + 6 * cfg->max_clones + 3 * cfg->max_landmarks + 2));
#undef IGV_ALLOC
  // defaults of StateParams (State.h:48-53)
  h->params.noise_g = 0.005; h->params.noise_a = 0.05; h->params.noise_bg = 0.001; h->params.noise_ba = 0.01;
  h->params.noise_cb = 2.0; h->params.noise_cb_rw = 0.2;
  h->params.g[0] = 0; h->params.g[1] = 0; h->params.g[2] = -9.8;
  for (int i = 0; i < 9; ++i) h->params.Rc[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 3; ++i) h->params.pc[i] = 0.0;
  *out = h;
  return IGV_OK;
}

igv_status igv_destroy(igv_batch* h) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_OK;
  if (h->stream) cudaStreamSynchronize(h->stream);
  igv_frame_graphs_destroy(h);
  void* ptrs[] = {h->P[0], h->P[1], h->X[0], h->X[1], h->flags, h->chi2, h->chi2_095, h->Hs, h->f_rows, h->f_gamma, h->n_acc,
                  h->Hc, h->Rpart, h->Gws, h->Zws, h->Sws, h->dxws, h->Hg, h->rg, h->Rg, h->cnt_g, h->gam_ws, h->Dws, h->Lws, h->pre_ws};
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (void* p : ptrs) if (p) cudaFree(p);
  void* tptrs[] = {h->trk.id, h->trk.mask, h->trk.st, h->trk.anchor, h->trk.pf, h->trk.pf_fej, h->trk.obs};
  for (void* p : tptrs) if (p) cudaFree(p);
  for (auto& sl : h->slots) { if (sl.mem) cudaFree(sl.mem); if (sl.consumed) cudaEventDestroy(sl.consumed); }
  for (auto& f : h->fences) if (f) cudaEventDestroy(f);
  if (h->copied) cudaEventDestroy(h->copied);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (char* p : h->retired) cudaFree(p);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return IGV_OK;
}

const char* igv_last_error(const igv_batch* h) { return h ? h->err.c_str() : "null handle"; }

igv_status igv_set_pointer_mode(igv_batch* h, int mode) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || (mode != IGV_PTR_HOST && mode != IGV_PTR_DEVICE)) return IGV_ERR_INVALID;
  h->ptr_mode = mode;
  return IGV_OK;
}

igv_status igv_set_compression(igv_batch* h, int kind) {
  IgvDeviceGuard dev_guard_(h);
  if (h) h->cfg_version++;
  if (!h || kind < IGV_COMPRESS_AUTO || kind > IGV_COMPRESS_GRAM) return IGV_ERR_INVALID;
  h->compress = kind;
  return IGV_OK;
}

igv_status igv_set_precision(igv_batch* h, int mode) {
  if (h) h->cfg_version++;
  if (!h || (mode != IGV_PREC_FP64 && mode != IGV_PREC_FP32_STACK && mode != IGV_PREC_TF32_GRAM)) return IGV_ERR_INVALID;
  h->stack_f32 = (mode != IGV_PREC_FP64) ? 1 : 0;
  h->gram_tc = (mode == IGV_PREC_TF32_GRAM) ? 1 : 0;
  return IGV_OK;
}

int igv_last_gram_tensor(const igv_batch* h) { return h ? h->last_gram_tc : -1; }

int igv_last_visual_path(const igv_batch* h) { return h ? h->last_visual_path : -1; }

igv_status igv_synchronize(igv_batch* h) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

long long igv_launch_count(const igv_batch* h) { return h ? h->launches : 0; }

igv_status igv_set_params(igv_batch* h, const igv_params* p) {
  IgvDeviceGuard dev_guard_(h);
  if (h) h->cfg_version++;
  if (!h || !p) return IGV_ERR_INVALID;
  h->params.noise_g = p->noise_g; h->params.noise_a = p->noise_a;
  h->params.noise_bg = p->noise_bg; h->params.noise_ba = p->noise_ba;
  h->params.noise_cb = p->noise_clockbias; h->params.noise_cb_rw = p->noise_cb_rw;
  for (int i = 0; i < 3; ++i) { h->params.g[i] = p->gravity[i]; h->params.pc[i] = p->T_cl2cr_p[i]; }
  for (int i = 0; i < 9; ++i) h->params.Rc[i] = p->T_cl2cr_R[i];
  return IGV_OK;
}

// chi^2 quantile by safeguarded Newton on P(dof/2, x/2) = p, started from the Wilson-Hilferty cube
// (replaces boost::math::quantile(chi_squared(dof), p), Update.cpp:31-32, StateManager.cpp:613-615).
double igv_chi2_quantile(double p, int dof) {
  if (!(p > 0.0 && p < 1.0) || dof < 1) return NAN;
  const double a = 0.5 * dof;
  // normal quantile for the start: bisection on erfc is plenty (only a starting point)
  double zlo = -8.0, zhi = 8.0;
  for (int it = 0; it < 60; ++it) {
    const double zm = 0.5 * (zlo + zhi);
    if (0.5 * std::erfc(-zm / std::sqrt(2.0)) < p) zlo = zm; else zhi = zm;
  }
  const double z = 0.5 * (zlo + zhi), t = 2.0 / (9.0 * dof);
  double x = dof * std::pow(1.0 - t + z * std::sqrt(t), 3.0);
  if (!(x > 0.0)) x = a;
  double lo = 0.0, hi = std::max(2.0 * x, 4.0 * dof + 60.0);
  for (int it = 0; it < 300; ++it) {
    const double f = reg_gamma_p(a, 0.5 * x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = 0.5 * std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a));
    double xn = x - f / dens;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    const bool done = std::fabs(xn - x) <= 2e-15 * std::max(1.0, x);
    x = xn;
    if (done) break;
  }
  return x;
}

igv_status igv_set_chi2_table(igv_batch* h, const double* table, int max_dof) {
  IgvDeviceGuard dev_guard_(h);
  if (h) h->cfg_version++;
  if (!h || !table || max_dof < 1 || max_dof > 1024) return IGV_ERR_INVALID;
  IGV_CUDA(h, cudaMemcpyAsync(h->chi2, table, sizeof(double) * max_dof, cudaMemcpyHostToDevice, h->stream));
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  h->chi2_n = max_dof;
  h->chi2_host.assign(table, table + max_dof);
  return IGV_OK;
}

// ---- state ----------------------------------------------------------------------------------------
igv_status igv_state_init(igv_batch* h, const double* R_i2w, const double* p, const double* v, const double* bg,
                          const double* ba, const double* R_ext, const double* p_ext, const double* cov_diag21) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !R_i2w || !p || !v || !bg || !ba || !R_ext || !p_ext || !cov_diag21) return IGV_ERR_INVALID;
  arena_reset(h);
  const size_t B = h->B;
  const double *dR, *dp, *dv, *dbg, *dba, *dRe, *dpe, *dd;
  IGV_TRY(stage(h, R_i2w, B * 9, &dR)); IGV_TRY(stage(h, p, B * 3, &dp)); IGV_TRY(stage(h, v, B * 3, &dv));
  IGV_TRY(stage(h, bg, B * 3, &dbg)); IGV_TRY(stage(h, ba, B * 3, &dba));
  IGV_TRY(stage(h, R_ext, B * 9, &dRe)); IGV_TRY(stage(h, p_ext, B * 3, &dpe));
  // cov_diag is shared and tiny: always a host array
  double* ddiag = nullptr;
  {
    const int pm = h->ptr_mode;
    h->ptr_mode = IGV_PTR_HOST;
    igv_status s = stage(h, cov_diag21, 21, &dd);
    h->ptr_mode = pm;
    if (s != IGV_OK) return s;
    (void)ddiag;
  }
  h->vars.clear();
  h->vars.push_back({VK_SE23, 0, 9, 0});
  h->vars.push_back({VK_BG, 9, 3, 0});
  h->vars.push_back({VK_BA, 12, 3, 0});
  h->vars.push_back({VK_EXT, 15, 6, 0});
  reindex(h);
  IGV_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(int) * B, h->stream));
  igv_launch_state_init(h, dR, dp, dv, dbg, dba, dRe, dpe, dd);
  if (h->trk.T > 0) { h->trk.col_of_slot.clear(); igv_launch_trk_reset(h); }   // a new State starts with an empty map
  return check_launch(h);
}

int igv_dim(const igv_batch* h) { return h ? h->N : -1; }
int igv_num_variables(const igv_batch* h) { return h ? (int)h->vars.size() : -1; }
int igv_num_clones(const igv_batch* h) { return h ? h->layout().n_clones : -1; }
int igv_clone_idx(const igv_batch* h, int slot) {
  if (!h) return -1;
  IgvLayout L = h->layout();
  return (slot >= 0 && slot < L.n_clones) ? L.idx_clone[slot] : -1;
}
int igv_gnss_idx(const igv_batch* h, int gtype) {
  if (!h || gtype < 0 || gtype > 5) return -1;
  return h->layout().idx_gnss[gtype];
}
int igv_state_size(const igv_batch* h) { return h ? h->xsize : -1; }

igv_status igv_state_get(igv_batch* h, double* dst) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !dst) return IGV_ERR_INVALID;
  const cudaMemcpyKind k = h->ptr_mode == IGV_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  IGV_CUDA(h, cudaMemcpyAsync(dst, h->Xc(), sizeof(double) * h->B * h->xsize, k, h->stream));
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}
igv_status igv_state_get_async(igv_batch* h, double* dst) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !dst) return IGV_ERR_INVALID;
  const cudaMemcpyKind k = h->ptr_mode == IGV_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  IGV_CUDA(h, cudaMemcpyAsync(dst, h->Xc(), sizeof(double) * h->B * h->xsize, k, h->stream));
  return IGV_OK;
}
igv_status igv_state_set(igv_batch* h, const double* src) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !src) return IGV_ERR_INVALID;
  const cudaMemcpyKind k = h->ptr_mode == IGV_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  IGV_CUDA(h, cudaMemcpyAsync(h->Xc(), src, sizeof(double) * h->B * h->xsize, k, h->stream));
  return IGV_OK;
}

// ---- covariance -------------------------------------------------------------------------------------
igv_status igv_cov_get(igv_batch* h, double* dst, int ld) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !dst || ld < h->N) return IGV_ERR_INVALID;
  arena_reset(h);
  double* dev;
  IGV_TRY(out_buf(h, dst, (size_t)h->B * ld * h->N, &dev));
  igv_launch_cov_copy(h, dev, ld, true);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, dst, dev, (size_t)h->B * ld * h->N));
  if (h->ptr_mode == IGV_PTR_DEVICE) IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}
igv_status igv_cov_set(igv_batch* h, const double* src, int ld) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !src || ld < h->N) return IGV_ERR_INVALID;
  arena_reset(h);
  const double* dev;
  IGV_TRY(stage(h, src, (size_t)h->B * ld * h->N, &dev));
  igv_launch_cov_copy(h, const_cast<double*>(dev), ld, false);
  return check_launch(h);
}
igv_status igv_cov_get_blocks(igv_batch* h, int n_blocks, const int* idx, const int* size, double* dst) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !dst) return IGV_ERR_INVALID;
  arena_reset(h);
  IgvBlocks blk;
  IGV_TRY(make_blocks(h, n_blocks, idx, size, &blk));
  double* dev;
  const size_t cnt = (size_t)h->B * blk.n * blk.n;
  IGV_TRY(out_buf(h, dst, cnt, &dev));
  igv_launch_cov_blocks(h, blk, dev);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, dst, dev, cnt));
  if (h->ptr_mode == IGV_PTR_DEVICE) IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

static igv_status add_variable(igv_batch* h, IgvVarKind kind, int tag, int size, const double* cov_block_host) {
  if (h->N + size > h->cfg.max_dim) return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds max_dim");
  const double* dev;
  const int pm = h->ptr_mode;
  h->ptr_mode = IGV_PTR_HOST;
  igv_status s = stage(h, cov_block_host, (size_t)size * size, &dev);
  h->ptr_mode = pm;
  if (s != IGV_OK) return s;
  igv_launch_add_variable(h, size, dev);
  h->vars.push_back({kind, h->N, size, tag});
  reindex(h);
  return check_launch(h);
}

igv_status igv_add_gnss_variable(igv_batch* h, int gtype, const double* value, double cov) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || gtype < 0 || gtype > 5) return IGV_ERR_INVALID;
  arena_reset(h);
  if (h->layout().idx_gnss[gtype] >= 0) return fail(h, IGV_ERR_STATE, "GNSS variable already in the state");
  const double* dv;
  IGV_TRY(stage(h, value, (size_t)h->B, &dv));
  IGV_TRY(add_variable(h, VK_GNSS, gtype, 1, &cov));
  igv_launch_set_gnss_value(h, gtype, dv);
  return check_launch(h);
}

static igv_status marginalize_var(igv_batch* h, size_t vi) {
  const IgvVar v = h->vars[vi];
  if (v.kind == VK_SE23 || v.kind == VK_BG || v.kind == VK_BA || v.kind == VK_EXT)
    return fail(h, IGV_ERR_STATE, "core variables cannot be marginalised");
  int clone_slot = -1, lm_slot = -1;
  if (v.kind == VK_CLONE) {
    clone_slot = 0;
    for (size_t k = 0; k < vi; ++k) if (h->vars[k].kind == VK_CLONE) ++clone_slot;
    // a landmark must have left its anchor before the clone goes (LandmarkUpdate::changeLandmarkAnchor precedes margSwPose,
    // IngvioFilter.cpp:167-175, :193-196)
    for (const auto& w : h->vars)
      if (w.kind == VK_LANDMARK && w.tag == clone_slot) return fail(h, IGV_ERR_STATE, "a landmark is still anchored at this clone");
  }
  if (v.kind == VK_LANDMARK) {
    lm_slot = 0;
    for (size_t k = 0; k < vi; ++k) if (h->vars[k].kind == VK_LANDMARK) ++lm_slot;
  }
  igv_launch_marginalize(h, v.idx, v.size, clone_slot, lm_slot);
  if (clone_slot >= 0) {
    trk_on_marg_clone(h, clone_slot);
    for (auto& w : h->vars) if (w.kind == VK_LANDMARK && w.tag > clone_slot) --w.tag;   // later clones move down one slot
  }
  h->vars.erase(h->vars.begin() + vi);
  reindex(h);
  return check_launch(h);
}

igv_status igv_marg_gnss_variable(igv_batch* h, int gtype) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || gtype < 0 || gtype > 5) return IGV_ERR_INVALID;
  for (size_t i = 0; i < h->vars.size(); ++i)
    if (h->vars[i].kind == VK_GNSS && h->vars[i].tag == gtype) {
      igv_status s = marginalize_var(h, i);
      if (s != IGV_OK) return s;
      igv_launch_set_gnss_value(h, gtype, nullptr);   // state->_gnss.erase(gtype): mirror slot back to 0
      return check_launch(h);
    }
  return fail(h, IGV_ERR_STATE, "GNSS variable not in the state");
}

igv_status igv_add_variable_independent(igv_batch* h, int size, const double* cov_block) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || size < 1 || size > 6 || !cov_block) return IGV_ERR_INVALID;
  arena_reset(h);
  return add_variable(h, VK_OPAQUE, 0, size, cov_block);
}

igv_status igv_marginalize(igv_batch* h, int idx) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  for (size_t i = 0; i < h->vars.size(); ++i)
    if (h->vars[i].idx == idx) return marginalize_var(h, i);
  return fail(h, IGV_ERR_STATE, "Marg is not in the current state");  // StateManager.cpp:157-161
}

igv_status igv_marginalize_clone(igv_batch* h, int slot) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  int s = 0;
  for (size_t i = 0; i < h->vars.size(); ++i)
    if (h->vars[i].kind == VK_CLONE) {
      if (s == slot) return marginalize_var(h, i);
      ++s;
    }
  return fail(h, IGV_ERR_STATE, "clone slot not in the sliding window");
}

// ---- propagation ------------------------------------------------------------------------------------
igv_status igv_propagate_cov(igv_batch* h, const double* Phi, const double* G, const double* dt) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !Phi || !G || !dt) return IGV_ERR_INVALID;
  if ((size_t)20 * h->N * sizeof(double) > 96 * 1024)
    return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds the propagation strip");
  arena_reset(h);
  const double *dP, *dG, *ddt;
  IGV_TRY(stage(h, Phi, (size_t)h->B * 225, &dP));
  IGV_TRY(stage(h, G, (size_t)h->B * 180, &dG));
  IGV_TRY(stage(h, dt, (size_t)h->B, &ddt));
  igv_launch_propagate(h, 1, nullptr, nullptr, ddt, dP, dG);
  return check_launch(h);
}

igv_status igv_propagate_imu(igv_batch* h, int n_steps, const double* gyro, const double* accel, const double* dt) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || n_steps < 0 || !gyro || !accel || !dt) return IGV_ERR_INVALID;
  if (n_steps == 0) return IGV_OK;
  if ((size_t)20 * h->N * sizeof(double) > 96 * 1024)   // k_propagate's strip (checked BEFORE the mean is advanced)
    return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds the propagation strip");
  arena_reset(h);
  const double *dg, *da, *ddt;
  IGV_TRY(stage(h, gyro, (size_t)h->B * n_steps * 3, &dg));
  IGV_TRY(stage(h, accel, (size_t)h->B * n_steps * 3, &da));
  IGV_TRY(stage(h, dt, (size_t)h->B * n_steps, &ddt));
  igv_launch_propagate(h, n_steps, dg, da, ddt, nullptr, nullptr);
  return check_launch(h);
}

static igv_status augment(igv_batch* h, const double* R, const double* cR, const double* cp) {
  IgvLayout L = h->layout();
  if (L.n_clones >= h->cfg.max_clones) return fail(h, IGV_ERR_CAPACITY, "sliding window is full");
  if (h->N + 6 > h->cfg.max_dim) return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds max_dim");
  igv_launch_augment(h, R, cR, cp);
  h->vars.push_back({VK_CLONE, h->N, 6, 0});
  reindex(h);
  trk_on_augment(h);
  return check_launch(h);
}
igv_status igv_augment_clone(igv_batch* h) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  return augment(h, nullptr, nullptr, nullptr);
}
igv_status igv_augment_clone_cov(igv_batch* h, const double* R_i2w, const double* clone_R, const double* clone_p) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !R_i2w || ((clone_R == nullptr) != (clone_p == nullptr))) return IGV_ERR_INVALID;
  arena_reset(h);
  const double *dR, *dcR, *dcp;
  IGV_TRY(stage(h, R_i2w, (size_t)h->B * 9, &dR));
  IGV_TRY(stage(h, clone_R, (size_t)h->B * 9, &dcR));
  IGV_TRY(stage(h, clone_p, (size_t)h->B * 3, &dcp));
  return augment(h, dR, dcR, dcp);
}

// ---- EKF update -------------------------------------------------------------------------------------
static igv_status ekf_common(igv_batch* h, int n_blocks, const int* blk_idx, const int* blk_size, int rows,
                             const double* H, int ldh, const double* res, const double* R, int r_kind, double* dx_out,
                             double* gamma_out, bool gamma_only) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !H || !res || rows < 1 || ldh < rows) return IGV_ERR_INVALID;
  if (r_kind != IGV_R_ISO && r_kind != IGV_R_DIAG && r_kind != IGV_R_FULL) return IGV_ERR_INVALID;
  if (!R) return IGV_ERR_INVALID;
  if (rows > h->max_rows) return fail(h, IGV_ERR_CAPACITY, "rows exceed the EKF workspace (max_rows)");
  arena_reset(h);
  IgvEkfLaunch e{};
  IGV_TRY(make_blocks(h, n_blocks, blk_idx, blk_size, &e.blk));
  const size_t B = h->B;
  IGV_TRY(stage(h, H, B * ldh * e.blk.n, &e.H));
  IGV_TRY(stage(h, res, B * rows, &e.res));
  const size_t rcount = (r_kind == IGV_R_ISO) ? 1 : (r_kind == IGV_R_DIAG ? rows : (size_t)rows * rows);
  IGV_TRY(stage(h, R, B * rcount, &e.R));
  e.rows = rows; e.strideH = (long)ldh * e.blk.n; e.h_ld = ldh; e.h_rowmajor = 0;
  e.strideRes = rows; e.res_inc = 1; e.strideR = (long)rcount; e.r_kind = r_kind;
  double *ddx = nullptr, *dgam = nullptr;
  IGV_TRY(out_buf(h, dx_out, B * h->N, &ddx));
  IGV_TRY(out_buf(h, gamma_out, B, &dgam));
  e.dx_out = ddx; e.gamma_only = gamma_only ? 1 : 0; e.gamma_out = dgam; e.gate_rows = nullptr;
  e.apply_boxplus = gamma_only ? 0 : 1;
  igv_launch_ekf(h, e);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, dx_out, ddx, B * h->N));
  IGV_TRY(fetch(h, gamma_out, dgam, B));
  return IGV_OK;
}

igv_status igv_ekf_update(igv_batch* h, int n_blocks, const int* blk_idx, const int* blk_size, int rows, const double* H,
                          int ldh, const double* res, const double* R, int r_kind, double* dx_out) {
  IgvDeviceGuard dev_guard_(h);
  return ekf_common(h, n_blocks, blk_idx, blk_size, rows, H, ldh, res, R, r_kind, dx_out, nullptr, false);
}
igv_status igv_chi2_whiten(igv_batch* h, int n_blocks, const int* blk_idx, const int* blk_size, int rows, const double* H,
                           int ldh, const double* res, const double* R, int r_kind, double* gamma_out) {
  IgvDeviceGuard dev_guard_(h);
  if (!gamma_out) return IGV_ERR_INVALID;
  return ekf_common(h, n_blocks, blk_idx, blk_size, rows, H, ldh, res, R, r_kind, nullptr, gamma_out, true);
}

igv_status igv_box_plus(igv_batch* h, const double* dx) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !dx) return IGV_ERR_INVALID;
  arena_reset(h);
  const double* d;
  IGV_TRY(stage(h, dx, (size_t)h->B * h->N, &d));
  igv_launch_boxplus(h, d);
  return check_launch(h);
}

// ---- fused visual update ------------------------------------------------------------------------------
igv_status igv_msckf_update(igv_batch* h, const igv_msckf_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  if (a->n_feats < 0 || a->n_feats > h->cfg.max_feats) return fail(h, IGV_ERR_CAPACITY, "n_feats exceeds max_feats");
  IgvLayout L = h->layout();
  if (a->obs_slots < L.n_clones) return fail(h, IGV_ERR_INVALID, "obs_slots smaller than the clone count");
  if (a->mode != IGV_VIS_ALL_OBS && a->mode != IGV_VIS_SELECTED) return IGV_ERR_INVALID;
  if (a->n_feats == 0 || L.n_clones == 0) return IGV_OK;   // `if (update_ids.size() == 0) return;`
  if (!a->pf_w || !a->anchor_slot || !a->obs || !a->obs_mask || !a->chi2_dof) return IGV_ERR_INVALID;
  if (h->chi2_n < 1) return fail(h, IGV_ERR_STATE, "chi^2 table not set (igv_set_chi2_table)");
  arena_reset(h);
  const size_t B = h->B, F = a->n_feats, SW = a->obs_slots;
  IgvMsckfLaunch m{};
  m.mode = a->mode; m.F = a->n_feats; m.obs_slots = a->obs_slots; m.max_valid = a->max_valid; m.noise = a->noise;
  IGV_TRY(stage(h, a->pf_w, B * F * 3, &m.pf));
  IGV_TRY(stage(h, a->anchor_slot, B * F, &m.anchor));
  IGV_TRY(stage(h, a->obs, B * F * SW * h->rho, &m.obs));
  IGV_TRY(stage(h, a->obs_mask, B * F * SW, &m.mask));
  IGV_TRY(stage(h, a->chi2_dof, B * F, &m.dof));
  IGV_TRY(stage(h, a->feat_ok, B * F, &m.feat_ok));
  igv_launch_msckf_features(h, m);
  IGV_TRY(check_launch(h));
  igv_launch_qr_compress(h, a->n_feats, a->max_valid);
  IGV_TRY(check_launch(h));
  const int n = 6 * L.n_clones;
  IgvEkfLaunch e{};
  e.blk.n_blocks = L.n_clones; e.blk.n = n;
  for (int s = 0; s < L.n_clones; ++s) { e.blk.idx[s] = L.idx_clone[s]; e.blk.size[s] = 6; }
  e.rows = n;
  e.H = h->Hc; e.strideH = (long)h->ncols_max * (h->ncols_max + 1); e.h_ld = n + 1; e.h_rowmajor = 1; e.h_upper = 1;
  e.res = h->Hc + n; e.strideRes = e.strideH; e.res_inc = n + 1;
  e.R = nullptr; e.strideR = 0; e.r_kind = IGV_R_ISO; e.r_iso_value = a->noise * a->noise;
  double* ddx = nullptr;
  IGV_TRY(out_buf(h, a->dx_out, B * h->N, &ddx));
  e.dx_out = ddx; e.gamma_only = 0; e.gamma_out = nullptr; e.gate_rows = nullptr; e.only_if = h->n_acc;
  e.apply_boxplus = 1;
  igv_launch_ekf(h, e);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, a->dx_out, ddx, B * h->N));
  if (a->n_accepted_out) {
    if (h->ptr_mode == IGV_PTR_DEVICE)
      IGV_CUDA(h, cudaMemcpyAsync(a->n_accepted_out, h->n_acc, sizeof(int) * B, cudaMemcpyDeviceToDevice, h->stream));
    else IGV_TRY(fetch(h, a->n_accepted_out, h->n_acc, B));
  }
  if (a->gamma_out) {
    const cudaMemcpyKind k = h->ptr_mode == IGV_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    IGV_CUDA(h, cudaMemcpy2DAsync(a->gamma_out, sizeof(double) * F, h->f_gamma, sizeof(double) * h->cfg.max_feats,
                                  sizeof(double) * F, B, k, h->stream));
    if (h->ptr_mode == IGV_PTR_HOST) IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return IGV_OK;
}

// ---- triangulation ----------------------------------------------------------------------------------------
igv_status igv_triangulate(igv_batch* h, const igv_tri_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a || !a->obs || !a->obs_mask || !a->pf_out || !a->ok_out) return IGV_ERR_INVALID;
  if (a->n_feats < 0 || a->n_feats > h->cfg.max_feats) return fail(h, IGV_ERR_CAPACITY, "n_feats exceeds max_feats");
  IgvLayout L = h->layout();
  if (a->obs_slots < L.n_clones) return fail(h, IGV_ERR_INVALID, "obs_slots smaller than the clone count");
  if (L.n_clones > 64) return fail(h, IGV_ERR_CAPACITY, "triangulation supports at most 64 clones");
  if (a->n_feats == 0) return IGV_OK;
  arena_reset(h);
  const size_t B = h->B, F = a->n_feats, SW = a->obs_slots;
  const double* dobs; const unsigned char* dmask; const int* danc;
  IGV_TRY(stage(h, a->obs, B * F * SW * h->rho, &dobs));
  IGV_TRY(stage(h, a->obs_mask, B * F * SW, &dmask));
  IGV_TRY(stage(h, a->anchor_slot, B * F, &danc));
  double* dpf; unsigned char* dok;
  IGV_TRY(out_buf(h, a->pf_out, B * F * 3, &dpf));
  IGV_TRY(out_buf(h, a->ok_out, B * F, &dok));
  igv_launch_triangulate(h, a->n_feats, a->obs_slots, dobs, dmask, danc, a->prm, dpf, dok);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, a->pf_out, dpf, B * F * 3));
  IGV_TRY(fetch(h, a->ok_out, dok, B * F));
  return IGV_OK;
}

// ---- GNSS residual generator (gnss_comm::psr_res / dopp_res) ------------------------------------------------
igv_status igv_gnss_residuals(igv_batch* h, const igv_gnss_res_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  if (a->n_sats < 0 || a->n_sats > h->cfg.max_sats) return fail(h, IGV_ERR_CAPACITY, "n_sats exceeds max_sats");
  if (a->n_sats == 0) return IGV_OK;
  if (!a->sat_pos || !a->sat_vel || !a->sat_clk || !a->obs || !a->obs_std || !a->ttx || !a->sys || !a->T_enu2ecef ||
      !a->unit || !a->res_pos || !a->res_vel || !a->sigma_psr || !a->sigma_dopp)
    return IGV_ERR_INVALID;
  arena_reset(h);
  const size_t B = h->B, S = a->n_sats;
  IgvGnssResLaunch g{};
  g.S = a->n_sats; g.psr_amp = a->psr_noise_amp; g.dopp_amp = a->dopp_noise_amp;
  IGV_TRY(stage(h, a->sat_pos, B * S * 3, &g.sat_pos));
  IGV_TRY(stage(h, a->sat_vel, B * S * 3, &g.sat_vel));
  IGV_TRY(stage(h, a->sat_clk, B * S * 3, &g.sat_clk));
  IGV_TRY(stage(h, a->obs, B * S * 3, &g.obs));
  IGV_TRY(stage(h, a->obs_std, B * S * 3, &g.obs_std));
  IGV_TRY(stage(h, a->ttx, B * S * 2, &g.ttx));
  IGV_TRY(stage(h, a->sys, B * S, &g.sys));
  IGV_TRY(stage(h, a->T_enu2ecef, B * 12, &g.T));
  IGV_TRY(stage(h, a->iono, B * 8, &g.iono));
  IGV_TRY(stage(h, a->clock_init, B * 5, &g.clock_init));
  IGV_TRY(out_buf(h, a->unit, B * S * 3, &g.unit));
  IGV_TRY(out_buf(h, a->res_pos, B * S, &g.res_pos));
  IGV_TRY(out_buf(h, a->res_vel, B * S, &g.res_vel));
  IGV_TRY(out_buf(h, a->sigma_psr, B * S, &g.sig_psr));
  IGV_TRY(out_buf(h, a->sigma_dopp, B * S, &g.sig_dopp));
  IGV_TRY(out_buf(h, a->azel, B * S * 2, &g.azel));
  IGV_TRY(out_buf(h, a->atmos, B * S * 2, &g.atmos));
  igv_launch_gnss_residuals(h, g);
  IGV_TRY(check_launch(h));
  if (h->ptr_mode == IGV_PTR_HOST) {
    auto back = [&](double* user, const double* dev, size_t n) {
      if (user) cudaMemcpyAsync(user, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    };
    back(a->unit, g.unit, B * S * 3); back(a->res_pos, g.res_pos, B * S); back(a->res_vel, g.res_vel, B * S);
    back(a->sigma_psr, g.sig_psr, B * S); back(a->sigma_dopp, g.sig_dopp, B * S);
    back(a->azel, g.azel, B * S * 2); back(a->atmos, g.atmos, B * S * 2);
    IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return IGV_OK;
}

// ---- ephemeris -> satellite states (gnss_comm::sat_states) ------------------------------------------------------
igv_status igv_sat_states(igv_batch* h, const igv_sat_state_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  if (a->n_sats < 0 || a->n_sats > h->cfg.max_sats) return fail(h, IGV_ERR_CAPACITY, "n_sats exceeds max_sats");
  if (a->n_sats == 0) return IGV_OK;
  if (!a->eph || !a->t_obs_rel || !a->psr || !a->sys || !a->sat_pos || !a->sat_vel || !a->sat_clk) return IGV_ERR_INVALID;
  arena_reset(h);
  const size_t B = h->B, S = a->n_sats;
  const double *eph, *tob, *psr;
  const int* sys;
  IGV_TRY(stage(h, a->eph, B * S * IGV_EPH_STRIDE, &eph));
  IGV_TRY(stage(h, a->t_obs_rel, B * S, &tob));
  IGV_TRY(stage(h, a->psr, B * S, &psr));
  IGV_TRY(stage(h, a->sys, B * S, &sys));
  double *pos, *vel, *clk, *ttx;
  IGV_TRY(out_buf(h, a->sat_pos, B * S * 3, &pos));
  IGV_TRY(out_buf(h, a->sat_vel, B * S * 3, &vel));
  IGV_TRY(out_buf(h, a->sat_clk, B * S * 3, &clk));
  IGV_TRY(out_buf(h, a->ttx_rel, B * S, &ttx));
  igv_launch_sat_states(h, a->n_sats, eph, tob, psr, sys, pos, vel, clk, ttx);
  IGV_TRY(check_launch(h));
  if (h->ptr_mode == IGV_PTR_HOST) {
    auto back = [&](double* user, const double* dev, size_t n) {
      if (user) cudaMemcpyAsync(user, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
    };
    back(a->sat_pos, pos, B * S * 3); back(a->sat_vel, vel, B * S * 3); back(a->sat_clk, clk, B * S * 3);
    back(a->ttx_rel, ttx, B * S);
    IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return IGV_OK;
}

// ---- fused GNSS update --------------------------------------------------------------------------------
igv_status igv_gnss_update(igv_batch* h, const igv_gnss_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  if (a->n_sats < 0 || a->n_sats > h->cfg.max_sats) return fail(h, IGV_ERR_CAPACITY, "n_sats exceeds max_sats");
  if (a->n_sats == 0) return IGV_OK;                       // GnssUpdate.cpp:92-93
  IgvLayout L = h->layout();
  // GnssManager::checkGnssStates (GnssManager.cpp:86-98)
  if (L.idx_gnss[IGV_GNSS_YOF] < 0 || L.idx_gnss[IGV_GNSS_FS] < 0) return IGV_OK;
  bool any = false;
  for (int i = 0; i < 4; ++i) any = any || L.idx_gnss[i] >= 0;
  if (!any) return IGV_OK;
  if (!a->unit || !a->res_pos || !a->res_vel || !a->sigma_psr || !a->sigma_dopp || !a->sys || !a->R_enu2ecef)
    return IGV_ERR_INVALID;
  if ((a->chi2_test || a->strong_reject) && h->chi2_n < 14)
    return fail(h, IGV_ERR_STATE, "chi^2 table not set (igv_set_chi2_table)");
  arena_reset(h);
  const size_t B = h->B, S = a->n_sats;
  IgvGnssLaunch g{};
  g.S = a->n_sats; g.adjust_yof = a->is_adjust_yof; g.chi2_test = a->chi2_test;
  IGV_TRY(stage(h, a->unit, B * S * 3, &g.unit));
  IGV_TRY(stage(h, a->res_pos, B * S, &g.res_pos));
  IGV_TRY(stage(h, a->res_vel, B * S, &g.res_vel));
  IGV_TRY(stage(h, a->sigma_psr, B * S, &g.sig_psr));
  IGV_TRY(stage(h, a->sigma_dopp, B * S, &g.sig_dopp));
  IGV_TRY(stage(h, a->sys, B * S, &g.sys));
  IGV_TRY(stage(h, a->R_enu2ecef, B * 9, &g.R_enu2ecef));
  // var_order: SE23, YOF, clock biases present, FS
  IgvBlocks& blk = g.blk;
  blk.n_blocks = 0; blk.n = 0;
  auto push = [&](int idx, int size) { blk.idx[blk.n_blocks] = idx; blk.size[blk.n_blocks] = size; blk.n_blocks++; blk.n += size; };
  for (int i = 0; i < 6; ++i) g.col_of_gnss[i] = -1;
  push(0, 9);
  g.col_of_gnss[IGV_GNSS_YOF] = blk.n; push(L.idx_gnss[IGV_GNSS_YOF], 1);
  for (int i = 0; i < 4; ++i) if (L.idx_gnss[i] >= 0) { g.col_of_gnss[i] = blk.n; push(L.idx_gnss[i], 1); }
  g.col_of_gnss[IGV_GNSS_FS] = blk.n; push(L.idx_gnss[IGV_GNSS_FS], 1);
  igv_launch_gnss_rows(h, g);
  IGV_TRY(check_launch(h));
  IgvEkfLaunch e{};
  e.blk = blk; e.rows = 2 * a->n_sats;
  const int ldh = 2 * h->cfg.max_sats;
  e.H = h->Hg; e.strideH = (long)ldh * 16; e.h_ld = ldh; e.h_rowmajor = 0;
  e.res = h->rg; e.strideRes = ldh; e.res_inc = 1;
  e.R = h->Rg; e.strideR = ldh; e.r_kind = IGV_R_DIAG;
  double* ddx = nullptr;
  IGV_TRY(out_buf(h, a->dx_out, B * h->N, &ddx));
  e.dx_out = ddx; e.gamma_only = 0; e.gamma_out = nullptr;
  e.gate_rows = a->strong_reject ? h->cnt_g : nullptr;
  e.only_if = h->cnt_g;   // `if rows == 0` nothing to do
  e.apply_boxplus = 1;
  igv_launch_ekf(h, e);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, a->dx_out, ddx, B * h->N));
  return IGV_OK;
}

// ---- delayed init / linear replace ------------------------------------------------------------------------
// Shared tail of igv_add_variable_delayed / igv_gnss_add_new_tracked_sys. All pointers are DEVICE pointers here.
// noise2_dev / rows_dev (optional, B each): per-sequence measurement variance and true row count (rows beyond it are
// zero padding: they change neither the Givens split nor the posterior, only the dof of the gate).
static igv_status delayed_common(igv_batch* h, int gtype, const double* dval, const IgvBlocks& blk, int rows, int k,
                                 int lm_anchor, const double* dHo, const double* dHn, const double* dr, double noise_iso,
                                 const double* noise2_dev, const int* rows_dev, double chi2_mult, int do_chi2,
                                 double prior_cov_if_rejected, int* accepted_out, double* dx_out) {
  const size_t B = h->B;
  int* dacc = h->n_acc;
  if ((size_t)rows * (blk.n + 2) + 16 > h->dws_stride) return fail(h, IGV_ERR_CAPACITY, "delayed-init workspace too small");
  igv_launch_delayed_init(h, blk, rows, k, dHo, dHn, dr, noise_iso, noise2_dev, rows_dev, chi2_mult, do_chi2,
                          prior_cov_if_rejected, dacc);
  IGV_TRY(check_launch(h));
  if (lm_anchor >= 0) {   // a landmark: value = the triangulated world position, tied to its anchor clone
    int lm_slot = 0;
    for (const auto& v : h->vars) if (v.kind == VK_LANDMARK) ++lm_slot;
    h->vars.push_back({VK_LANDMARK, h->N, 3, lm_anchor});
    reindex(h);
    igv_launch_set_lm_value(h, lm_slot, dval);
    IGV_TRY(check_launch(h));
  } else {
    h->vars.push_back({gtype >= 0 ? VK_GNSS : VK_OPAQUE, h->N, 1, gtype >= 0 ? gtype : 0});
    reindex(h);
    if (gtype >= 0) { igv_launch_set_gnss_value(h, gtype, dval); IGV_TRY(check_launch(h)); }
  }
  // EKF on the remaining rows (StateManager.cpp:626-627), only where the variable was accepted
  IgvEkfLaunch e{};
  e.blk = blk; e.rows = rows - k;
  e.H = h->Dws + k; e.strideH = (long)rows * blk.n; e.h_ld = rows; e.h_rowmajor = 0;
  e.res = h->Dws + B * rows * blk.n + k; e.strideRes = rows; e.res_inc = 1;
  e.R = noise2_dev; e.strideR = noise2_dev ? 1 : 0; e.r_kind = IGV_R_ISO; e.r_iso_value = noise_iso * noise_iso;
  e.only_if = dacc; e.apply_boxplus = 1;
  double* ddx = nullptr;
  IGV_TRY(out_buf(h, dx_out, B * h->N, &ddx));
  e.dx_out = ddx;
  igv_launch_ekf(h, e);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, dx_out, ddx, B * h->N));
  if (accepted_out) {
    if (h->ptr_mode == IGV_PTR_DEVICE)
      IGV_CUDA(h, cudaMemcpyAsync(accepted_out, dacc, sizeof(int) * B, cudaMemcpyDeviceToDevice, h->stream));
    else IGV_TRY(fetch(h, accepted_out, dacc, B));
  }
  return IGV_OK;
}

igv_status igv_add_variable_delayed(igv_batch* h, int gtype, const double* value, int n_blocks, const int* blk_idx,
                                    const int* blk_size, int rows, const double* H_old, const double* H_new,
                                    const double* res, double noise_iso, double chi2_mult, int do_chi2,
                                    double prior_cov_if_rejected, int* accepted_out, double* dx_out) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !H_old || !H_new || !res || rows < 1 || rows > 128 || gtype > 5) return IGV_ERR_INVALID;
  if (rows <= 1) return fail(h, IGV_ERR_INVALID, "H_new rows should be larger than H_new cols");  // StateManager.cpp:574-578
  if (gtype >= 0 && h->layout().idx_gnss[gtype] >= 0) return fail(h, IGV_ERR_STATE, "New var already in state");
  if (h->N + 1 > h->cfg.max_dim) return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds max_dim");
  if (rows - 1 > h->max_rows) return fail(h, IGV_ERR_CAPACITY, "rows exceed the EKF workspace");
  arena_reset(h);
  IgvBlocks blk;
  IGV_TRY(make_blocks(h, n_blocks, blk_idx, blk_size, &blk));
  const size_t B = h->B;
  const double *dHo, *dHn, *dr, *dval = nullptr;
  IGV_TRY(stage(h, H_old, B * rows * blk.n, &dHo));
  IGV_TRY(stage(h, H_new, B * rows, &dHn));
  IGV_TRY(stage(h, res, B * rows, &dr));
  if (gtype >= 0) IGV_TRY(stage(h, value, B, &dval));
  if (blk.n > h->dws_cols) return fail(h, IGV_ERR_CAPACITY, "too many measured columns for the delayed-init workspace");
  return delayed_common(h, gtype, dval, blk, rows, 1, -1, dHo, dHn, dr, noise_iso, nullptr, nullptr, chi2_mult, do_chi2,
                        prior_cov_if_rejected, accepted_out, dx_out);
}

// GnssUpdate::addNewTrackedSys for ONE system (GnssUpdate.cpp:317-476; the reference loops over sys_to_add and calls
// StateManager::addVariableDelayed per system): rows of the satellites of that system (all satellites for the clock
// drift FS) are built on the device -- H_x = [u^T R_w2e [x]x | -u^T R_w2e on p (clock bias) or v (FS) | yof], H_f = 1,
// res = -res_pos / -res_vel, noise = sqrt(mean sigma_i^2) -- and handed to the delayed initialisation.
igv_status igv_gnss_add_new_tracked_sys(igv_batch* h, const igv_gnss_new_sys_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  if (a->gtype < 0 || a->gtype > IGV_GNSS_FS) return IGV_ERR_INVALID;
  if (a->n_sats < 2 || a->n_sats > h->cfg.max_sats) return fail(h, IGV_ERR_CAPACITY, "n_sats outside [2, max_sats]");
  if (!a->unit || !a->res_pos || !a->res_vel || !a->sigma_psr || !a->sigma_dopp || !a->sys || !a->R_enu2ecef || !a->value)
    return IGV_ERR_INVALID;
  IgvLayout L = h->layout();
  if (L.idx_gnss[IGV_GNSS_YOF] < 0) return IGV_OK;                        // GnssUpdate.cpp:329-330: nothing to do
  if (L.idx_gnss[a->gtype] >= 0) return fail(h, IGV_ERR_STATE, "New var already in state");
  if (h->N + 1 > h->cfg.max_dim) return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds max_dim");
  if (a->n_sats - 1 > h->max_rows || a->n_sats > 128) return fail(h, IGV_ERR_CAPACITY, "rows exceed the EKF workspace");
  arena_reset(h);
  const size_t B = h->B, S = a->n_sats;
  IgvGnssNewRowsLaunch g{};
  g.S = a->n_sats; g.gtype = a->gtype; g.adjust_yof = a->is_adjust_yof;
  IGV_TRY(stage(h, a->unit, B * S * 3, &g.unit));
  IGV_TRY(stage(h, a->res_pos, B * S, &g.res_pos));
  IGV_TRY(stage(h, a->res_vel, B * S, &g.res_vel));
  IGV_TRY(stage(h, a->sigma_psr, B * S, &g.sig_psr));
  IGV_TRY(stage(h, a->sigma_dopp, B * S, &g.sig_dopp));
  IGV_TRY(stage(h, a->sys, B * S, &g.sys));
  IGV_TRY(stage(h, a->R_enu2ecef, B * 9, &g.R_enu2ecef));
  IGV_TRY(stage(h, a->R_ecef2enu, B * 9, &g.R_ecef2enu));
  const double* dval;
  IGV_TRY(stage(h, a->value, B, &dval));
  // workspace in the staging slot: Hx (S x 10 col-major) | Hf (S) | res (S) | noise^2 | count
  char* mem = nullptr;
  IGV_TRY(arena_reserve(h, sizeof(double) * B * (S * 12 + 1) + sizeof(int) * B, &mem));
  g.Hx = reinterpret_cast<double*>(mem);
  g.Hf = g.Hx + B * S * 10;
  g.res = g.Hf + B * S;
  g.noise2 = g.res + B * S;
  g.count = reinterpret_cast<int*>(g.noise2 + B);
  igv_launch_gnss_new_rows(h, g);
  IGV_TRY(check_launch(h));
  IgvBlocks blk;
  blk.n_blocks = 2; blk.n = 10;
  blk.idx[0] = 0; blk.size[0] = 9;
  blk.idx[1] = L.idx_gnss[IGV_GNSS_YOF]; blk.size[1] = 1;
  return delayed_common(h, a->gtype, dval, blk, a->n_sats, 1, -1, g.Hx, g.Hf, g.res, 0.0, g.noise2, g.count,
                        a->chi2_mult > 0.0 ? a->chi2_mult : 0.95, 1, a->prior_cov_if_rejected, a->accepted_out, a->dx_out);
}

// ---- SLAM landmarks ----------------------------------------------------------------------------------------------
int igv_num_landmarks(const igv_batch* h) { return h ? h->layout().n_lm : -1; }
int igv_landmark_idx(const igv_batch* h, int lm_slot) {
  if (!h) return -1;
  IgvLayout L = h->layout();
  return (lm_slot >= 0 && lm_slot < L.n_lm) ? L.idx_lm[lm_slot] : -1;
}
int igv_landmark_anchor(const igv_batch* h, int lm_slot) {
  if (!h) return -1;
  IgvLayout L = h->layout();
  return (lm_slot >= 0 && lm_slot < L.n_lm) ? L.lm_anchor[lm_slot] : -1;
}

igv_status igv_landmark_init(igv_batch* h, const igv_lm_init_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a || !a->pf_w || !a->obs || !a->obs_mask) return IGV_ERR_INVALID;
  if (h->rho != 2) return fail(h, IGV_ERR_STATE, "landmarks: mono handles only");
  IgvLayout L = h->layout();
  if (L.n_lm >= h->cfg.max_landmarks) return fail(h, IGV_ERR_CAPACITY, "no vacant landmark slot (max_landmarks)");
  if (a->obs_slots < L.n_clones || L.n_clones < 2) return fail(h, IGV_ERR_INVALID, "obs_slots smaller than the clone count");
  if (a->anchor_slot < 0 || a->anchor_slot >= L.n_clones) return fail(h, IGV_ERR_INVALID, "anchor clone not in the window");
  if (h->N + 3 > h->cfg.max_dim) return fail(h, IGV_ERR_CAPACITY, "state dimension exceeds max_dim");
  const int rows = 2 * L.n_clones, n = 6 * L.n_clones;
  if (rows > 128 || rows - 3 > h->max_rows || n > h->dws_cols) return fail(h, IGV_ERR_CAPACITY, "window too large for the delayed-init workspace");
  arena_reset(h);
  const size_t B = h->B, SW = a->obs_slots;
  IgvLmInitLaunch g{};
  g.SW = a->obs_slots; g.anchor_slot = a->anchor_slot; g.rows_max = rows;
  IGV_TRY(stage(h, a->pf_w, B * 3, &g.pf));
  IGV_TRY(stage(h, a->obs, B * SW * 2, &g.obs));
  IGV_TRY(stage(h, a->obs_mask, B * SW, &g.mask));
  char* mem = nullptr;
  IGV_TRY(arena_reserve(h, sizeof(double) * B * ((size_t)rows * (n + 4)) + sizeof(int) * B, &mem));
  g.Hx = reinterpret_cast<double*>(mem);
  g.Hf = g.Hx + B * rows * n;
  g.res = g.Hf + B * rows * 3;
  g.count = reinterpret_cast<int*>(g.res + B * rows);
  igv_launch_lm_init_rows(h, g);
  IGV_TRY(check_launch(h));
  IgvBlocks blk;
  blk.n_blocks = L.n_clones; blk.n = n;
  for (int s = 0; s < L.n_clones; ++s) { blk.idx[s] = L.idx_clone[s]; blk.size[s] = 6; }
  return delayed_common(h, -1, g.pf, blk, rows, 3, a->anchor_slot, g.Hx, g.Hf, g.res, a->noise, nullptr, g.count,
                        a->chi2_mult > 0.0 ? a->chi2_mult : 0.95, 1, a->prior_cov_if_rejected, a->accepted_out, nullptr);
}

igv_status igv_landmark_update(igv_batch* h, const igv_lm_update_args* a) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !a) return IGV_ERR_INVALID;
  IgvLayout L = h->layout();
  if (L.n_lm == 0) return IGV_OK;                                   // LandmarkUpdate.cpp:35
  if (!a->uv || !a->valid) return IGV_ERR_INVALID;
  if (h->chi2_n < 2) return fail(h, IGV_ERR_STATE, "chi^2 table not set (igv_set_chi2_table)");
  const int rows = 2 * L.n_lm, ncols = 15 + 6 * L.n_clones + 3 * L.n_lm;
  if (rows > h->max_rows) return fail(h, IGV_ERR_CAPACITY, "rows exceed the EKF workspace (max_rows)");
  arena_reset(h);
  const size_t B = h->B, nl = L.n_lm;
  IgvLmUpdateLaunch g{};
  IGV_TRY(stage(h, a->uv, B * nl * 2, &g.uv));
  IGV_TRY(stage(h, a->valid, B * nl, &g.valid));
  g.noise2 = a->noise * a->noise;
  g.H = h->Lws; g.ldh = rows; g.ncols = ncols;
  g.res = h->Lws + B * (size_t)rows * ncols;
  double* dgam = nullptr;
  IGV_TRY(out_buf(h, a->gamma_out, B * nl, &dgam));
  g.gamma = dgam; g.n_acc = h->n_acc;
  igv_launch_lm_update_rows(h, g);
  IGV_TRY(check_launch(h));
  IgvEkfLaunch e{};
  IgvBlocks& blk = e.blk;
  blk.n_blocks = 0; blk.n = 0;
  auto push = [&](int idx, int size) { blk.idx[blk.n_blocks] = idx; blk.size[blk.n_blocks] = size; blk.n_blocks++; blk.n += size; };
  push(0, 9); push(15, 6);
  for (int s = 0; s < L.n_clones; ++s) push(L.idx_clone[s], 6);
  for (int l = 0; l < L.n_lm; ++l) push(L.idx_lm[l], 3);
  e.rows = rows; e.H = g.H; e.strideH = (long)rows * ncols; e.h_ld = rows; e.h_rowmajor = 0;
  e.res = g.res; e.strideRes = rows; e.res_inc = 1;
  e.R = nullptr; e.strideR = 0; e.r_kind = IGV_R_ISO; e.r_iso_value = g.noise2;
  e.only_if = h->n_acc; e.apply_boxplus = 1;
  igv_launch_ekf(h, e);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, a->gamma_out, dgam, B * nl));
  if (a->n_accepted_out) {
    if (h->ptr_mode == IGV_PTR_DEVICE)
      IGV_CUDA(h, cudaMemcpyAsync(a->n_accepted_out, h->n_acc, sizeof(int) * B, cudaMemcpyDeviceToDevice, h->stream));
    else IGV_TRY(fetch(h, a->n_accepted_out, h->n_acc, B));
  }
  return IGV_OK;
}

igv_status igv_landmark_change_anchor(igv_batch* h, int lm_slot, int new_clone_slot) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  IgvLayout L = h->layout();
  if (lm_slot < 0 || lm_slot >= L.n_lm) return fail(h, IGV_ERR_STATE, "landmark slot not in the state");
  if (new_clone_slot < 0 || new_clone_slot >= L.n_clones) return fail(h, IGV_ERR_STATE, "clone slot not in the sliding window");
  if (L.n_clones < 2) return IGV_OK;                                 // MapServerManager.cpp:347
  const int old = L.lm_anchor[lm_slot];
  if (old == new_clone_slot) return IGV_OK;
  arena_reset(h);
  char* mem = nullptr;
  IGV_TRY(arena_reserve(h, sizeof(double) * (size_t)h->B * 45, &mem));
  double* H = reinterpret_cast<double*>(mem);
  igv_launch_lm_anchor_H(h, lm_slot, H);
  IGV_TRY(check_launch(h));
  IgvBlocks blk;
  blk.n_blocks = 3; blk.n = 15;
  blk.idx[0] = L.idx_clone[old]; blk.size[0] = 6;
  blk.idx[1] = L.idx_clone[new_clone_slot]; blk.size[1] = 6;
  blk.idx[2] = L.idx_lm[lm_slot]; blk.size[2] = 3;
  igv_launch_replace_var_linear(h, L.idx_lm[lm_slot], 3, blk, H);
  IGV_TRY(check_launch(h));
  int l = 0;
  for (auto& v : h->vars)
    if (v.kind == VK_LANDMARK) { if (l == lm_slot) v.tag = new_clone_slot; ++l; }
  return IGV_OK;
}

igv_status igv_landmark_marginalize(igv_batch* h, int lm_slot) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  int l = 0;
  for (size_t i = 0; i < h->vars.size(); ++i)
    if (h->vars[i].kind == VK_LANDMARK) {
      if (l == lm_slot) return marginalize_var(h, i);
      ++l;
    }
  return fail(h, IGV_ERR_STATE, "[StateManager]: Landmark id not exists in state! Cannot marg!");   // StateManager.cpp:342-346
}

igv_status igv_replace_var_linear(igv_batch* h, int target_idx, int target_size, int n_blocks, const int* blk_idx,
                                  const int* blk_size, const double* H) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !H || target_size < 1 || target_size > 6) return IGV_ERR_INVALID;
  bool found = false;
  for (const auto& v : h->vars) if (v.idx == target_idx && v.size == target_size) found = true;
  if (!found) return fail(h, IGV_ERR_STATE, "Target var not in state, cannot linearly replace");
  arena_reset(h);
  IgvBlocks blk;
  IGV_TRY(make_blocks(h, n_blocks, blk_idx, blk_size, &blk));
  const double* dH;
  IGV_TRY(stage(h, H, (size_t)h->B * target_size * blk.n, &dH));
  igv_launch_replace_var_linear(h, target_idx, target_size, blk, dH);
  return check_launch(h);
}

// ---- read-outs --------------------------------------------------------------------------------------------
// ---- track table (MapServer on the device) ------------------------------------------------------------------------
igv_status igv_tracks_create(igv_batch* h, int max_tracks) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || max_tracks < 1 || max_tracks > 4096) return IGV_ERR_INVALID;
  if (h->trk.T != 0) return fail(h, IGV_ERR_STATE, "track table already created");
  if (h->cfg.max_clones < 1) return fail(h, IGV_ERR_CAPACITY, "track table needs max_clones >= 1");
  IgvTrackTable& t = h->trk;
  const size_t n = (size_t)h->B * max_tracks, C = (size_t)h->cfg.max_clones;
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.id), sizeof(int) * n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.mask), sizeof(unsigned long long) * n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.st), n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.anchor), sizeof(int) * n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.pf), sizeof(double) * 3 * n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.pf_fej), sizeof(double) * 3 * n));
  IGV_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&t.obs), sizeof(double) * n * C * h->rho));
  IGV_CUDA(h, cudaMemsetAsync(t.obs, 0, sizeof(double) * n * C * h->rho, h->stream));
  t.T = max_tracks;
  t.C = (int)C;
  t.col_of_slot.clear();
  const int nc = h->layout().n_clones;
  for (int s = 0; s < nc; ++s) t.col_of_slot.push_back(s);
  igv_launch_trk_reset(h);
  return check_launch(h);
}

igv_status igv_tracks_reset(igv_batch* h) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  igv_launch_trk_reset(h);
  return check_launch(h);
}

int igv_tracks_capacity(const igv_batch* h) { return h ? h->trk.T : 0; }

igv_status igv_tracks_collect(igv_batch* h, const int* n_meas, int meas_stride, const unsigned long long* ids,
                              const double* uv) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (!n_meas || meas_stride < 0 || meas_stride > 4096) return IGV_ERR_INVALID;
  if (meas_stride == 0) return IGV_OK;
  if (!ids || !uv) return IGV_ERR_INVALID;
  if (h->trk.col_of_slot.empty())
    return fail(h, IGV_ERR_STATE, "[FeatureInfoManager]: Meas timestamp not in sw!");   // MapServerManager.cpp:107-111
  arena_reset(h);
  const size_t B = h->B, M = meas_stride;
  const int* dn; const unsigned long long* did; const double* duv;
  IGV_TRY(stage(h, n_meas, B, &dn));
  IGV_TRY(stage(h, ids, B * M, &did));
  IGV_TRY(stage(h, uv, B * M * h->rho, &duv));
  igv_launch_trk_collect(h, dn, meas_stride, did, duv);
  return check_launch(h);
}

igv_status igv_tracks_mark_lost(igv_batch* h) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  igv_launch_trk_mark_lost(h);
  return check_launch(h);
}

igv_status igv_tracks_gather(igv_batch* h, const igv_track_gather_args* a) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (!a || !a->track_entry || !a->n_sel || !a->obs || !a->mask_all || !a->mask_upd || !a->anchor_slot ||
      !a->chi2_dof || !a->feat_ok)
    return IGV_ERR_INVALID;
  if (a->rule != IGV_TRK_LOST && a->rule != IGV_TRK_SEEN_AT) return IGV_ERR_INVALID;
  if (a->n_feats < 1 || a->n_feats > h->cfg.max_feats) return fail(h, IGV_ERR_CAPACITY, "n_feats exceeds max_feats");
  const int nc = (int)h->trk.col_of_slot.size();
  if (a->obs_slots < nc || a->obs_slots < 1) return fail(h, IGV_ERR_INVALID, "obs_slots smaller than the clone count");
  IgvTrkGatherLaunch g{};
  g.rule = a->rule; g.n_selected = a->n_selected; g.min_obs = a->min_obs; g.dof_fixed = a->dof_fixed;
  g.F = a->n_feats; g.SW = a->obs_slots;
  if (a->rule == IGV_TRK_SEEN_AT) {
    if (a->n_selected < 1) return fail(h, IGV_ERR_INVALID, "SEEN_AT needs at least one selected clone");
    IGV_TRY(trk_slot_bits(h, a->n_selected, a->selected_slots, &g.sel_cols));
  }
  arena_reset(h);
  const size_t B = h->B, F = a->n_feats, SW = a->obs_slots;
  IGV_TRY(out_buf(h, a->track_entry, B * F, &g.entry));
  IGV_TRY(out_buf(h, a->n_sel, B, &g.n_sel));
  IGV_TRY(out_buf(h, a->track_id, B * F, &g.track_id));
  IGV_TRY(out_buf(h, a->obs, B * F * SW * h->rho, &g.obs));
  IGV_TRY(out_buf(h, a->mask_all, B * F * SW, &g.mask_all));
  IGV_TRY(out_buf(h, a->mask_upd, B * F * SW, &g.mask_upd));
  IGV_TRY(out_buf(h, a->anchor_slot, B * F, &g.anchor_slot));
  IGV_TRY(out_buf(h, a->chi2_dof, B * F, &g.dof));
  IGV_TRY(out_buf(h, a->feat_ok, B * F, &g.feat_ok));
  igv_launch_trk_gather(h, g);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, a->track_entry, g.entry, B * F));
  IGV_TRY(fetch(h, a->n_sel, g.n_sel, B));
  IGV_TRY(fetch(h, a->track_id, g.track_id, B * F));
  IGV_TRY(fetch(h, a->obs, g.obs, B * F * SW * h->rho));
  IGV_TRY(fetch(h, a->mask_all, g.mask_all, B * F * SW));
  IGV_TRY(fetch(h, a->mask_upd, g.mask_upd, B * F * SW));
  IGV_TRY(fetch(h, a->anchor_slot, g.anchor_slot, B * F));
  IGV_TRY(fetch(h, a->chi2_dof, g.dof, B * F));
  IGV_TRY(fetch(h, a->feat_ok, g.feat_ok, B * F));
  return IGV_OK;
}

igv_status igv_tracks_commit_tri(igv_batch* h, int n_feats, const int* track_entry, const double* pf,
                                 const unsigned char* ok, unsigned char* feat_ok) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (n_feats < 0 || !track_entry || !pf || !ok) return IGV_ERR_INVALID;
  if (n_feats == 0) return IGV_OK;
  arena_reset(h);
  const size_t n = (size_t)h->B * n_feats;
  const int* de; const double* dpf; const unsigned char* dok; unsigned char* dfo = nullptr;
  IGV_TRY(stage(h, track_entry, n, &de));
  IGV_TRY(stage(h, pf, 3 * n, &dpf));
  IGV_TRY(stage(h, ok, n, &dok));
  if (feat_ok) {
    if (h->ptr_mode == IGV_PTR_DEVICE) {
      dfo = feat_ok;
    } else {   // in/out argument: staged copy in, fetched back
      const unsigned char* in;
      IGV_TRY(stage(h, static_cast<const unsigned char*>(feat_ok), n, &in));
      dfo = const_cast<unsigned char*>(in);
    }
  }
  igv_launch_trk_commit_tri(h, n_feats, de, dpf, dok, dfo);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, feat_ok, static_cast<const unsigned char*>(dfo), n));
  return IGV_OK;
}

igv_status igv_tracks_erase(igv_batch* h, int n_feats, const int* track_entry) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (n_feats < 0 || !track_entry) return IGV_ERR_INVALID;
  if (n_feats == 0) return IGV_OK;
  arena_reset(h);
  const int* de;
  IGV_TRY(stage(h, track_entry, (size_t)h->B * n_feats, &de));
  igv_launch_trk_erase(h, n_feats, de);
  return check_launch(h);
}

igv_status igv_tracks_clean_obs(igv_batch* h, int n_slots, const int* clone_slots) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  unsigned long long bits;
  IGV_TRY(trk_slot_bits(h, n_slots, clone_slots, &bits));
  if (bits == 0ull) return IGV_OK;
  igv_launch_trk_clean(h, bits, 1);
  return check_launch(h);
}

igv_status igv_tracks_change_anchor(igv_batch* h, int n_old, const int* old_slots, double min_depth) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  unsigned long long bits;
  IGV_TRY(trk_slot_bits(h, n_old, old_slots, &bits));
  if (bits == 0ull) return IGV_OK;
  igv_launch_trk_change_anchor(h, bits, min_depth);
  return check_launch(h);
}

igv_status igv_tracks_erase_invalid(igv_batch* h, double min_depth) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (h->trk.col_of_slot.empty()) return IGV_OK;
  igv_launch_trk_erase_invalid(h, min_depth);
  return check_launch(h);
}

igv_status igv_tracks_get(igv_batch* h, const igv_track_dump* d) {
  IgvDeviceGuard dev_guard_(h);
  IGV_TRY(trk_ready(h));
  if (!d) return IGV_ERR_INVALID;
  if (d->obs && d->obs_slots < (int)h->trk.col_of_slot.size())
    return fail(h, IGV_ERR_INVALID, "obs_slots smaller than the clone count");
  arena_reset(h);
  const size_t n = (size_t)h->B * h->trk.T, SW = d->obs ? d->obs_slots : 0;
  igv_track_dump dev = *d;
  IGV_TRY(out_buf(h, d->id, n, &dev.id));
  IGV_TRY(out_buf(h, d->used, n, &dev.used));
  IGV_TRY(out_buf(h, d->to_marg, n, &dev.to_marg));
  IGV_TRY(out_buf(h, d->is_tri, n, &dev.is_tri));
  IGV_TRY(out_buf(h, d->slot_mask, n, &dev.slot_mask));
  IGV_TRY(out_buf(h, d->anchor_slot, n, &dev.anchor_slot));
  IGV_TRY(out_buf(h, d->pf, 3 * n, &dev.pf));
  IGV_TRY(out_buf(h, d->pf_fej, 3 * n, &dev.pf_fej));
  IGV_TRY(out_buf(h, d->obs, n * SW * h->rho, &dev.obs));
  IGV_TRY(out_buf(h, d->n_tracks, (size_t)h->B, &dev.n_tracks));
  igv_launch_trk_dump(h, dev);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, d->id, dev.id, n));
  IGV_TRY(fetch(h, d->used, dev.used, n));
  IGV_TRY(fetch(h, d->to_marg, dev.to_marg, n));
  IGV_TRY(fetch(h, d->is_tri, dev.is_tri, n));
  IGV_TRY(fetch(h, d->slot_mask, dev.slot_mask, n));
  IGV_TRY(fetch(h, d->anchor_slot, dev.anchor_slot, n));
  IGV_TRY(fetch(h, d->pf, dev.pf, 3 * n));
  IGV_TRY(fetch(h, d->pf_fej, dev.pf_fej, 3 * n));
  IGV_TRY(fetch(h, d->obs, dev.obs, n * SW * h->rho));
  IGV_TRY(fetch(h, d->n_tracks, dev.n_tracks, (size_t)h->B));
  if (h->ptr_mode == IGV_PTR_DEVICE) IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

igv_status igv_get_flags(igv_batch* h, int* flags_out, int clear) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !flags_out) return IGV_ERR_INVALID;
  const cudaMemcpyKind k = h->ptr_mode == IGV_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  IGV_CUDA(h, cudaMemcpyAsync(flags_out, h->flags, sizeof(int) * h->B, k, h->stream));
  if (clear) IGV_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(int) * h->B, h->stream));
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

igv_status igv_cov_trace(igv_batch* h, double* trace_out) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !trace_out) return IGV_ERR_INVALID;
  arena_reset(h);
  double* dev;
  IGV_TRY(out_buf(h, trace_out, (size_t)h->B, &dev));
  igv_launch_trace(h, dev);
  IGV_TRY(check_launch(h));
  IGV_TRY(fetch(h, trace_out, dev, (size_t)h->B));
  if (h->ptr_mode == IGV_PTR_DEVICE) IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  return IGV_OK;
}

igv_status igv_cov_trace_async(igv_batch* h, double* trace_out) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !trace_out) return IGV_ERR_INVALID;
  arena_reset(h);
  double* dev;
  IGV_TRY(out_buf(h, trace_out, (size_t)h->B, &dev));
  igv_launch_trace(h, dev);
  IGV_TRY(check_launch(h));
  if (h->ptr_mode == IGV_PTR_HOST)
    IGV_CUDA(h, cudaMemcpyAsync(trace_out, dev, sizeof(double) * h->B, cudaMemcpyDeviceToHost, h->stream));
  return IGV_OK;
}

igv_status igv_fence_record(igv_batch* h, int fence) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || fence < 0 || fence >= 4) return IGV_ERR_INVALID;
  IGV_CUDA(h, cudaEventRecord(h->fences[fence], h->stream));
  return IGV_OK;
}
igv_status igv_fence_wait(igv_batch* h, int fence) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || fence < 0 || fence >= 4) return IGV_ERR_INVALID;
  IGV_CUDA(h, cudaEventSynchronize(h->fences[fence]));
  return IGV_OK;
}

igv_status igv_profile_enable(igv_batch* h, int on) {
  IgvDeviceGuard dev_guard_(h);
  if (!h) return IGV_ERR_INVALID;
  h->prof_on = on != 0;
  return IGV_OK;
}

igv_status igv_profile_read(igv_batch* h, double* ms_out, long long* launches_out, int reset) {
  IgvDeviceGuard dev_guard_(h);
  if (!h || !ms_out || !launches_out) return IGV_ERR_INVALID;
  IGV_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto& ev : h->prof_events) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev.e0, ev.e1);
    h->prof_ms[ev.kind] += ms;
    h->prof_cnt[ev.kind] += 1;
    cudaEventDestroy(ev.e0);
    cudaEventDestroy(ev.e1);
  }
  h->prof_events.clear();
  for (int k = 0; k < IGV_K_COUNT; ++k) { ms_out[k] = h->prof_ms[k]; launches_out[k] = h->prof_cnt[k]; }
  if (reset) for (int k = 0; k < IGV_K_COUNT; ++k) { h->prof_ms[k] = 0.0; h->prof_cnt[k] = 0; }
  return IGV_OK;
}

}  // extern "C"
