// Track table: the MapServer of every sequence on the device.
//
// Reference: MapServer.h:69-134 (FeatureInfo: id, observation maps keyed by clone timestamp, anchor, _isToMarg, _isTri,
// landmark), MapServerManager.cpp:101-273 (collect*Meas, markMarg*Features), :275-341 (triangulateFeatureInfo*),
// :454-490 (eraseInvalidFeatures), the track selections of RemoveLostUpdate.cpp:45-59 / SwMargUpdate.cpp:61-85 /
// KeyframeUpdate.cpp:455-480, SwMargUpdate.cpp:191-259 and KeyframeUpdate.cpp:251-327 (clean*ObsAtMargTime,
// changeMSCKFAnchor), fed from the wire format of feature_tracker/msg/{Mono,Stereo}Meas.msg (uint64 id, float64 uv).
//
// Layout: SoA over T table entries per sequence; the observation map of a track is a 64-bit mask over PHYSICAL clone
// columns plus a T x C x rho array of image coordinates, so the sliding window never moves data (the handle maps
// window slots to columns, IgvTrkCols). This is integer / byte work bound by HBM latency, not by arithmetic: one CTA
// per sequence where an ordered decision is needed (message order in collect, id order in gather -- the iteration
// order of std::map<int, ...>), one thread per track everywhere else. Every result is independent of scheduling.
#include "igv_internal.h"

// Dynamic shared memory. (tests/emul/ compiles this file for the CPU with its own definition, to run the kernels
// against the oracle without a GPU; IGV_EMULATE is never defined in the library build.)
#ifndef IGV_EMULATE
#define IGV_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace {

__device__ __forceinline__ int narrow_id(unsigned long long v) {  // `int _id = msg.id` (MapServer.cpp:24)
  return (int)(unsigned int)(v & 0xffffffffull);
}

struct TrkPtrs {
  int T, C, rho, B;
  int* id; unsigned long long* mask; unsigned char* st; int* anchor; double* pf; double* pf_fej; double* obs;
  int* flags;
};

__global__ void k_trk_reset(TrkPtrs p) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  p.id[g] = 0; p.mask[g] = 0ull; p.st[g] = 0; p.anchor[g] = -1;
  for (int i = 0; i < 3; ++i) { p.pf[3 * g + i] = 0.0; p.pf_fej[3 * g + i] = 0.0; }
}

// Dynamic shared memory of k_trk_collect / k_trk_gather (also used by tests/emul, so the two cannot drift apart).
inline size_t trk_collect_smem(int T, int M) { return sizeof(int) * (3 * (size_t)T + 2 * (size_t)M) + (size_t)M; }
inline size_t trk_gather_smem(int T, int F) {
  return sizeof(int) * (2 * (size_t)T + (((size_t)F + 1) & ~size_t(1))) + sizeof(unsigned long long) * (size_t)F;
}

// One CTA per sequence. Phases: (1) warp 0 compacts the table into a list of used entries (id, entry) and a list of
// free entries with ballots, the other warps bring the message ids to shared memory; (2) per measurement: first
// occurrence of its id in the message? existing entry? -- both loops are branch-free with warp-uniform trip counts
// (a `break` here lets the lanes of a warp drift apart for good: the first version of this kernel spent 0.7 ms in
// 32 serialised copies of the scan, profiles/r01_track_table.md); (3) warp 0 hands free entries to the new ids in
// message order; (4) per measurement: write the observation (at most one writer per entry after (2)).
__global__ void __launch_bounds__(256) k_trk_collect(TrkPtrs p, IgvTrkCols cols, const int* __restrict__ n_meas, int M,
                                                     const unsigned long long* __restrict__ ids,
                                                     const double* __restrict__ uv) {
  IGV_DYN_SMEM(int, smem_i);
  int* s_uid = smem_i;                // T   ids of the used entries, compacted
  int* s_ut = s_uid + p.T;            // T   their table entries
  int* s_free = s_ut + p.T;           // T   free entries, ascending
  int* s_mid = s_free + p.T;          // M   message ids narrowed to int
  int* s_slot = s_mid + M;            // M   table entry of each measurement (-1: none)
  unsigned char* s_flag = reinterpret_cast<unsigned char*>(s_slot + M);   // M   bit 0: first occurrence, bit 1: new id
  __shared__ int s_nused, s_nfree;
  const int b = blockIdx.x, tid = threadIdx.x;
  const long tb = (long)b * p.T;
  int n = n_meas[b];
  n = n < 0 ? 0 : (n > M ? M : n);
  if (tid < 32) {
    const unsigned lt = (1u << tid) - 1u;
    int nused = 0, nfree = 0;
    for (int t0 = 0; t0 < p.T; t0 += 32) {
      const int t = t0 + tid;
      const bool in = t < p.T;
      const bool us = in && (p.st[tb + t] & IGV_TRK_USED);
      const unsigned mu = __ballot_sync(0xffffffffu, us);
      const unsigned mf = __ballot_sync(0xffffffffu, in && !us);
      if (us) { const int k = nused + __popc(mu & lt); s_uid[k] = p.id[tb + t]; s_ut[k] = t; }
      if (in && !us) s_free[nfree + __popc(mf & lt)] = t;
      nused += __popc(mu);
      nfree += __popc(mf);
    }
    if (tid == 0) { s_nused = nused; s_nfree = nfree; }
  } else {
    for (int i = tid - 32; i < n; i += blockDim.x - 32) s_mid[i] = narrow_id(ids[(long)b * M + i]);
  }
  __syncthreads();
  const int nused = s_nused, nfree = s_nfree;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + tid;
    const bool act = i < n;
    const int mid = act ? s_mid[i] : 0;
    const int jend = min(n, (i | 31) + 1);          // warp-uniform: the largest i of this warp, + 1
    int dup = 0;
    for (int j = 0; j < jend; ++j) dup |= (j < i && s_mid[j] == mid) ? 1 : 0;
    int slot = -1;
    for (int k = 0; k < nused; ++k) slot = (s_uid[k] == mid) ? s_ut[k] : slot;   // ids are unique in the table
    if (act) {
      const int eff = dup ? 0 : 1;
      if (!eff) slot = -1;
      s_slot[i] = slot;
      s_flag[i] = (unsigned char)(eff | ((eff && slot < 0) ? 2 : 0));
    }
  }
  __syncthreads();
  if (tid < 32) {
    const unsigned lt = (1u << tid) - 1u;
    int nnew = 0;
    bool over = false;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + tid;
      const bool nw = i < n && (s_flag[i] & 2);
      const unsigned m = __ballot_sync(0xffffffffu, nw);
      if (nw) {
        const int k = nnew + __popc(m & lt);
        if (k < nfree) s_slot[i] = s_free[k]; else { s_slot[i] = -1; over = true; }
      }
      nnew += __popc(m);
    }
    if (over) atomicOr(&p.flags[b], IGV_FLAG_TRACKS_FULL);
  }
  __syncthreads();
  const unsigned long long bit = 1ull << cols.cur_col;
  for (int i = tid; i < n; i += blockDim.x) {
    if (!(s_flag[i] & 1)) continue;
    const int slot = s_slot[i];
    if (slot < 0) continue;   // table full
    const long g = tb + slot;
    bool fresh = (s_flag[i] & 2) != 0;
    unsigned long long m = 0ull;
    if (!fresh) { m = p.mask[g]; if (m == 0ull) fresh = true; }   // empty observation map: re-initialised (:113-123)
    if (fresh) {
      p.id[g] = s_mid[i];
      p.st[g] = IGV_TRK_USED;
      p.mask[g] = bit;
      p.anchor[g] = cols.cur_col;
      for (int k = 0; k < 3; ++k) { p.pf[3 * g + k] = 0.0; p.pf_fej[3 * g + k] = 0.0; }
    } else {
      if (m & bit) continue;                                      // ":126-130 already in obs, skip adding"
      p.mask[g] = m | bit;
      p.st[g] = p.st[g] & ~IGV_TRK_TO_MARG;
    }
    double* o = p.obs + ((size_t)g * p.C + cols.cur_col) * p.rho;
    const double* z = uv + ((size_t)b * M + i) * p.rho;
    for (int k = 0; k < p.rho; ++k) o[k] = z[k];
  }
}

__global__ void k_trk_mark_lost(TrkPtrs p, IgvTrkCols cols) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  const unsigned char st = p.st[g];
  if (!(st & IGV_TRK_USED)) return;
  const bool seen = cols.cur_col >= 0 && ((p.mask[g] >> cols.cur_col) & 1ull);
  if (!seen) p.st[g] = st | IGV_TRK_TO_MARG;
}

// One CTA per sequence: select, order by id (rank by counting: ids are unique inside a sequence), emit.
__global__ void __launch_bounds__(256) k_trk_gather(TrkPtrs p, IgvTrkCols cols, IgvTrkGatherLaunch a) {
  IGV_DYN_SMEM(int, smem_i);
  int* s_sel_t = smem_i;                     // T
  int* s_sel_id = s_sel_t + p.T;             // T
  int* s_rank_t = s_sel_id + p.T;            // F
  unsigned long long* s_rank_mask = reinterpret_cast<unsigned long long*>(s_rank_t + ((a.F + 1) & ~1));  // F
  __shared__ int s_cnt;
  const int b = blockIdx.x, tid = threadIdx.x;
  const long tb = (long)b * p.T;
  if (tid == 0) s_cnt = 0;
  for (int f = tid; f < a.F; f += blockDim.x) { s_rank_t[f] = -1; s_rank_mask[f] = 0ull; }
  __syncthreads();
  for (int t = tid; t < p.T; t += blockDim.x) {
    const unsigned char st = p.st[tb + t];
    if (!(st & IGV_TRK_USED)) continue;
    bool sel;
    if (a.rule == IGV_TRK_LOST) sel = (st & IGV_TRK_TO_MARG) != 0;
    else sel = (p.mask[tb + t] & a.sel_cols) == a.sel_cols;
    if (sel) {
      const int k = atomicAdd(&s_cnt, 1);
      s_sel_t[k] = t;
      s_sel_id[k] = p.id[tb + t];
    }
  }
  __syncthreads();
  const int n = s_cnt;
  for (int k = tid; k < n; k += blockDim.x) {
    const int my = s_sel_id[k];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (s_sel_id[j] < my) ? 1 : 0;
    if (rank < a.F) { s_rank_t[rank] = s_sel_t[k]; s_rank_mask[rank] = p.mask[tb + s_sel_t[k]]; }
  }
  __syncthreads();
  if (tid == 0) {
    a.n_sel[b] = n < a.F ? n : a.F;
    if (n > a.F) atomicOr(&p.flags[b], IGV_FLAG_GATHER_CUT);
  }
  const long fb = (long)b * a.F;
  for (int f = tid; f < a.F; f += blockDim.x) {
    const int t = s_rank_t[f];
    int entry = -1, tid_out = 0, an = 0, dof = 1;
    unsigned char ok = 0;
    if (t >= 0) {
      const int nobs = __popcll(s_rank_mask[f]);
      entry = t;
      tid_out = p.id[tb + t];
      const int ac = p.anchor[tb + t];
      const int as = (ac >= 0) ? (int)cols.slot_of_col[ac] : -1;
      if (a.rule == IGV_TRK_LOST) {
        dof = nobs - 1 > 1 ? nobs - 1 : 1;
        ok = nobs >= a.min_obs ? 1 : 0;
      } else {
        dof = a.dof_fixed > 0 ? a.dof_fixed : (a.n_selected - 1 > 1 ? a.n_selected - 1 : 1);
        ok = 1;
      }
      if (as < 0) ok = 0; else an = as;   // anchor clone no longer in the window: nothing to linearise about
    }
    a.entry[fb + f] = entry;
    if (a.track_id) a.track_id[fb + f] = tid_out;
    a.anchor_slot[fb + f] = an;
    a.dof[fb + f] = dof;
    a.feat_ok[fb + f] = ok;
  }
  const int rows = a.F * a.SW;
  for (int e = tid; e < rows; e += blockDim.x) {
    const int f = e / a.SW, s = e - f * a.SW;
    const int t = s_rank_t[f];
    const int col = (s < cols.n_slots) ? (int)cols.col_of_slot[s] : -1;
    const bool has = t >= 0 && col >= 0 && ((s_rank_mask[f] >> col) & 1ull);
    const size_t oe = (size_t)fb * a.SW + e;
    a.mask_all[oe] = has ? 1 : 0;
    a.mask_upd[oe] = (has && (a.rule == IGV_TRK_LOST || ((a.sel_cols >> col) & 1ull))) ? 1 : 0;
    const double* src = has ? p.obs + ((size_t)(tb + t) * p.C + col) * p.rho : nullptr;
    for (int k = 0; k < p.rho; ++k) a.obs[oe * p.rho + k] = has ? src[k] : 0.0;
  }
}

__global__ void k_trk_commit_tri(TrkPtrs p, int F, const int* __restrict__ entry, const double* __restrict__ pf,
                                 const unsigned char* __restrict__ ok, unsigned char* feat_ok) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * F) return;
  const int b = (int)(g / F);
  const int t = entry[g];
  const bool good = t >= 0 && ok[g];
  if (feat_ok) feat_ok[g] = (feat_ok[g] && good) ? 1 : 0;
  if (!good) return;
  const long e = (long)b * p.T + t;
  const unsigned char st = p.st[e];
  for (int i = 0; i < 3; ++i) p.pf[3 * e + i] = pf[3 * g + i];              // setValuePosXyz (:297,:303)
  if (!(st & IGV_TRK_TRI)) {
    for (int i = 0; i < 3; ++i) p.pf_fej[3 * e + i] = pf[3 * g + i];        // setFejPosXyz at the first success (:296)
    p.st[e] = st | IGV_TRK_TRI;
  }
}

__device__ __forceinline__ void erase_entry(const TrkPtrs& p, long e) {
  p.st[e] = 0; p.mask[e] = 0ull; p.anchor[e] = -1;
}

__global__ void k_trk_erase(TrkPtrs p, int F, const int* __restrict__ entry) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * F) return;
  const int t = entry[g];
  if (t < 0) return;
  erase_entry(p, (long)(g / F) * p.T + t);
}

__global__ void k_trk_clean(TrkPtrs p, unsigned long long bits, int erase_empty) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  if (!(p.st[g] & IGV_TRK_USED)) return;
  const unsigned long long m = p.mask[g] & ~bits;
  p.mask[g] = m;
  if (erase_empty) {
    if (m == 0ull) erase_entry(p, g);
  } else {
    const int ac = p.anchor[g];
    if (ac >= 0 && ((bits >> ac) & 1ull)) p.anchor[g] = -1;   // the anchor clone left the window
  }
}

// depth of the landmark in the camera of window slot s:  (R^T (pf - p)).z, R_c2w row-major in the mean mirror
__device__ __forceinline__ double depth_in_slot(const double* Xb, int s, const double* pf) {
  const double* c = Xb + IGV_X_CORE + 12 * s;
  const double d0 = pf[0] - c[9], d1 = pf[1] - c[10], d2 = pf[2] - c[11];
  return c[2] * d0 + c[5] * d1 + c[8] * d2;
}

__global__ void k_trk_change_anchor(TrkPtrs p, IgvTrkCols cols, const double* __restrict__ X, int xsize,
                                    unsigned long long old_cols, double min_depth) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  const unsigned char st = p.st[g];
  if (!(st & IGV_TRK_USED)) return;
  const int ac = p.anchor[g];
  if (ac < 0 || !((old_cols >> ac) & 1ull)) return;
  if (!(st & IGV_TRK_TRI)) { erase_entry(p, g); return; }
  const int b = (int)(g / p.T);
  const double z = depth_in_slot(X + (size_t)b * xsize, cols.n_slots - 1, p.pf + 3 * g);
  if (z <= min_depth) { erase_entry(p, g); return; }
  p.anchor[g] = cols.cur_col;
}

__global__ void k_trk_erase_invalid(TrkPtrs p, IgvTrkCols cols, const double* __restrict__ X, int xsize,
                                    double min_depth) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  const unsigned char st = p.st[g];
  if (!(st & IGV_TRK_USED) || !(st & IGV_TRK_TRI)) return;
  const int ac = p.anchor[g];
  const int as = ac >= 0 ? (int)cols.slot_of_col[ac] : -1;
  if (as < 0) { erase_entry(p, g); return; }                  // getAnchoredPose() == nullptr (:468-472)
  const int b = (int)(g / p.T);
  if (depth_in_slot(X + (size_t)b * xsize, as, p.pf + 3 * g) <= min_depth) erase_entry(p, g);
}

__global__ void k_trk_dump(TrkPtrs p, IgvTrkCols cols, igv_track_dump d) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)p.B * p.T) return;
  const unsigned char st = p.st[g];
  const bool used = (st & IGV_TRK_USED) != 0;
  const unsigned long long m = used ? p.mask[g] : 0ull;
  if (d.id) d.id[g] = used ? p.id[g] : 0;
  if (d.used) d.used[g] = used ? 1 : 0;
  if (d.to_marg) d.to_marg[g] = (used && (st & IGV_TRK_TO_MARG)) ? 1 : 0;
  if (d.is_tri) d.is_tri[g] = (used && (st & IGV_TRK_TRI)) ? 1 : 0;
  if (d.slot_mask) {
    unsigned long long sm = 0ull;
    for (int s = 0; s < cols.n_slots; ++s) if ((m >> cols.col_of_slot[s]) & 1ull) sm |= 1ull << s;
    d.slot_mask[g] = sm;
  }
  if (d.anchor_slot) {
    const int ac = used ? p.anchor[g] : -1;
    d.anchor_slot[g] = ac >= 0 ? (int)cols.slot_of_col[ac] : -1;
  }
  const bool tri = used && (st & IGV_TRK_TRI);
  if (d.pf) for (int i = 0; i < 3; ++i) d.pf[3 * g + i] = tri ? p.pf[3 * g + i] : 0.0;
  if (d.pf_fej) for (int i = 0; i < 3; ++i) d.pf_fej[3 * g + i] = tri ? p.pf_fej[3 * g + i] : 0.0;
  if (d.obs) {
    for (int s = 0; s < d.obs_slots; ++s) {
      const int col = s < cols.n_slots ? (int)cols.col_of_slot[s] : -1;
      const bool has = col >= 0 && ((m >> col) & 1ull);
      for (int k = 0; k < p.rho; ++k)
        d.obs[((size_t)g * d.obs_slots + s) * p.rho + k] = has ? p.obs[((size_t)g * p.C + col) * p.rho + k] : 0.0;
    }
  }
}

__global__ void k_trk_count(TrkPtrs p, int* n_tracks) {
  const int b = blockIdx.x;
  int c = 0;
  for (int t = threadIdx.x; t < p.T; t += blockDim.x) c += (p.st[(long)b * p.T + t] & IGV_TRK_USED) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  __shared__ int s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    n_tracks[b] = tot;
  }
}

#if !defined(IGV_EMULATE) || defined(IGV_EMULATE_LAUNCHERS)   // tests/emul: kernels only, or (full model) launchers too
TrkPtrs ptrs(const igv_batch* h) {
  TrkPtrs p;
  p.T = h->trk.T; p.C = h->trk.C; p.rho = h->rho; p.B = h->B;
  p.id = h->trk.id; p.mask = h->trk.mask; p.st = h->trk.st; p.anchor = h->trk.anchor; p.pf = h->trk.pf;
  p.pf_fej = h->trk.pf_fej; p.obs = h->trk.obs; p.flags = h->flags;
  return p;
}

inline unsigned per_track_grid(const igv_batch* h) { return (unsigned)(((long)h->B * h->trk.T + 255) / 256); }

#endif  // IGV_EMULATE

}  // namespace

// ---- host metadata: window slot <-> physical column (shared by the C-ABI and by tests/emul) ----------------------
IgvTrkCols igv_trk_cols(const IgvTrackTable& t) {
  IgvTrkCols c;
  for (int i = 0; i < IGV_MAX_CLONES; ++i) { c.col_of_slot[i] = -1; c.slot_of_col[i] = -1; }
  c.n_slots = (int)t.col_of_slot.size();
  for (int s = 0; s < c.n_slots; ++s) {
    c.col_of_slot[s] = (signed char)t.col_of_slot[s];
    c.slot_of_col[t.col_of_slot[s]] = (signed char)s;
  }
  c.cur_col = c.n_slots > 0 ? t.col_of_slot[c.n_slots - 1] : -1;
  return c;
}
// A clone was appended to the window: it takes the lowest free column. Returns it, or -1 when every column is taken.
int igv_trk_col_alloc(IgvTrackTable& t) {
  unsigned long long used = 0ull;
  for (int c : t.col_of_slot) used |= 1ull << c;
  int col = 0;
  while (col < t.C && ((used >> col) & 1ull)) ++col;
  if (col >= t.C) return -1;
  t.col_of_slot.push_back(col);
  return col;
}
// The clone at window slot `slot` left the window: later slots move down by one, its column becomes free.
int igv_trk_col_release(IgvTrackTable& t, int slot) {
  if (slot < 0 || slot >= (int)t.col_of_slot.size()) return -1;
  const int col = t.col_of_slot[slot];
  t.col_of_slot.erase(t.col_of_slot.begin() + slot);
  return col;
}
bool igv_trk_slot_bits(const IgvTrackTable& t, int n, const int* slots, unsigned long long* bits) {
  *bits = 0ull;
  for (int i = 0; i < n; ++i) {
    if (slots[i] < 0 || slots[i] >= (int)t.col_of_slot.size()) return false;
    *bits |= 1ull << t.col_of_slot[slots[i]];
  }
  return true;
}

#if !defined(IGV_EMULATE) || defined(IGV_EMULATE_LAUNCHERS)   // tests/emul: kernels only, or (full model) launchers too
void igv_launch_trk_reset(igv_batch* h) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_reset<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h));
  h->launches++;
}

void igv_launch_trk_collect(igv_batch* h, const int* n_meas, int meas_stride, const unsigned long long* ids,
                            const double* uv) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  const int T = h->trk.T, M = meas_stride;
  const size_t smem = trk_collect_smem(T, M);
  if (smem > 48 * 1024) IGV_SMEM_OPTIN((k_trk_collect), 220 * 1024);
  k_trk_collect<<<h->B, 256, smem, h->stream>>>(ptrs(h), igv_trk_cols(h->trk), n_meas, M, ids, uv);
  h->launches++;
}

void igv_launch_trk_mark_lost(igv_batch* h) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_mark_lost<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h), igv_trk_cols(h->trk));
  h->launches++;
}

void igv_launch_trk_gather(igv_batch* h, const IgvTrkGatherLaunch& g) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  const size_t smem = trk_gather_smem(h->trk.T, g.F);
  if (smem > 48 * 1024) IGV_SMEM_OPTIN((k_trk_gather), 220 * 1024);
  k_trk_gather<<<h->B, 256, smem, h->stream>>>(ptrs(h), igv_trk_cols(h->trk), g);
  h->launches++;
}

void igv_launch_trk_commit_tri(igv_batch* h, int F, const int* entry, const double* pf, const unsigned char* ok,
                               unsigned char* feat_ok) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_commit_tri<<<(unsigned)(((long)h->B * F + 255) / 256), 256, 0, h->stream>>>(ptrs(h), F, entry, pf, ok, feat_ok);
  h->launches++;
}

void igv_launch_trk_erase(igv_batch* h, int F, const int* entry) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_erase<<<(unsigned)(((long)h->B * F + 255) / 256), 256, 0, h->stream>>>(ptrs(h), F, entry);
  h->launches++;
}

void igv_launch_trk_clean(igv_batch* h, unsigned long long col_bits, int erase_empty) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_clean<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h), col_bits, erase_empty);
  h->launches++;
}

void igv_launch_trk_change_anchor(igv_batch* h, unsigned long long old_cols, double min_depth) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_change_anchor<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h), igv_trk_cols(h->trk), h->Xc(), h->xsize, old_cols,
                                                                min_depth);
  h->launches++;
}

void igv_launch_trk_erase_invalid(igv_batch* h, double min_depth) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_erase_invalid<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h), igv_trk_cols(h->trk), h->Xc(), h->xsize, min_depth);
  h->launches++;
}

void igv_launch_trk_dump(igv_batch* h, const igv_track_dump& d) {
  IgvProfScope prof_scope_(h, IGV_K_OTHER);
  k_trk_dump<<<per_track_grid(h), 256, 0, h->stream>>>(ptrs(h), igv_trk_cols(h->trk), d);
  h->launches++;
  if (d.n_tracks) {
    k_trk_count<<<h->B, 256, 0, h->stream>>>(ptrs(h), d.n_tracks);
    h->launches++;
  }
}
#endif  // IGV_EMULATE
