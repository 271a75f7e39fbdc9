// TEST INFRASTRUCTURE: runs the track-table kernels of ingvio_b200/csrc/k_tracks.cu (the same source the library
// compiles with nvcc) on the CPU through tests/emul/cuda_emul.h, behind a small C interface for ctypes
// (tests/emul/trk_emul.py). The host-side column bookkeeping (igv_trk_col_alloc / _release / _slot_bits /
// igv_trk_cols) is the library's own code as well. Built by tests/emul/trk_emul.py with g++ -std=c++20 -pthread.
#define IGV_EMULATE 1
#include "cuda_emul.h"

#include "../../ingvio_b200/csrc/k_tracks.cu"
#include "../../ingvio_b200/csrc/k_tri.cu"      // k_triangulate: one thread per track, plain FP64 arithmetic

namespace {
struct Emu {
  int B, T, C, rho, xsize;
  IgvTrackTable trk;
  std::vector<int> id, anchor, flags;
  std::vector<unsigned long long> mask;
  std::vector<unsigned char> st;
  std::vector<double> pf, pf_fej, obs, X;
  TrkPtrs ptrs() {
    TrkPtrs p;
    p.T = T; p.C = C; p.rho = rho; p.B = B;
    p.id = id.data(); p.mask = mask.data(); p.st = st.data(); p.anchor = anchor.data(); p.pf = pf.data();
    p.pf_fej = pf_fej.data(); p.obs = obs.data(); p.flags = flags.data();
    return p;
  }
  unsigned per_track_grid() const { return (unsigned)(((long)B * T + 255) / 256); }
};
}  // namespace

extern "C" {

void* emu_create(int B, int T, int C, int rho, int xsize) {
  Emu* e = new Emu();
  e->B = B; e->T = T; e->C = C; e->rho = rho; e->xsize = xsize;
  const size_t n = (size_t)B * T;
  e->id.assign(n, 0); e->anchor.assign(n, -1); e->flags.assign(B, 0); e->mask.assign(n, 0ull); e->st.assign(n, 0);
  e->pf.assign(3 * n, 0.0); e->pf_fej.assign(3 * n, 0.0); e->obs.assign(n * C * rho, 0.0);
  e->X.assign((size_t)B * xsize, 0.0);
  e->trk.T = T; e->trk.C = C;
  TrkPtrs p = e->ptrs();
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_reset(p); });
  return e;
}
void emu_destroy(void* h) { delete static_cast<Emu*>(h); }
int* emu_flags(void* h) { return static_cast<Emu*>(h)->flags.data(); }
void emu_set_X(void* h, const double* X) {
  Emu* e = static_cast<Emu*>(h);
  std::memcpy(e->X.data(), X, sizeof(double) * e->X.size());
}
int emu_on_augment(void* h) { return igv_trk_col_alloc(static_cast<Emu*>(h)->trk); }
// igv_api.cu: trk_on_marg_clone
void emu_on_marg(void* h, int slot) {
  Emu* e = static_cast<Emu*>(h);
  if (slot < 0 || slot >= (int)e->trk.col_of_slot.size()) return;
  TrkPtrs p = e->ptrs();
  const unsigned long long bits = 1ull << e->trk.col_of_slot[slot];
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_clean(p, bits, 0); });
  igv_trk_col_release(e->trk, slot);
}
void emu_collect(void* h, const int* n_meas, int M, const unsigned long long* ids, const double* uv) {
  Emu* e = static_cast<Emu*>(h);
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  const size_t smem = trk_collect_smem(e->T, M);
  emul::launch(e->B, 256, smem, [&] { k_trk_collect(p, c, n_meas, M, ids, uv); });
}
void emu_mark_lost(void* h) {
  Emu* e = static_cast<Emu*>(h);
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_mark_lost(p, c); });
}
int emu_gather(void* h, int rule, int n_selected, const int* selected_slots, int min_obs, int dof_fixed, int F, int SW,
               int* entry, int* n_sel, int* track_id, double* obs, unsigned char* mask_all, unsigned char* mask_upd,
               int* anchor_slot, int* dof, unsigned char* feat_ok) {
  Emu* e = static_cast<Emu*>(h);
  IgvTrkGatherLaunch g{};
  g.rule = rule; g.n_selected = n_selected; g.min_obs = min_obs; g.dof_fixed = dof_fixed; g.F = F; g.SW = SW;
  if (rule == IGV_TRK_SEEN_AT && !igv_trk_slot_bits(e->trk, n_selected, selected_slots, &g.sel_cols)) return 1;
  g.entry = entry; g.n_sel = n_sel; g.track_id = track_id; g.obs = obs; g.mask_all = mask_all; g.mask_upd = mask_upd;
  g.anchor_slot = anchor_slot; g.dof = dof; g.feat_ok = feat_ok;
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  const size_t smem = trk_gather_smem(e->T, F);
  emul::launch(e->B, 256, smem, [&] { k_trk_gather(p, c, g); });
  return 0;
}
void emu_commit_tri(void* h, int F, const int* entry, const double* pf, const unsigned char* ok, unsigned char* feat_ok) {
  Emu* e = static_cast<Emu*>(h);
  TrkPtrs p = e->ptrs();
  emul::launch((unsigned)(((long)e->B * F + 255) / 256), 256, 0, [&] { k_trk_commit_tri(p, F, entry, pf, ok, feat_ok); });
}
void emu_erase(void* h, int F, const int* entry) {
  Emu* e = static_cast<Emu*>(h);
  TrkPtrs p = e->ptrs();
  emul::launch((unsigned)(((long)e->B * F + 255) / 256), 256, 0, [&] { k_trk_erase(p, F, entry); });
}
int emu_clean(void* h, int n, const int* slots) {
  Emu* e = static_cast<Emu*>(h);
  unsigned long long bits;
  if (!igv_trk_slot_bits(e->trk, n, slots, &bits)) return 1;
  TrkPtrs p = e->ptrs();
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_clean(p, bits, 1); });
  return 0;
}
int emu_change_anchor(void* h, int n, const int* slots, double min_depth) {
  Emu* e = static_cast<Emu*>(h);
  unsigned long long bits;
  if (!igv_trk_slot_bits(e->trk, n, slots, &bits)) return 1;
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  const double* X = e->X.data();
  const int xs = e->xsize;
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_change_anchor(p, c, X, xs, bits, min_depth); });
  return 0;
}
void emu_erase_invalid(void* h, double min_depth) {
  Emu* e = static_cast<Emu*>(h);
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  const double* X = e->X.data();
  const int xs = e->xsize;
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_erase_invalid(p, c, X, xs, min_depth); });
}
// igv_launch_triangulate (k_tri.cu) on the clone poses last given with emu_set_X
void emu_triangulate(void* h, int n_clones, int F, int obs_slots, const double* obs, const unsigned char* mask, const int* anchor,
                     const igv_tri_params* prm, const double* Rc, const double* pc, double* pf_out, unsigned char* ok_out) {
  Emu* e = static_cast<Emu*>(h);
  TriArgs a;
  a.X = e->X.data(); a.xsize = e->xsize; a.n_clones = n_clones;
  a.F = F; a.obs_slots = obs_slots; a.rho = e->rho;
  a.obs = obs; a.mask = mask; a.anchor = anchor; a.prm = *prm;
  for (int i = 0; i < 9; ++i) a.Rc[i] = Rc[i];
  for (int i = 0; i < 3; ++i) a.pc[i] = pc[i];
  a.pf_out = pf_out; a.ok_out = ok_out; a.B = e->B;
  const long n = (long)e->B * F;
  emul::launch((unsigned)((n + 127) / 128), 128, 0, [&] { k_triangulate(a); });
}
void emu_dump(void* h, int obs_slots, int* id, unsigned char* used, unsigned char* to_marg, unsigned char* is_tri,
              unsigned long long* slot_mask, int* anchor_slot, double* pf, double* pf_fej, double* obs, int* n_tracks) {
  Emu* e = static_cast<Emu*>(h);
  igv_track_dump d{};
  d.obs_slots = obs_slots; d.id = id; d.used = used; d.to_marg = to_marg; d.is_tri = is_tri; d.slot_mask = slot_mask;
  d.anchor_slot = anchor_slot; d.pf = pf; d.pf_fej = pf_fej; d.obs = obs; d.n_tracks = n_tracks;
  TrkPtrs p = e->ptrs();
  IgvTrkCols c = igv_trk_cols(e->trk);
  emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_dump(p, c, d); });
  if (n_tracks) emul::launch(e->B, 256, 0, [&] { k_trk_count(p, n_tracks); });
}
}
