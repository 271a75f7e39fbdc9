// TEST INFRASTRUCTURE: the subset of the C-ABI (include/ingvio_b200.h) that the C++ map-server mirror
// (ingvio_b200/host/ingvio_map_server.hpp) calls, backed by the track-table kernel source executed on the CPU
// (tests/emul/trk_emul.cpp). It exists so that tests/cpp/test_map_server_mirror.cpp -- the reference's MapServer gtest
// restated against the mirror -- can check the mirror's own logic on a machine without a GPU; the GPU test links the same
// source against libingvio_b200.so. Covariance / mean algebra is NOT provided here (those entry points are exercised on the
// GPU only: here igv_propagate_imu only drifts the position, igv_triangulate IS the device kernel (k_tri.cu) run on the CPU,
// igv_msckf_update and igv_set_chi2_table do nothing and igv_cov_get returns the
// identity -- enough for tests/cpp/test_updaters_frames.cpp to exercise the wiring of the updater mirror, nothing more); B = 1.
#include "trk_emul.cpp"

// The opaque handle of the C-ABI is the library's own igv_batch (igv_internal.h; cfg / err / launches are reused), extended
// by what the shim needs. No CUDA call is ever made on it.
struct Shim : igv_batch {
  void* emu = nullptr;           // created by igv_tracks_create (the table size is not known before)
  int n_clones = 0;
  std::vector<double> Xh;        // mean mirror: only the clone poses are kept
  std::vector<double> last_gyro, last_accel, last_dt;   // arguments of the last igv_propagate_imu (igv_shim_last_propagate)
};
static Shim* S(igv_batch* h) { return static_cast<Shim*>(h); }
static const Shim* S(const igv_batch* h) { return static_cast<const Shim*>(h); }

extern "C" {

igv_status igv_create(const igv_config* cfg, igv_batch** out) {
  if (!cfg || !out || cfg->batch != 1) return IGV_ERR_INVALID;
  Shim* h = new Shim();
  h->cfg = *cfg;
  h->Xh.assign(IGV_X_CORE + 12 * (size_t)cfg->max_clones, 0.0);
  *out = h;
  return IGV_OK;
}
igv_status igv_destroy(igv_batch* hb) {
  Shim* h = S(hb);
  if (h) { if (h->emu) emu_destroy(h->emu); delete h; }
  return IGV_OK;
}
const char* igv_last_error(const igv_batch* h) { return h ? h->err.c_str() : "null handle"; }
long long igv_launch_count(const igv_batch* h) { return h ? h->launches : 0; }
igv_status igv_set_params(igv_batch* hb, const igv_params* p) {
  if (!hb || !p) return IGV_ERR_INVALID;
  for (int i = 0; i < 9; ++i) hb->params.Rc[i] = p->T_cl2cr_R[i];
  for (int i = 0; i < 3; ++i) hb->params.pc[i] = p->T_cl2cr_p[i];
  return IGV_OK;
}
igv_status igv_state_init(igv_batch* hb, const double* R, const double* p, const double* v, const double* bg, const double* ba,
                          const double* Re, const double* pe, const double*) {
  Shim* h = S(hb);
  if (!h) return IGV_ERR_INVALID;
  h->n_clones = 0;
  std::fill(h->Xh.begin(), h->Xh.end(), 0.0);
  for (int i = 0; i < 9; ++i) { h->Xh[i] = R[i]; h->Xh[21 + i] = Re[i]; }
  for (int i = 0; i < 3; ++i) { h->Xh[9 + i] = p[i]; h->Xh[12 + i] = v[i]; h->Xh[15 + i] = bg[i]; h->Xh[18 + i] = ba[i]; h->Xh[30 + i] = pe[i]; }
  if (h->emu) {   // a new State starts with an empty map (igv_api.cu: igv_state_init)
    Emu* e = static_cast<Emu*>(h->emu);
    e->trk.col_of_slot.clear();
    TrkPtrs p = e->ptrs();
    emul::launch(e->per_track_grid(), 256, 0, [&] { k_trk_reset(p); });
  }
  ++h->launches;
  return IGV_OK;
}
int igv_dim(const igv_batch* h) { return h ? 21 + 6 * S(h)->n_clones : -1; }
igv_status igv_augment_clone_cov(igv_batch* hb, const double*, const double* clone_R, const double* clone_p) {
  Shim* h = S(hb);
  if (!h) return IGV_ERR_INVALID;
  if (h->n_clones >= h->cfg.max_clones) { h->err = "sliding window is full"; return IGV_ERR_CAPACITY; }
  double* c = h->Xh.data() + IGV_X_CORE + 12 * h->n_clones;
  for (int i = 0; i < 9; ++i) c[i] = clone_R ? clone_R[i] : 0.0;
  for (int i = 0; i < 3; ++i) c[9 + i] = clone_p ? clone_p[i] : 0.0;
  ++h->n_clones;
  if (h->emu) emu_on_augment(h->emu);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_marginalize(igv_batch* hb, int idx) {
  Shim* h = S(hb);
  if (!h) return IGV_ERR_INVALID;
  const int slot = (idx - 21) / 6;
  if (idx < 21 || (idx - 21) % 6 != 0 || slot >= h->n_clones) { h->err = "Marg is not in the current state"; return IGV_ERR_STATE; }
  if (h->emu) emu_on_marg(h->emu, slot);
  double* base = h->Xh.data() + IGV_X_CORE;
  for (int s = slot; s + 1 < h->n_clones; ++s) std::memcpy(base + 12 * s, base + 12 * (s + 1), sizeof(double) * 12);
  --h->n_clones;
  ++h->launches;
  return IGV_OK;
}
int igv_tracks_capacity(const igv_batch* h) { return (h && S(h)->emu) ? static_cast<Emu*>(S(h)->emu)->T : 0; }
igv_status igv_tracks_create(igv_batch* hb, int max_tracks) {
  Shim* h = S(hb);
  if (!h || max_tracks < 1 || max_tracks > 4096) return IGV_ERR_INVALID;
  if (h->emu) { h->err = "track table already created"; return IGV_ERR_STATE; }
  h->emu = emu_create(1, max_tracks, h->cfg.max_clones, h->cfg.stereo ? 4 : 2, (int)h->Xh.size());
  for (int s = 0; s < h->n_clones; ++s) emu_on_augment(h->emu);   // igv_tracks_create: existing clones take columns 0..n-1
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_collect(igv_batch* hb, const int* n_meas, int meas_stride, const unsigned long long* ids, const double* uv) {
  Shim* h = S(hb);
  if (!h || !h->emu) { if (h) h->err = "track table not created (igv_tracks_create)"; return IGV_ERR_STATE; }
  if (!n_meas || meas_stride < 0 || meas_stride > 4096) return IGV_ERR_INVALID;
  if (meas_stride == 0) return IGV_OK;
  if (h->n_clones == 0) { h->err = "[FeatureInfoManager]: Meas timestamp not in sw!"; return IGV_ERR_STATE; }
  emu_collect(h->emu, n_meas, meas_stride, ids, uv);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_mark_lost(igv_batch* hb) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  emu_mark_lost(h->emu);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_erase_invalid(igv_batch* hb, double min_depth) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  if (h->n_clones == 0) return IGV_OK;
  emu_set_X(h->emu, h->Xh.data());
  emu_erase_invalid(h->emu, min_depth);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_get(igv_batch* hb, const igv_track_dump* d) {
  Shim* h = S(hb);
  if (!h || !h->emu || !d) return IGV_ERR_STATE;
  if (d->obs && d->obs_slots < h->n_clones) return IGV_ERR_INVALID;
  emu_dump(h->emu, d->obs_slots, d->id, d->used, d->to_marg, d->is_tri, d->slot_mask, d->anchor_slot, d->pf, d->pf_fej, d->obs,
           d->n_tracks);
  ++h->launches;
  return IGV_OK;
}

// ---- stubs / table calls used by the updater mirror (tests/cpp/test_updaters_frames.cpp) ----
int igv_state_size(const igv_batch* h) { return h ? (int)S(h)->Xh.size() : -1; }
igv_status igv_state_get(igv_batch* hb, double* dst) {
  Shim* h = S(hb);
  if (!h || !dst) return IGV_ERR_INVALID;
  std::memcpy(dst, h->Xh.data(), sizeof(double) * h->Xh.size());
  return IGV_OK;
}
igv_status igv_cov_get(igv_batch* hb, double* dst, int ld) {
  Shim* h = S(hb);
  const int N = igv_dim(hb);
  if (!h || !dst || ld < N) return IGV_ERR_INVALID;
  for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) dst[(size_t)j * ld + i] = (i == j) ? 1.0 : 0.0;
  return IGV_OK;
}
igv_status igv_set_chi2_table(igv_batch*, const double*, int) { return IGV_OK; }
igv_status igv_propagate_imu(igv_batch* hb, int n_steps, const double* gyro, const double* accel, const double* dt) {
  Shim* h = S(hb);
  if (!h) return IGV_ERR_INVALID;
  h->last_gyro.assign(gyro, gyro + 3 * n_steps);
  h->last_accel.assign(accel, accel + 3 * n_steps);
  h->last_dt.assign(dt, dt + n_steps);
  for (int k = 0; k < n_steps; ++k) { h->Xh[9] += 2.0 * dt[k]; h->Xh[10] += 0.5 * dt[k]; }   // drift sideways: parallax
  ++h->launches;
  return IGV_OK;
}
// shim only: what the last igv_propagate_imu was given (tests/cpp/test_imu_buffer.cpp); clears the record
int igv_shim_last_propagate(igv_batch* hb, double* gyro, double* accel, double* dt, int cap) {
  Shim* h = S(hb);
  const int n = (int)h->last_dt.size();
  for (int i = 0; i < n && i < cap; ++i) {
    dt[i] = h->last_dt[i];
    for (int k = 0; k < 3; ++k) { gyro[3 * i + k] = h->last_gyro[3 * i + k]; accel[3 * i + k] = h->last_accel[3 * i + k]; }
  }
  h->last_dt.clear(); h->last_gyro.clear(); h->last_accel.clear();
  return n;
}
igv_status igv_msckf_update(igv_batch* hb, const igv_msckf_args* a) { if (!hb || !a) return IGV_ERR_INVALID; ++S(hb)->launches; return IGV_OK; }
igv_status igv_triangulate(igv_batch* hb, const igv_tri_args* a) {   // the real kernel (k_tri.cu) on the CPU
  Shim* h = S(hb);
  if (!h || !h->emu || !a || !a->obs || !a->obs_mask || !a->pf_out || !a->ok_out) return IGV_ERR_INVALID;
  if (a->n_feats < 0 || a->n_feats > h->cfg.max_feats || a->obs_slots < h->n_clones) return IGV_ERR_INVALID;
  emu_set_X(h->emu, h->Xh.data());
  emu_triangulate(h->emu, h->n_clones, a->n_feats, a->obs_slots, a->obs, a->obs_mask, a->anchor_slot, &a->prm, h->params.Rc,
                  h->params.pc, a->pf_out, a->ok_out);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_gather(igv_batch* hb, const igv_track_gather_args* a) {
  Shim* h = S(hb);
  if (!h || !h->emu || !a) return IGV_ERR_STATE;
  if (a->obs_slots < h->n_clones || a->n_feats < 1 || a->n_feats > h->cfg.max_feats) return IGV_ERR_INVALID;
  std::vector<int> ids((size_t)a->n_feats);
  if (emu_gather(h->emu, a->rule, a->n_selected, a->selected_slots, a->min_obs, a->dof_fixed, a->n_feats, a->obs_slots, a->track_entry,
                 a->n_sel, a->track_id ? a->track_id : ids.data(), a->obs, a->mask_all, a->mask_upd, a->anchor_slot, a->chi2_dof,
                 a->feat_ok) != 0) { h->err = "clone slot not in the sliding window"; return IGV_ERR_STATE; }
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_commit_tri(igv_batch* hb, int F, const int* entry, const double* pf, const unsigned char* ok, unsigned char* feat_ok) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  emu_commit_tri(h->emu, F, entry, pf, ok, feat_ok);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_erase(igv_batch* hb, int F, const int* entry) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  emu_erase(h->emu, F, entry);
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_clean_obs(igv_batch* hb, int n, const int* slots) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  if (emu_clean(h->emu, n, slots) != 0) { h->err = "clone slot not in the sliding window"; return IGV_ERR_STATE; }
  ++h->launches;
  return IGV_OK;
}
igv_status igv_tracks_change_anchor(igv_batch* hb, int n, const int* slots, double min_depth) {
  Shim* h = S(hb);
  if (!h || !h->emu) return IGV_ERR_STATE;
  emu_set_X(h->emu, h->Xh.data());
  if (emu_change_anchor(h->emu, n, slots, min_depth) != 0) { h->err = "clone slot not in the sliding window"; return IGV_ERR_STATE; }
  ++h->launches;
  return IGV_OK;
}
}
