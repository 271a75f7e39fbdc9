// Internal declarations shared by the C-ABI host code and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ingvio_b200.h"

#define IGV_MAX_CLONES 64
#define IGV_MAX_BLOCKS 100
#define IGV_MAX_LM 32
#define IGV_X_CORE 39          // packed mean: 33 core doubles + 6 GNSS scalars, then 12 per clone

// Variable layout of the batch, passed BY VALUE to kernels (all sequences share it).
struct IgvLayout {
  int N;                         // current state dimension
  int ld;                        // leading dimension of P
  int xsize;                     // doubles per sequence in the mean mirror
  int n_clones;
  int idx_clone[IGV_MAX_CLONES]; // Type::idx() of clone slot s (0 = oldest)
  int idx_gnss[6];               // Type::idx() of GPS,GLO,GAL,BDS,FS,YOF or -1
  int n_lm;                      // SLAM landmarks in the state
  int lm_off;                    // offset of the landmark values (3 each) in the packed mean
  int idx_lm[IGV_MAX_LM];        // Type::idx() of landmark slot l
  int lm_anchor[IGV_MAX_LM];     // anchor clone slot of landmark slot l
};

struct IgvBlocks {               // a var_order as (idx,size) blocks
  int n_blocks;
  int n;                         // sum of sizes
  int idx[IGV_MAX_BLOCKS];
  int size[IGV_MAX_BLOCKS];
};

struct IgvDevParams {
  double noise_g, noise_a, noise_bg, noise_ba, noise_cb, noise_cb_rw;
  double g[3];
  double Rc[9];                  // T_cl2cr rotation, row-major
  double pc[3];
};

// Track table (k_tracks.cu): the MapServer of every sequence, SoA over T table entries. Observations live in C
// PHYSICAL clone columns; `col_of_slot` (host metadata, passed by value like IgvLayout) maps window slots to them.
struct IgvTrackTable {
  int T = 0, C = 0;                        // entries per sequence, physical clone columns (= cfg.max_clones)
  int* id = nullptr;                       // B x T   MapServer key (message id narrowed to int)
  unsigned long long* mask = nullptr;      // B x T   bit c: observation at physical column c
  unsigned char* st = nullptr;             // B x T   IGV_TRK_USED | IGV_TRK_TO_MARG | IGV_TRK_TRI
  int* anchor = nullptr;                   // B x T   physical column of the anchor clone, -1 none
  double* pf = nullptr;                    // B x T x 3 landmark value (world)
  double* pf_fej = nullptr;                // B x T x 3
  double* obs = nullptr;                   // B x T x C x rho
  std::vector<int> col_of_slot;            // window slot -> physical column
};
enum { IGV_TRK_USED = 1, IGV_TRK_TO_MARG = 2, IGV_TRK_TRI = 4 };
struct IgvTrkCols {                        // by-value kernel argument
  int n_slots;                             // clones in the window
  int cur_col;                             // physical column of the newest clone, -1 if the window is empty
  signed char col_of_slot[IGV_MAX_CLONES];
  signed char slot_of_col[IGV_MAX_CLONES];
};

enum IgvVarKind { VK_SE23, VK_BG, VK_BA, VK_EXT, VK_GNSS, VK_CLONE, VK_OPAQUE, VK_LANDMARK };
struct IgvVar {
  IgvVarKind kind;
  int idx, size;
  int tag;                       // gnss type for VK_GNSS, anchor clone slot for VK_LANDMARK
};

// Test / tuning knobs, read from the environment ONCE per handle (igv_create), never on the launch path.
struct IgvKnobs {
  int fuse = -1;        // IGV_FUSE: 1 forces the fused per-track kernel, 0 forbids it
  int qr_cfg = 0;       // IGV_QR_CFG: forces one compression kernel (k_qr.cu launch_qr)
  int qr_split = 0;     // IGV_QR_SPLIT: forces the row split
  int gram_cfg = 0;     // IGV_GRAM_CFG: 1 forces the super-block Gram kernel
  int factor_cfg = 0;   // IGV_FACTOR_CFG: 1 forces the column-by-column factorisation
  int fuse_minw = 12;   // IGV_FUSE_MINW: fewest warps per CTA for which the fused per-track kernel is taken
  int tri_cfg = 0;      // IGV_TRI_CFG: 1 forces the thread-per-track triangulation kernel
  int tri_minb = 4;     // IGV_TRI_MINB: resident blocks per SM the group kernel is compiled for (3: 168 regs .. 6: 80 regs)
  int graph = -1;       // IGV_GRAPH: 0 disables CUDA-graph replay of igv_frame_step
  int feat_const = 1;   // IGV_FEAT_CONST: 0 forbids the compile-time-sized instances of the per-track kernel
  int ekf_t_small = 0, ekf_t_big = 0;       // (0 = automatic) IGV_EKF_T_SMALL / IGV_EKF_T_BIG: threads per CTA of k_ekf_update (rows <= 32 / above)
  double tc_pivot_tol = 1e-13; // IGV_TC_PIVOT_TOL: relative pivot threshold of the factorisation behind k_gram_tc (A/B runs)
  int tc_drain = 0;     // IGV_TC_DRAIN: 16-row stages per FP32 accumulation of k_gram_tc (0 = its default)
  int feat_ps = 1;      // IGV_FEAT_PS: 0 keeps the window's covariance block in global memory (A/B runs)
  int feat_warps = 0;   // IGV_FEAT_WARPS: cap on the warps per CTA of the unfused per-track kernel (0 = automatic)
  int prop_tma = 1;     // IGV_PROP_TMA: 0 forbids the bulk-TMA staging of the IMU transition records in k_propagate
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: one bit per device ordinal,
// set the first time a handle of that device launches the kernel (place the macro in the launcher: the static
// is per launcher instantiation, hence per kernel).
#define IGV_SMEM_OPTIN(kernel, bytes)                                                                 \
  do {                                                                                                \
    static std::atomic<unsigned long long> done_{0ull};                                               \
    int dev_ = 0;                                                                                     \
    cudaGetDevice(&dev_); /* the entry point's IgvDeviceGuard made the handle's device current */     \
    const unsigned long long bit_ = 1ull << (dev_ & 63);                                              \
    if (!(done_.load(std::memory_order_acquire) & bit_)) {                                            \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));        \
      done_.fetch_or(bit_, std::memory_order_release);                                                \
    }                                                                                                 \
  } while (0)

// One captured frame of igv_frame_step (igv_frame.cu): the key it was captured for, the instantiated graph and the
// host-side bookkeeping the frame leaves behind.
struct IgvFrameGraph {
  std::vector<unsigned long long> key;
  void* exec = nullptr;                  // cudaGraphExec_t
  std::vector<IgvVar> vars_after;
  int N_after = 0, cur_after = 0, xcur_after = 0, last_visual_path = -1;
  std::vector<int> col_of_slot_after;
  long long launch_delta = 0, uses = 0;
};

struct igv_batch {
  igv_config cfg{};
  IgvKnobs knobs;
  unsigned long long cfg_version = 0;    // bumped by every setter whose values kernels receive by value
  std::vector<IgvFrameGraph> frame_graphs;
  long long graph_replays = 0;
  bool capturing = false;                // igv_frame_step is capturing the stream: no staging-ring events
  int B = 0, ld = 0, xsize = 0, max_rows = 0, qmax = 0, ncols_max = 0, rho = 2;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int ptr_mode = IGV_PTR_HOST;
  std::string err;
  long long launches = 0;
  std::vector<IgvVar> vars;       // insertion order == covariance order (State::_err_variables)
  int N = 0;
  IgvDevParams params{};
  // device memory
  double* P[2] = {nullptr, nullptr};   // ping-pong covariance, B x ld x ld col-major
  int cur = 0;
  double* X[2] = {nullptr, nullptr};   // mean mirror, B x xsize
  int xcur = 0;
  int* flags = nullptr;                // B
  double* chi2 = nullptr;              // table[d-1]
  int chi2_n = 0;
  double* chi2_095 = nullptr;          // 0.95 quantiles, dof 1..128 (the delayed initialisation's hard-coded gate)
  // workspaces
  double* Hs = nullptr;                // stacked projected blocks  B x F x qmax x (ncols_max+1) row-major
  int* f_rows = nullptr;               // B x F rows written per feature (0 = rejected)
  double* f_gamma = nullptr;           // B x F
  int* n_acc = nullptr;                // B
  double* Hc = nullptr;                // compressed [R | Q^T r]   B x ncols_max x (ncols_max+1) row-major
  double* Rpart = nullptr;             // partial triangles for split QR
  int qr_split_cap = 0;
  double* Gws = nullptr;               // partial Gram matrices of the stack  B x qr_split_cap x gram_n1p^2 (k_gram.cu)
  int gram_n1p = 0;
  int compress = 0;                    // IGV_COMPRESS_*
  int stack_f32 = 0;                   // IGV_PREC_FP32_STACK / IGV_PREC_TF32_GRAM: Hs holds floats
  int gram_tc = 0;                     // IGV_PREC_TF32_GRAM: Gram matrix of the float stack on tcgen05 (k_gram_tc.cuh)
  int last_gram_tc = 0;                // the last compression really ran k_gram_tc
  int last_visual_path = -1;           // igv_last_visual_path
  bool feat_fused = false;             // the last k_msckf_features launch accumulated the Gram matrix itself
  double* Zws = nullptr;               // B x max_rows x (ld+1)
  double* Sws = nullptr;               // B x max_rows x max_rows
  double* dxws = nullptr;              // B x ld
  double* Hg = nullptr;                // GNSS rows  B x (2*max_sats) x 16 col-major
  double* rg = nullptr;                // B x 2*max_sats
  double* Rg = nullptr;                // B x 2*max_sats
  int* cnt_g = nullptr;                // B accepted GNSS rows
  double* gam_ws = nullptr;            // B
  double* Dws = nullptr;               // delayed-init workspace: B x dws_stride
  size_t dws_stride = 0;
  int dws_rows = 128, dws_cols = 16;   // most rows / measured columns of a delayed initialisation
  double* Lws = nullptr;               // landmark-update rows: B x (2 L) x (lm_cols + 2)
  double* pre_ws = nullptr;            // per (sequence, IMU step) Phi (225) + G*sigma (180), row-major
  size_t pre_cap = 0;
  std::vector<double> chi2_host;
  // profiling (igv_profile_*)
  bool prof_on = false;
  struct ProfEv { int kind; cudaEvent_t e0, e1; };
  std::vector<ProfEv> prof_events;
  double prof_ms[IGV_K_COUNT] = {0};
  long long prof_cnt[IGV_K_COUNT] = {0};
  // host->device staging (HOST pointer mode): a ring of arenas, one per API call. Bulk arguments travel on a copy
  // stream, so the copies of the NEXT calls overlap the kernels of the current ones; a slot is reused only after
  // the kernels that read it have finished (event `consumed`).
  static constexpr int kSlots = 12;
  struct Slot { char* mem = nullptr; size_t cap = 0, off = 0; cudaEvent_t consumed = nullptr; bool used = false; };
  Slot slots[kSlots];
  size_t slot_cap = 0;                  // common target capacity of the slots
  int slot = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copied = nullptr;
  bool copies_pending = false;          // staged copies the compute stream has not been ordered after yet
  cudaEvent_t fences[4] = {nullptr, nullptr, nullptr, nullptr};   // igv_fence_record / igv_fence_wait
  std::vector<char*> retired;
  IgvTrackTable trk;                    // igv_tracks_* (empty until igv_tracks_create)

  IgvLayout layout() const;
  double* Pc() const { return P[cur]; }
  double* Xc() const { return X[xcur]; }
};

// Every C-ABI entry point that touches the device runs under this guard: the handle's device becomes current for
// the call (allocations, events and launches all use the CURRENT device) and the caller's device is restored.
struct IgvDeviceGuard {
  int prev = -1; bool switched = false;
  explicit IgvDeviceGuard(const igv_batch* h) {
    if (!h) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != h->cfg.device) switched = (cudaSetDevice(h->cfg.device) == cudaSuccess);
  }
  ~IgvDeviceGuard() { if (switched) cudaSetDevice(prev); }
  IgvDeviceGuard(const IgvDeviceGuard&) = delete;
  IgvDeviceGuard& operator=(const IgvDeviceGuard&) = delete;
};

// Order the compute stream after every host->device copy staged so far (called before the first kernel that may
// read a staged argument: every launcher does it, through IgvProfScope or directly).
inline void igv_commit_copies(igv_batch* h) {
  if (!h->copies_pending) return;
  cudaEventRecord(h->copied, h->copy_stream);
  cudaStreamWaitEvent(h->stream, h->copied, 0);
  h->copies_pending = false;
}

struct IgvProfScope {
  igv_batch* h; int kind; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on;
  IgvProfScope(igv_batch* h_, int kind_) : h(h_), kind(kind_), on(h_->prof_on) {
    igv_commit_copies(h);
    if (on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
  }
  ~IgvProfScope() {
    if (on) { cudaEventRecord(e1, h->stream); h->prof_events.push_back({kind, e0, e1}); }
  }
};

// ---- kernel launchers (defined in the k_*.cu files) --------------------------------------------
void igv_launch_state_init(igv_batch* h, const double* R, const double* p, const double* v, const double* bg,
                           const double* ba, const double* Rext, const double* pext, const double* diag21);
void igv_launch_cov_copy(igv_batch* h, double* user, int ld_user, bool to_user);
void igv_launch_cov_blocks(igv_batch* h, const IgvBlocks& blk, double* dst);
void igv_launch_add_variable(igv_batch* h, int size, const double* cov_block_dev);
void igv_launch_marginalize(igv_batch* h, int start, int size, int clone_slot, int lm_slot = -1);
void igv_launch_set_gnss_value(igv_batch* h, int gtype, const double* value_dev);
void igv_launch_propagate(igv_batch* h, int n_steps, const double* gyro, const double* accel, const double* dt,
                          const double* Phi, const double* G);
void igv_launch_augment(igv_batch* h, const double* R_i2w, const double* clone_R, const double* clone_p);
void igv_launch_boxplus(igv_batch* h, const double* dx);
void igv_launch_trace(igv_batch* h, double* out);

struct IgvEkfLaunch {
  IgvBlocks blk;
  int rows;
  const double* H; long strideH; int h_ld; int h_rowmajor;
  int h_upper;               // H is upper triangular (zero left of the diagonal): set for the compressed visual rows
  const double* res; long strideRes; int res_inc;
  const double* R; long strideR; int r_kind;
  double r_iso_value;        // used when r_kind == IGV_R_ISO and R == nullptr
  const int* only_if;        // optional per-sequence mask: update only where nonzero
  double* dx_out;            // device, B x N or null
  int gamma_only; double* gamma_out;
  const int* gate_rows;      // optional per-sequence row count for the joint chi^2 gate (GNSS strong reject)
  int apply_boxplus;
};
void igv_launch_ekf(igv_batch* h, const IgvEkfLaunch& a);

struct IgvMsckfLaunch {
  int mode, F, obs_slots, max_valid;
  const double* pf; const int* anchor; const double* obs; const unsigned char* mask; const int* dof;
  const unsigned char* feat_ok;
  double noise;
};
void igv_launch_msckf_features(igv_batch* h, const IgvMsckfLaunch& a);
void igv_launch_qr_compress(igv_batch* h, int F, int max_valid);
int igv_gram_n1p(int ncols_max);
bool igv_gram_supported(int n);
void igv_launch_gram_compress(igv_batch* h, int F, int max_valid, int split);
void igv_launch_gram_factor(igv_batch* h, int nparts);
void igv_launch_triangulate(igv_batch* h, int F, int obs_slots, const double* obs, const unsigned char* mask,
                            const int* anchor, const igv_tri_params& prm, double* pf_out, unsigned char* ok_out);

struct IgvGnssLaunch {
  int S; const double* unit; const double* res_pos; const double* res_vel; const double* sig_psr;
  const double* sig_dopp; const int* sys; const double* R_enu2ecef; int adjust_yof; int chi2_test;
  IgvBlocks blk;             // var_order used for the rows: SE23, YOF, present clock biases, FS
  int col_of_gnss[6];        // column of each GNSS scalar in H (or -1)
};
void igv_launch_gnss_rows(igv_batch* h, const IgvGnssLaunch& a);

struct IgvGnssResLaunch {
  int S;
  const double* sat_pos; const double* sat_vel; const double* sat_clk; const double* obs; const double* obs_std;
  const double* ttx; const int* sys; const double* T; const double* iono;
  double psr_amp, dopp_amp;
  double* unit; double* res_pos; double* res_vel; double* sig_psr; double* sig_dopp; double* azel; double* atmos;
  const double* clock_init;
};
void igv_launch_gnss_residuals(igv_batch* h, const IgvGnssResLaunch& l);
void igv_launch_sat_states(igv_batch* h, int S, const double* eph, const double* t_obs, const double* psr, const int* sys,
                           double* pos, double* vel, double* clk, double* ttx);

// ---- track table (k_tracks.cu) ----
IgvTrkCols igv_trk_cols(const IgvTrackTable& t);
int igv_trk_col_alloc(IgvTrackTable& t);
int igv_trk_col_release(IgvTrackTable& t, int slot);
bool igv_trk_slot_bits(const IgvTrackTable& t, int n, const int* slots, unsigned long long* bits);
void igv_launch_trk_reset(igv_batch* h);
void igv_launch_trk_collect(igv_batch* h, const int* n_meas, int meas_stride, const unsigned long long* ids,
                            const double* uv);
void igv_launch_trk_mark_lost(igv_batch* h);
struct IgvTrkGatherLaunch {
  int rule, n_selected, min_obs, dof_fixed, F, SW;
  unsigned long long sel_cols;   // physical-column bit mask of the selected clones
  int* entry; int* n_sel; int* track_id; double* obs; unsigned char* mask_all; unsigned char* mask_upd;
  int* anchor_slot; int* dof; unsigned char* feat_ok;
};
void igv_launch_trk_gather(igv_batch* h, const IgvTrkGatherLaunch& g);
void igv_launch_trk_commit_tri(igv_batch* h, int F, const int* entry, const double* pf, const unsigned char* ok,
                               unsigned char* feat_ok);
void igv_launch_trk_erase(igv_batch* h, int F, const int* entry);
void igv_launch_trk_clean(igv_batch* h, unsigned long long col_bits, int erase_empty);
void igv_launch_trk_change_anchor(igv_batch* h, unsigned long long old_cols, double min_depth);
void igv_launch_trk_erase_invalid(igv_batch* h, double min_depth);
void igv_launch_trk_dump(igv_batch* h, const igv_track_dump& d);

void igv_launch_delayed_init(igv_batch* h, const IgvBlocks& blk, int rows, int k, const double* Hold, const double* Hnew,
                             const double* res, double noise_iso, const double* noise2_dev, const int* rows_dev,
                             double chi2_mult, int do_chi2, double prior_cov, int* accepted_dev);
struct IgvLmInitLaunch {
  int SW, anchor_slot, rows_max;
  const double* pf; const double* obs; const unsigned char* mask;
  double* Hx; double* Hf; double* res; int* count;
};
void igv_launch_lm_init_rows(igv_batch* h, const IgvLmInitLaunch& l);
struct IgvLmUpdateLaunch {
  const double* uv; const unsigned char* valid; double noise2;
  double* H; int ldh; int ncols; double* res; double* gamma; int* n_acc;
};
void igv_launch_lm_update_rows(igv_batch* h, const IgvLmUpdateLaunch& l);
void igv_launch_lm_anchor_H(igv_batch* h, int lm_slot, double* H);
void igv_launch_set_lm_value(igv_batch* h, int lm_slot, const double* pf_dev);
struct IgvGnssNewRowsLaunch {
  int S, gtype, adjust_yof;
  const double* unit; const double* res_pos; const double* res_vel; const double* sig_psr; const double* sig_dopp;
  const int* sys; const double* R_enu2ecef; const double* R_ecef2enu;
  double* Hx; double* Hf; double* res; double* noise2; int* count;
};
void igv_launch_gnss_new_rows(igv_batch* h, const IgvGnssNewRowsLaunch& g);
void igv_frame_graphs_destroy(igv_batch* h);
void igv_arena_quiesce(igv_batch* h);
void igv_launch_replace_var_linear(igv_batch* h, int tidx, int tsize, const IgvBlocks& blk, const double* H);
