/*
 * ingvio_b200.h -- C-ABI of the B200-native (sm_100a CUDA) invariant-EKF hot path of InGVIO.
 *
 * The reference (ChangwuLiu/InGVIO, /root/reference) has no FFI: its seam is the C++ class API
 *   State{_cov,_err_variables} (private, friend StateManager)   ingvio_estimator/src/State.h:72-136
 *   StateManager static methods                                 ingvio_estimator/src/StateManager.h:33-128
 *   UpdateBase and the *Update classes                          ingvio_estimator/src/Update.h:36-97
 * This header is what a maintainer binds from inside those method bodies (INTEGRATION.md shows the
 * stubs). Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - One opaque handle = a BATCH of B independent filters ("sequences") that share one variable
 *     layout (same variables in the same order) and live on one device. B = 1 is the drop-in case
 *     for the single-filter reference. Calls on one handle must be serialised by the caller
 *     (the reference is single-threaded: ingvio_estimator/src/IngvioNode.cpp:36); handles are
 *     independent.
 *   - All floating point data is FP64 (the reference computes in double everywhere).
 *   - Matrices are COLUMN-MAJOR (Eigen default) unless a parameter says otherwise; 3x3 rotations
 *     in the state mirror are ROW-MAJOR 9-vectors.
 *   - Bulk array arguments carry a leading batch dimension: element [b][...] at ptr + b*stride,
 *     with the per-sequence stride equal to the stated per-sequence size. They are HOST pointers in
 *     IGV_PTR_HOST mode (default; copied H2D on the handle's stream) or DEVICE pointers in
 *     IGV_PTR_DEVICE mode (consumed in place). Small index/shape arguments are always host values.
 *   - Every function returns an igv_status; nothing ever calls exit() (the reference's
 *     std::exit paths, e.g. StateManager.cpp:157-161, become IGV_ERR_STATE).
 *   - Work is enqueued on the handle's CUDA stream; igv_synchronize() or any *_get call waits.
 */
#ifndef INGVIO_B200_H
#define INGVIO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct igv_batch igv_batch;

typedef enum {
  IGV_OK = 0,
  IGV_ERR_INVALID = 1,   /* bad argument / shape                                                */
  IGV_ERR_CUDA = 2,      /* CUDA runtime error (see igv_last_error)                             */
  IGV_ERR_STATE = 3,     /* operation not valid for the current variable layout                 */
  IGV_ERR_CAPACITY = 4   /* would exceed max_dim / max_clones / max_feats / max_sats / max_rows  */
} igv_status;

enum { IGV_PTR_HOST = 0, IGV_PTR_DEVICE = 1 };

/* State::GNSSType (State.h:75) */
enum { IGV_GNSS_GPS = 0, IGV_GNSS_GLO = 1, IGV_GNSS_GAL = 2, IGV_GNSS_BDS = 3, IGV_GNSS_FS = 4, IGV_GNSS_YOF = 5 };

/* measurement-noise argument kinds of igv_ekf_update / igv_chi2_whiten */
enum { IGV_R_ISO = 0,   /* R = sigma^2 I, R points to ONE double per sequence: sigma^2          */
       IGV_R_DIAG = 1,  /* R = diag(r), R points to `rows` doubles per sequence                 */
       IGV_R_FULL = 2   /* dense rows x rows, column-major                                      */ };

/* visual-update flavours (which reference updater the call reproduces) */
enum { IGV_VIS_ALL_OBS = 0,   /* RemoveLostUpdate: every observation inside the window          */
       IGV_VIS_SELECTED = 1   /* KeyframeUpdate / SwMargUpdate: observations at selected clones  */ };

/* how igv_msckf_update compresses the stacked Jacobian to [R | Q^T r] (the reference: Eigen::SPQR,
 * RemoveLostUpdate.cpp:139-155, SwMargUpdate.cpp:161-176, KeyframeUpdate.cpp:557-572) */
enum { IGV_COMPRESS_AUTO = 0,         /* GRAM where supported (<= 215 columns), else HOUSEHOLDER          */
       IGV_COMPRESS_HOUSEHOLDER = 1,  /* blocked Householder QR of the stack (k_qr.cu)                    */
       IGV_COMPRESS_GRAM = 2          /* R from the Cholesky elimination of [H r]^T [H r] (k_gram.cu)     */ };

/* per-sequence status bits (igv_get_flags) */
enum { IGV_FLAG_NEG_DIAG = 1,      /* negative covariance diagonal after an update (StateManager.cpp:413-421) */
       IGV_FLAG_CHOL_FAIL = 2,     /* innovation covariance not positive definite; update skipped            */
       IGV_FLAG_GNSS_REJECTED = 4, /* joint chi^2 "strong reject" fired (GnssUpdate.cpp:286-287)             */
       IGV_FLAG_TRACKS_FULL = 8,   /* track table full: new tracks of a frame were dropped (igv_tracks_collect) */
       IGV_FLAG_GATHER_CUT = 16,   /* more selected tracks than max_feats: the highest ids were left out        */
       IGV_FLAG_WEAK_PIVOT = 32    /* informational: the Gram-form compression kept a column whose pivot is below 1e-11 of
                                      the column's squared norm, within two decades of the 1e-13 threshold under which
                                      it is treated as dependent, so that row of [R | Q^T r] carries about five digits.
                                      The posterior's backward error is the same eps ||H||^2 as with a Householder R
                                      (tests/test_gpu_illcond.py); IGV_COMPRESS_HOUSEHOLDER avoids the Gram form        */ };

typedef struct {
  int batch;        /* B >= 1                                                                    */
  int max_dim;      /* largest state dimension N (leading dimension of the covariance)           */
  int max_clones;   /* sliding-window capacity                                                   */
  int max_feats;    /* most tracks per visual update                                             */
  int max_sats;     /* most satellites per GNSS epoch                                            */
  int stereo;       /* 0: 2 rows per observation, 1: 4 rows (left + right camera)                */
  int device;       /* CUDA device ordinal                                                       */
  void* stream;     /* cudaStream_t to enqueue on, or NULL for a stream owned by the handle      */
  int max_landmarks;/* SLAM landmarks kept in the state (StateParams::_max_landmarks), 0..32; 0 = none;
                       max_dim must leave room for 3 per landmark                                  */
} igv_config;

/* StateParams (State.h:36-70) + gravity (ImuPropagator.h) */
typedef struct {
  double noise_g, noise_a, noise_bg, noise_ba;   /* State.cpp:39-42                              */
  double noise_clockbias, noise_cb_rw;           /* as StateParams holds them AFTER State.cpp:51-52 */
  double gravity[3];                             /* world gravity vector, e.g. {0,0,-9.8}        */
  double T_cl2cr_R[9];                           /* row-major; left->right camera (State.cpp:33)  */
  double T_cl2cr_p[3];
} igv_params;

/* ---- lifecycle ------------------------------------------------------------------------------ */
igv_status igv_create(const igv_config* cfg, igv_batch** out);
igv_status igv_destroy(igv_batch* h);
const char* igv_last_error(const igv_batch* h);
igv_status igv_set_pointer_mode(igv_batch* h, int mode);
/* Lifetime of bulk arguments in HOST pointer mode: they are copied to the device ASYNCHRONOUSLY on the library's copy
 * stream. Calls with an output argument return after the stream has been synchronised; calls without one
 * (igv_propagate_imu, igv_tracks_collect, igv_box_plus, igv_msckf_update without outputs, ...) may return while the copy
 * from a PAGE-LOCKED buffer is still in flight: such a buffer must not be modified until igv_synchronize (or an
 * igv_fence_record / igv_fence_wait pair) has returned. Pageable buffers are staged by the driver before the copy call
 * returns and may be reused at once. */
igv_status igv_synchronize(igv_batch* h);
igv_status igv_set_compression(igv_batch* h, int kind);   /* IGV_COMPRESS_* (default AUTO) */
/* Arithmetic mode of the visual update (BASELINE configs[4], "FP32 vs FP64").
 *   IGV_PREC_FP64       everything in double (the parity path, default);
 *   IGV_PREC_FP32_STACK the projected per-track blocks [H | r] are STORED in single precision (half the HBM traffic of
 *                       the stack: the bound of wide windows / stereo / small batches, where the stack is materialised);
 *                       Jacobians, null-space projection and the chi^2 gate are still evaluated in double (no gate
 *                       decision can flip), the Gram matrix is accumulated in double and the whole EKF update is
 *                       double. Measured tolerance: tests/test_gpu_precision.py, profiles/r02_precision_sweep.md. */
/*   IGV_PREC_TF32_GRAM  as IGV_PREC_FP32_STACK, and the Gram matrix of the float stack is formed on the 5th-generation
 *                       tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM): every float is split into two TF32
 *                       terms and hi^T hi + hi^T lo + lo^T hi is accumulated in FP32 over 128 rows at a time, the
 *                       128-row sums in FP64; factorisation and EKF update stay FP64. Used for stacks of 129..192
 *                       columns (windows of 22..31 clones); narrower and wider ones take IGV_PREC_FP32_STACK's kernel.
 *                       A throughput / accuracy trade-off for wide windows (c5: 1.74 x the FP64 path), NOT a parity mode:
 *                       the unit's truncating FP32 accumulation leaves ~1.5e-6 relative in the Gram matrix, which at
 *                       c5 moves the posterior by up to 2 cm / 5e-4 relative in P within 40 frames
 *                       (tests/test_gpu_precision.py, profiles/r02_tcgen05_eval.md). */
enum { IGV_PREC_FP64 = 0, IGV_PREC_FP32_STACK = 1, IGV_PREC_TF32_GRAM = 2 };
igv_status igv_set_precision(igv_batch* h, int mode);
int igv_last_gram_tensor(const igv_batch* h);   /* 1 if the last visual update's Gram matrix came from the tcgen05 kernel */
/* which kernels the last igv_msckf_update used: 0 Householder QR of the materialised stack, 1 Gram matrix of the
 * materialised stack, 2 Gram matrix accumulated inside the per-track kernel (no stack in HBM); -1 before any update */
int igv_last_visual_path(const igv_batch* h);
long long igv_launch_count(const igv_batch* h);          /* kernels launched so far on this handle */
igv_status igv_set_params(igv_batch* h, const igv_params* p);
/* chi^2 quantile table: table[d-1] = quantile(d), d = 1..max_dof.  Replaces
 * UpdateBase::setChiSquaredTable (Update.cpp:27-34, boost::math::quantile). */
igv_status igv_set_chi2_table(igv_batch* h, const double* table, int max_dof);
/* quantile(chi_squared(dof), p) without Boost (Update.cpp:31-32); igv_add_variable_delayed uses it with the
 * reference's hard-coded p = 0.95 (StateManager.cpp:613-615), independent of the table above. NaN on bad input. */
double igv_chi2_quantile(double p, int dof);

/* ---- state: variables and mean --------------------------------------------------------------
 * State::State + State::initStateAndCov (State.cpp:60-91,126-167): variables SE23@0, bg@9, ba@12,
 * cam-IMU extrinsics@15 (N = 21); covariance = diag(cov_diag[21]).
 * R_* are row-major 3x3; every array is per sequence (B x ...), cov_diag is shared (21). */
igv_status igv_state_init(igv_batch* h, const double* R_i2w, const double* p, const double* v,
                          const double* bg, const double* ba, const double* R_ext, const double* p_ext,
                          const double* cov_diag21);
int igv_dim(const igv_batch* h);          /* State::curr_cov_size()            State.h:108 */
int igv_num_variables(const igv_batch* h);/* State::curr_err_variable_size()   State.h:109 */
int igv_num_clones(const igv_batch* h);
int igv_clone_idx(const igv_batch* h, int slot);   /* Type::idx() of the slot-th oldest clone, -1 if none */
int igv_gnss_idx(const igv_batch* h, int gtype);   /* Type::idx() of a GNSS scalar, -1 if absent          */
/* Packed mean per sequence, doubles:
 *  [0:9) R_i2w  [9:12) p  [12:15) v  [15:18) bg  [18:21) ba  [21:30) R_ext  [30:33) p_ext
 *  [33:39) gnss values (GPS,GLO,GAL,BDS,FS,YOF; 0 if absent)
 *  [39 + 12*s ...) clone s: R_c2w (9), p (3)        -> igv_state_size() doubles in total          */
int igv_state_size(const igv_batch* h);
igv_status igv_state_get(igv_batch* h, double* dst);
igv_status igv_state_set(igv_batch* h, const double* src);
/* Pipelined read-back: the *_async calls enqueue the copy and return; igv_fence_record marks a point of the handle's
 * stream (fence 0..3) and igv_fence_wait blocks the host until the device has reached it. With HOST pointers the
 * destination must be page-locked for the copy to overlap. A caller that reads the result of frame k after
 * submitting frame k+1 lets the bulk host->device copies of k+1 overlap the kernels of k. */
igv_status igv_state_get_async(igv_batch* h, double* dst);
igv_status igv_cov_trace_async(igv_batch* h, double* trace_out /* B */);
igv_status igv_fence_record(igv_batch* h, int fence);
igv_status igv_fence_wait(igv_batch* h, int fence);

/* ---- covariance lifecycle (StateManager.cpp:121-242) ---------------------------------------- */
igv_status igv_cov_get(igv_batch* h, double* dst, int ld);         /* getFullCov; B x (ld*N)       */
igv_status igv_cov_set(igv_batch* h, const double* src, int ld);   /* also checkpoint restore      */
/* getMarginalCov (StateManager.cpp:128-153): dst is B x (n x n), n = sum(size) */
igv_status igv_cov_get_blocks(igv_batch* h, int n_blocks, const int* idx, const int* size, double* dst);
/* addGNSSVariable / margGNSSVariable (StateManager.cpp:216-242). value: B doubles. */
igv_status igv_add_gnss_variable(igv_batch* h, int gtype, const double* value, double cov);
igv_status igv_marg_gnss_variable(igv_batch* h, int gtype);
/* addVariableIndependent (StateManager.cpp:194-214) for an opaque variable (no mean mirror);
 * cov_block is size x size column-major, shared by all sequences. */
igv_status igv_add_variable_independent(igv_batch* h, int size, const double* cov_block);
/* marginalize (StateManager.cpp:155-192) of the variable that starts at `idx`. */
igv_status igv_marginalize(igv_batch* h, int idx);
/* margSlidingWindowPose (StateManager.cpp:316-338): slot 0 is the oldest clone. */
igv_status igv_marginalize_clone(igv_batch* h, int slot);

/* ---- propagation ---------------------------------------------------------------------------- */
/* StateManager::propagateStateCov (StateManager.cpp:42-119). Phi: B x 225 (15x15 col-major),
 * G: B x 180 (15x12 col-major), dt: B. Covariance only (the caller owns the mean). */
igv_status igv_propagate_cov(igv_batch* h, const double* Phi, const double* G, const double* dt);
/* ImuPropagator::propagateUntil loop body x n_steps (ImuPropagator.cpp:98-162 analytic branch +
 * :260-271 + StateManager.cpp:42-119), on the device mean mirror: gyro/accel: B x n_steps x 3 (raw,
 * biases are subtracted on device), dt: B x n_steps (steps with dt < 1e-6 are skipped, :262). */
igv_status igv_propagate_imu(igv_batch* h, int n_steps, const double* gyro, const double* accel,
                             const double* dt);
/* StateManager::augmentSlidingWindowPose (StateManager.cpp:253-296) from the device mean mirror. */
igv_status igv_augment_clone(igv_batch* h);
/* Same, covariance only, rotation supplied by the caller (J depends on R_i2w only, :279-282).
 * R_i2w: B x 9 row-major. The clone mean in the mirror is set from clone_R/clone_p if non-NULL. */
igv_status igv_augment_clone_cov(igv_batch* h, const double* R_i2w, const double* clone_R,
                                 const double* clone_p);

/* ---- EKF update (StateManager.cpp:359-426) -------------------------------------------------- */
/* var_order is given as (idx,size) blocks; H: B x (ldh*n) col-major with n = sum(size);
 * res: B x rows; R by r_kind. Applies the retraction to the device mean mirror (boxPlus,
 * StateManager.cpp:244-251) and, if dx_out != NULL, returns dx (B x N) for the caller's Type objects. */
igv_status igv_ekf_update(igv_batch* h, int n_blocks, const int* blk_idx, const int* blk_size, int rows,
                          const double* H, int ldh, const double* res, const double* R, int r_kind,
                          double* dx_out);
/* UpdateBase::whitenResidual (Update.cpp:36-79): gamma[b] = res^T (H P_s H^T + R)^-1 res. */
igv_status igv_chi2_whiten(igv_batch* h, int n_blocks, const int* blk_idx, const int* blk_size, int rows,
                           const double* H, int ldh, const double* res, const double* R, int r_kind,
                           double* gamma_out);
/* StateManager::boxPlus (StateManager.cpp:244-251) on the device mean mirror; dx: B x N. */
igv_status igv_box_plus(igv_batch* h, const double* dx);

/* ---- fused MSCKF visual update --------------------------------------------------------------
 * RemoveLostUpdate::updateState{Mono,Stereo} (RemoveLostUpdate.cpp:40-167, :276-405),
 * KeyframeUpdate::updateState* (KeyframeUpdate.cpp:438-735), SwMargUpdate::updateState*
 * (SwMargUpdate.cpp:42-365) after track selection and triangulation: per-feature residual /
 * Jacobian over clone poses, left null-space projection, chi^2 gate, stacking, QR compression,
 * ekfUpdate, boxPlus. Clone poses are read from the device mean mirror. */
typedef struct {
  int mode;                  /* IGV_VIS_ALL_OBS | IGV_VIS_SELECTED                                */
  int n_feats;               /* F <= max_feats                                                    */
  const double* pf_w;        /* B x F x 3      triangulated landmark, world frame                 */
  const int* anchor_slot;    /* B x F          clone slot of the landmark's anchor pose           */
  const double* obs;         /* B x F x SW x rho  normalised image coordinates per clone slot     */
  const unsigned char* obs_mask; /* B x F x SW  1 if the track has an observation at that slot and
                                    (SELECTED mode) the slot is one of the selected clones        */
  const int* chi2_dof;       /* B x F  dof of the gate: #obs-1 (RemoveLostUpdate.cpp:95-96),
                                #selected-1 (SwMargUpdate.cpp:129-130), 2 (KeyframeUpdate.cpp:525-526) */
  int obs_slots;             /* SW: slot dimension of obs / obs_mask (>= current clone count)      */
  double noise;              /* visual_noise (sigma, normalised units)                             */
  int max_valid;             /* stop after this many accepted tracks (RemoveLostUpdate.h:38); <=0: no cap */
  double* dx_out;            /* optional B x N                                                    */
  int* n_accepted_out;       /* optional B                                                        */
  double* gamma_out;         /* optional B x F chi^2 statistics (NaN for tracks with < 2 usable obs) */
  const unsigned char* feat_ok; /* optional B x F: 0 skips the track (e.g. ok_out of igv_triangulate)   */
} igv_msckf_args;
igv_status igv_msckf_update(igv_batch* h, const igv_msckf_args* a);

/* ---- triangulation ("next" row) ----------------------------------------
 * Triangulator::triangulateMonoObs / triangulateStereoObs (Triangulator.cpp:173-359) for every track
 * of every sequence, over the clone poses of the device mean mirror, followed by the anchor-depth
 * check of FeatureInfoManager::triangulateFeatureInfo{Mono,Stereo} (MapServerManager.cpp:275-341).
 * pf_out / ok_out can be fed to igv_msckf_update (pf_w / feat_ok) without leaving the device. */
typedef struct {
  double trans_thres, huber_epsilon, conv_precision, init_damping;   /* Triangulator.h:38-41 */
  int outer_loop_max_iter, inner_loop_max_iter;                      /* :42-43 */
  double max_depth, min_depth;                                       /* :45-46 */
} igv_tri_params;
typedef struct {
  int n_feats;                   /* F <= max_feats                                                    */
  const double* obs;             /* B x F x SW x rho                                                  */
  const unsigned char* obs_mask; /* B x F x SW                                                        */
  int obs_slots;                 /* SW                                                                */
  const int* anchor_slot;        /* optional B x F: landmark must lie in front of this clone's camera */
  igv_tri_params prm;
  double* pf_out;                /* B x F x 3 world position (zeros when not ok)                      */
  unsigned char* ok_out;         /* B x F                                                             */
} igv_tri_args;
igv_status igv_triangulate(igv_batch* h, const igv_tri_args* a);

/* ---- fused GNSS update ------------------------------------------------------------------------
 * GnssUpdate::updateTrackedSys (GnssUpdate.cpp:84-293) from the psr_res / dopp_res output
 * boundary (gnss_comm/src/gnss_spp.cpp:99-146, :256-282). */
typedef struct {
  int n_sats;                /* S <= max_sats                                                     */
  const double* unit;        /* B x S x 3  receiver->satellite unit vectors (= -J[:,0:3]), ECEF    */
  const double* res_pos;     /* B x S      psr_res output                                          */
  const double* res_vel;     /* B x S      dopp_res output                                         */
  const double* sigma_psr;   /* B x S      psr_noise of GnssUpdate.cpp:187                         */
  const double* sigma_dopp;  /* B x S      dopp_noise of GnssUpdate.cpp:256                        */
  const int* sys;            /* B x S      IGV_GNSS_GPS..BDS of each satellite                     */
  const double* R_enu2ecef;  /* B x 9 row-major (GvioAligner::getRenu2ecef)                        */
  int is_adjust_yof;         /* GnssUpdate.cpp:164-167,239-242                                     */
  int chi2_test;             /* per-row gate (gnss_chi2_test, :190,:259)                           */
  int strong_reject;         /* joint gate if rows <= 14 (:286)                                    */
  double* dx_out;            /* optional B x N                                                    */
} igv_gnss_args;
igv_status igv_gnss_update(igv_batch* h, const igv_gnss_args* a);

/* GnssUpdate::addNewTrackedSys (GnssUpdate.cpp:317-476) for ONE newly seen system -- the reference loops over
 * sys_to_add and ends every iteration in StateManager::addVariableDelayed (StateManager.cpp:547-630). The measurement
 * rows are built on the device from the same psr_res / dopp_res boundary as igv_gnss_update, evaluated by the caller
 * with the SPP initial value of the new scalar in the receiver clock vector (GnssUpdate.cpp:351-370):
 *   clock bias g (IGV_GNSS_GPS..BDS): the satellites with sys == g (getResJacobianOfSys, :478-521),
 *        H_x = [u^T R_w2e [p]x, -u^T R_w2e, 0 | yof], res = -res_pos, noise = psr_amp sqrt(mean ura psr_std / sin^2 el);
 *   clock drift (IGV_GNSS_FS): every satellite, H_x = [u^T R_w2e [v]x, 0, -u^T R_w2e | yof], res = -res_vel (:374-425);
 *   H_f = ones; chi^2 gate at the reference's 0.95 quantile times chi2_mult (0.95 at both call sites).
 * The yaw-offset column uses R_ecef2enu * dR_z (getRecef2enu(), :403,:464) as the reference does.  Sequences may hold
 * different numbers of satellites of the system; one with fewer than two is not initialised (StateManager.cpp:574-578).
 * Rejected sequences keep a decoupled scalar with prior_cov_if_rejected, like igv_add_variable_delayed. */
typedef struct {
  int n_sats;                /* S, 2..max_sats                                                     */
  int gtype;                 /* IGV_GNSS_GPS..IGV_GNSS_BDS or IGV_GNSS_FS: the scalar to add        */
  const double* value;       /* B          SPP initial value (posSpp(3+i) / velSpp(3), :414,:467)   */
  const double* unit;        /* B x S x 3  as igv_gnss_args                                         */
  const double* res_pos;     /* B x S                                                               */
  const double* res_vel;     /* B x S                                                               */
  const double* sigma_psr;   /* B x S      psr_amp sqrt(ura psr_std / sin^2 el)                     */
  const double* sigma_dopp;  /* B x S      dopp_amp sqrt(ura dopp_std c / f / sin^2 el)             */
  const int* sys;            /* B x S                                                               */
  const double* R_enu2ecef;  /* B x 9 row-major                                                     */
  const double* R_ecef2enu;  /* B x 9 row-major, or NULL = transpose of R_enu2ecef                  */
  int is_adjust_yof;
  double chi2_mult;          /* 0 = the reference's 0.95                                            */
  double prior_cov_if_rejected;
  int* accepted_out;         /* optional B                                                          */
  double* dx_out;            /* optional B x (N+1)                                                  */
} igv_gnss_new_sys_args;
igv_status igv_gnss_add_new_tracked_sys(igv_batch* h, const igv_gnss_new_sys_args* a);

/* ---- GNSS residual generator ("next" row) ------------------------------------------
 * gnss_comm::psr_res (gnss_comm/src/gnss_spp.cpp:99-146) and gnss_comm::dopp_res (:256-282) with sat_azel,
 * ecef2geo, calculate_trop_delay (Saastamoinen + Niell) and calculate_ion_delay (Klobuchar) of
 * gnss_comm/src/gnss_utility.cpp:347-387, :762-901, evaluated at the receiver state GnssUpdate::updateTrackedSys
 * assembles from the filter mean (GnssUpdate.cpp:101-111), plus the measurement sigmas of GnssUpdate.cpp:177-186 and
 * :246-255. Inputs are the outputs of gnss_comm::sat_states (gnss_spp.cpp:50-97) and the raw L1 observations; the
 * outputs are the inputs of igv_gnss_update (in DEVICE pointer mode the two calls chain without a host round trip). */
typedef struct {
  int n_sats;                /* S <= max_sats                                                                  */
  const double* sat_pos;     /* B x S x 3  SatState::pos, ECEF at the transmit time [m]                         */
  const double* sat_vel;     /* B x S x 3  SatState::vel [m/s]                                                  */
  const double* sat_clk;     /* B x S x 3  SatState::dt [s], ddt [s/s], tgd [s]                                 */
  const double* obs;         /* B x S x 3  L1 pseudo-range [m], Doppler [Hz], carrier frequency [Hz] (<= 0: the
                                satellite has no L1 observation and its outputs stay zero, gnss_spp.cpp:117-119) */
  const double* obs_std;     /* B x S x 3  ephemeris ura, Obs::psr_std, Obs::dopp_std of the L1 observation     */
  const double* ttx;         /* B x S x 2  transmit time as day of year (time2doy) and GPS seconds of week
                                (time2gpst), the only uses gnss_comm makes of it here                           */
  const int* sys;            /* B x S      IGV_GNSS_GPS..BDS (= gnss_comm::sys2idx)                              */
  const double* T_enu2ecef;  /* B x 12     GvioAligner::getTenu2ecef: rotation (9, row-major) + translation (3) */
  const double* iono;        /* B x 8      Klobuchar parameters, or NULL (no ionospheric delay)                 */
  double psr_noise_amp, dopp_noise_amp;   /* _psr_noise_amp, _dopp_noise_amp (GnssUpdate.h)                      */
  double* unit;              /* out B x S x 3  receiver->satellite unit vectors (= -J[:, 0:3])                   */
  double* res_pos;           /* out B x S                                                                       */
  double* res_vel;           /* out B x S                                                                       */
  double* sigma_psr;         /* out B x S                                                                       */
  double* sigma_dopp;        /* out B x S                                                                       */
  double* azel;              /* out B x S x 2, optional                                                         */
  double* atmos;             /* out B x S x 2 (ionosphere, troposphere delay [m]), optional                     */
  const double* clock_init;  /* optional B x 5 (GPS, GLO, GAL, BDS clock bias [m], clock drift): receiver clock entries
                                that override the state's (NaN = keep the state's), the SPP initial values
                                addNewTrackedSys writes into xyzt / dopp (GnssUpdate.cpp:351-370); NULL = none     */
} igv_gnss_res_args;
igv_status igv_gnss_residuals(igv_batch* h, const igv_gnss_res_args* a);

/* ---- ephemeris -> satellite state -------------------------------------------------------------------------------
 * gnss_comm::sat_states (gnss_comm/src/gnss_spp.cpp:50-97): signal transmit time from the pseudo-range, satellite
 * clock (eph2svdt / geph2svdt), position and velocity (eph2pos, eph2vel: Kepler orbit with harmonic corrections, BDS
 * GEO rotation; geph2pos / geph2vel: RK4 of the GLONASS force model), gnss_comm/src/gnss_utility.cpp:390-733.
 * Times are seconds relative to the ephemeris epoch toe (time_diff(t, toe)); calendar types stay on the host.
 * One ephemeris record of IGV_EPH_STRIDE doubles per satellite:
 *   Kepler (GPS, GAL, BDS): A, e, i0, OMG0, omg, M0, delta_n, OMG_dot, i_dot, cuc, cus, crc, crs, cic, cis, af0, af1,
 *     af2, toe_tow (BDS: time2bdt(toe - 14 s)), tgd[0], time_diff(toe, toc), prn, 0, 0
 *   GLONASS: pos[3], vel[3], acc[3], tau_n, gamma, 0...
 * Outputs have the layout igv_gnss_residuals consumes (sat_pos, sat_vel, sat_clk = dt, ddt, tgd); ttx_rel (optional,
 * B x S) is time_diff(transmit time, toe), from which the caller derives day-of-year / seconds-of-week. */
#define IGV_EPH_STRIDE 24
typedef struct {
  int n_sats;
  const double* eph;         /* B x S x IGV_EPH_STRIDE                                                           */
  const double* t_obs_rel;   /* B x S  time_diff(obs->time, toe)                                                 */
  const double* psr;         /* B x S  L1 pseudo-range [m]; <= 0: no L1 observation, the state stays zero        */
  const int* sys;            /* B x S  IGV_GNSS_GPS..BDS                                                         */
  double* sat_pos;           /* out B x S x 3 */
  double* sat_vel;           /* out B x S x 3 */
  double* sat_clk;           /* out B x S x 3 */
  double* ttx_rel;           /* out B x S, optional */
} igv_sat_state_args;
igv_status igv_sat_states(igv_batch* h, const igv_sat_state_args* a);

/* ---- track table ("next" row) ------------------------------------------------------
 * The MapServer (MapServer.h:69-134: std::map<int, FeatureInfo> with per-feature observation maps keyed by clone
 * timestamp) as a device-resident table per sequence, fed from the tracker's wire format
 * (feature_tracker/msg/MonoMeas.msg, StereoMeas.msg: uint64 id, float64 u0 v0 [u1 v1]) and emitting the dense
 * per-track arrays igv_triangulate / igv_msckf_update consume, so a whole frame stays on the device.
 * A track holds: id (the message id narrowed to `int` like MapServer.cpp:24), one observation per clone of the
 * window, the anchor clone (AnchoredLandmark::getAnchoredPose), _isToMarg, _isTri and the landmark value / FEJ
 * position. Observations are stored per PHYSICAL clone column; the handle maps window slots to columns and follows
 * igv_augment_clone* / igv_marginalize_clone, so nothing moves when the window slides. Only MSCKF-type features are
 * modelled (SLAM landmarks are section 8f rank 3). Every call is one launch over all B sequences; selections come out
 * in ascending id order, the iteration order of the reference's std::map.                                          */
igv_status igv_tracks_create(igv_batch* h, int max_tracks /* per sequence, <= 4096 */);
igv_status igv_tracks_reset(igv_batch* h);                /* map_server.reset(new MapServer()) */
/* MapServerManager::collectMonoMeas / collectStereoMeas (MapServerManager.cpp:189-219) with
 * FeatureInfoManager::collect*Meas (:101-187) for one frame message per sequence, at the NEWEST clone (the reference
 * asserts state->_timestamp is in the window, :107-111). n_meas: B counts (<= meas_stride <= 4096); ids:
 * B x meas_stride; uv: B x meas_stride x rho. Message order is kept: a repeated id inside one message keeps its first
 * measurement (":126-130 skip adding"); new tracks take free table entries in message order. */
igv_status igv_tracks_collect(igv_batch* h, const int* n_meas, int meas_stride, const unsigned long long* ids,
                              const double* uv);
/* MapServerManager::markMargMonoFeatures / markMargStereoFeatures (:221-273): tracks without an observation at the
 * newest clone get _isToMarg. */
igv_status igv_tracks_mark_lost(igv_batch* h);

enum { IGV_TRK_LOST = 0,      /* RemoveLostUpdate.cpp:45-59 / :280-294: MSCKF tracks with _isToMarg               */
       IGV_TRK_SEEN_AT = 1    /* SwMargUpdate.cpp:61-85, KeyframeUpdate.cpp:455-480: observed at every selected clone */ };
typedef struct {
  int rule;                      /* IGV_TRK_LOST | IGV_TRK_SEEN_AT                                                 */
  int n_selected;                /* SEEN_AT: number of selected clones                                             */
  const int* selected_slots;     /* SEEN_AT: their window slots (HOST array)                                       */
  int min_obs;                   /* LOST: feat_ok needs this many observations (4 mono :51, 3 stereo :287)          */
  int dof_fixed;                 /* SEEN_AT: > 0 fixed gate dof (KeyframeUpdate.cpp:525-526: 2); else n_selected-1  */
  int n_feats;                   /* F <= max_feats: capacity of the outputs                                         */
  int obs_slots;                 /* SW >= clone count: slot dimension of the outputs                                */
  int* track_entry;              /* B x F      table entry of the f-th selected track, -1 beyond n_sel              */
  int* n_sel;                    /* B          selected tracks (capped at F, IGV_FLAG_GATHER_CUT)                   */
  int* track_id;                 /* optional B x F ids (0 beyond n_sel)                                             */
  double* obs;                   /* B x F x SW x rho  -> igv_triangulate.obs / igv_msckf_update.obs                 */
  unsigned char* mask_all;       /* B x F x SW  every observation of the track -> igv_triangulate.obs_mask          */
  unsigned char* mask_upd;       /* B x F x SW  LOST: = mask_all; SEEN_AT: observations at the selected clones
                                                -> igv_msckf_update.obs_mask                                       */
  int* anchor_slot;              /* B x F      window slot of the anchor clone (0 beyond n_sel)                     */
  int* chi2_dof;                 /* B x F      LOST: #obs-1; SEEN_AT: n_selected-1 or dof_fixed                     */
  unsigned char* feat_ok;        /* B x F      LOST: #obs >= min_obs; SEEN_AT: 1; 0 beyond n_sel                    */
} igv_track_gather_args;
igv_status igv_tracks_gather(igv_batch* h, const igv_track_gather_args* a);
/* Second half of FeatureInfoManager::triangulateFeatureInfo{Mono,Stereo} (MapServerManager.cpp:284-306): where ok,
 * the landmark value (and, at the first success, the FEJ value) is stored and _isTri set; feat_ok (optional, in/out)
 * is and-ed with ok so that it can go straight into igv_msckf_update. pf / ok: outputs of igv_triangulate.
 * (_numOfTri, read nowhere in the reference, is not kept.) */
igv_status igv_tracks_commit_tri(igv_batch* h, int n_feats, const int* track_entry, const double* pf,
                                 const unsigned char* ok, unsigned char* feat_ok);
/* map_server->erase(id) for the gathered tracks (RemoveLostUpdate.cpp:58-59,165-166). */
igv_status igv_tracks_erase(igv_batch* h, int n_feats, const int* track_entry);
/* SwMargUpdate::clean{Mono,Stereo}ObsAtMargTime (SwMargUpdate.cpp:191-213, :389-411), KeyframeUpdate twin
 * (KeyframeUpdate.cpp:251-278): drop the observations at these window slots, erase tracks left without any. */
igv_status igv_tracks_clean_obs(igv_batch* h, int n_slots, const int* clone_slots /* HOST */);
/* SwMargUpdate::changeMSCKFAnchor (SwMargUpdate.cpp:216-259; min_depth 0) / KeyframeUpdate::changeMSCKFAnchor
 * (KeyframeUpdate.cpp:280-327; min_depth 0.3): tracks anchored at one of old_slots move to the newest clone if
 * triangulated and deeper than min_depth in it, else they are erased. */
igv_status igv_tracks_change_anchor(igv_batch* h, int n_old, const int* old_slots /* HOST */, double min_depth);
/* MapServerManager::eraseInvalidFeatures (MapServerManager.cpp:454-490): triangulated tracks whose depth in the
 * anchor camera is <= min_depth (reference: 0.2) or whose anchor left the window are erased. */
igv_status igv_tracks_erase_invalid(igv_batch* h, double min_depth);
/* Dump in table-entry order (tests, checkpoints). All outputs optional. used/to_marg/is_tri: B x T; id: B x T;
 * slot_mask: B x T, bit s = observation at window slot s; anchor_slot: B x T (-1: none / left the window);
 * pf, pf_fej: B x T x 3; obs: B x T x obs_slots x rho (zeros where no observation); n_tracks: B. */
typedef struct {
  int obs_slots;
  int* id; unsigned char* used; unsigned char* to_marg; unsigned char* is_tri; unsigned long long* slot_mask;
  int* anchor_slot; double* pf; double* pf_fej; double* obs; int* n_tracks;
} igv_track_dump;
igv_status igv_tracks_get(igv_batch* h, const igv_track_dump* d);
int igv_tracks_capacity(const igv_batch* h);   /* max_tracks, 0 before igv_tracks_create */

/* ---- one whole frame cycle -------------------------------------------------------------------
 * IngvioFilter::callbackMonoFrame / callbackStereoFrame (IngvioFilter.cpp:143-231) behind one call:
 *   propagateAugmentAtEnd (n_imu IMU samples, then the clone) -> visual update (igv_msckf_update arguments, or NULL)
 *   -> marginalisation of the listed clone slots (in the given order; slots refer to the window at that moment)
 *   -> GNSS update (igv_gnss_update arguments, or NULL).
 * DEVICE pointer mode: the kernel sequence is captured into a CUDA graph the second time the same frame shape occurs
 * (variable layout, argument pointers and sizes) and replayed afterwards -- keep the argument buffers at fixed
 * addresses and refill them between frames. dx_out / n_accepted_out / gamma_out of the nested arguments are honoured
 * (device pointers). HOST pointer mode: the calls run one after the other (no graph). IGV_GRAPH=0 disables replay. */
typedef struct {
  int n_imu;                     /* K IMU samples of the frame interval (0: none)                    */
  const double* gyro;            /* B x K x 3                                                         */
  const double* accel;           /* B x K x 3                                                         */
  const double* dt;              /* B x K                                                             */
  const igv_msckf_args* visual;  /* or NULL                                                           */
  int n_marg;                    /* clones leaving the window after the visual update                 */
  const int* marg_slots;         /* n_marg slots (host array)                                         */
  const igv_gnss_args* gnss;     /* or NULL                                                           */
} igv_frame_args;
igv_status igv_frame_step(igv_batch* h, const igv_frame_args* a);
long long igv_graph_replays(const igv_batch* h);   /* frames served by a graph launch so far           */

/* ---- delayed initialisation / linear replacement ------------------------------------------------
 * StateManager::addVariableDelayed (StateManager.cpp:547-630, with :462-541): new 1-dim variable.
 * H_old: B x (rows x n_old) col-major (ld = rows), H_new: B x rows, res: B x rows.
 * accepted_out: B ints (0/1); sequences that fail the chi^2 keep a decoupled variable with the
 * supplied prior_cov_if_rejected so that the batch keeps one layout. gtype >= 0 registers the new
 * scalar as that GNSS state (value: B doubles), gtype < 0 adds an opaque scalar. dx_out (optional,
 * B x (N+1)) receives the correction of the residual EKF for the caller's Type objects. */
igv_status igv_add_variable_delayed(igv_batch* h, int gtype, const double* value, int n_blocks,
                                    const int* blk_idx, const int* blk_size, int rows, const double* H_old,
                                    const double* H_new, const double* res, double noise_iso,
                                    double chi2_mult, int do_chi2, double prior_cov_if_rejected,
                                    int* accepted_out, double* dx_out);
/* ---- SLAM landmarks in the state (mono) ------------------------------------------
 * An anchored landmark is a 3-dim variable holding the world position p_f, tied to an anchor clone
 * (AnchoredLandmark.h:27-108); its retraction is p_f <- Gamma0(dtheta_anchor) p_f + Gamma1(dtheta_anchor) dp
 * (AnchoredLandmark.cpp:227-243), applied by every update once landmarks exist. The packed mean gains 3 doubles per
 * landmark slot after the clones (igv_state_size). Landmark slots are in order of insertion. */
int igv_num_landmarks(const igv_batch* h);
int igv_landmark_idx(const igv_batch* h, int lm_slot);           /* Type::idx(), -1 if absent            */
int igv_landmark_anchor(const igv_batch* h, int lm_slot);        /* anchor clone slot                    */
/* LandmarkUpdate::initNewLandmarkMono, the body of its loop for ONE track (LandmarkUpdate.cpp:395-421 with
 * calcResJacobianSingleFeatAllMonoObs :426-500): rows over every clone of the window that observed the track,
 * H_x on the clones, H_f (rows x 3) on the landmark, then StateManager::addVariableDelayed (Givens split of H_f, 0.95
 * chi^2 gate times chi2_mult on the remaining rows, invertible 3 x 3 initialisation, EKF on the remaining rows).
 * Sequences that fail the gate (or have fewer than two observations) keep a decoupled landmark with
 * prior_cov_if_rejected on its diagonal, so that the batch keeps one layout (B = 1: marginalise it again = the
 * reference's `continue`). */
typedef struct {
  int obs_slots;             /* SW' >= number of clones                                            */
  int anchor_slot;           /* anchor clone slot of the new landmark (shared by the batch)         */
  const double* pf_w;        /* B x 3      triangulated world position                              */
  const double* obs;         /* B x SW' x 2 normalised image coordinates per clone slot             */
  const unsigned char* obs_mask; /* B x SW'                                                         */
  double noise;              /* _noise = visual_noise                                               */
  double chi2_mult;          /* 0 = the reference's 0.95                                            */
  double prior_cov_if_rejected;
  int* accepted_out;         /* optional B                                                          */
} igv_lm_init_args;
igv_status igv_landmark_init(igv_batch* h, const igv_lm_init_args* a);
/* LandmarkUpdate::updateLandmarkMono (LandmarkUpdate.cpp:32-149 with calcResJacobianSingleLandmarkMono :521-572):
 * every landmark of the state observed in the current image contributes 2 rows on [SE23 | cam-IMU extrinsics | anchor
 * clone | landmark], gated by the chi^2 test (dof 2) and stacked into one EKF update with R = noise^2 I. */
typedef struct {
  const double* uv;          /* B x L x 2  current observation of landmark slot l (L = igv_num_landmarks) */
  const unsigned char* valid;/* B x L      0: skip this landmark in this sequence                   */
  double noise;
  int* n_accepted_out;       /* optional B                                                          */
  double* gamma_out;         /* optional B x L gate statistic (NaN where skipped)                   */
} igv_lm_update_args;
igv_status igv_landmark_update(igv_batch* h, const igv_lm_update_args* a);
/* FeatureInfoManager::changeAnchoredPose (MapServerManager.cpp:343-378): the landmark becomes a linear function of
 * (old anchor, new anchor, itself) through StateManager::replaceVarLinear with H = [-[p_f]x 0 | [p_f]x 0 | I]; the world
 * position is unchanged. */
igv_status igv_landmark_change_anchor(igv_batch* h, int lm_slot, int new_clone_slot);
/* StateManager::margAnchoredLandmarkInState (StateManager.cpp:340-353). */
igv_status igv_landmark_marginalize(igv_batch* h, int lm_slot);

/* StateManager::replaceVarLinear (StateManager.cpp:632-693). H: size(target) x n, col-major, B x ... */
igv_status igv_replace_var_linear(igv_batch* h, int target_idx, int target_size, int n_blocks,
                                  const int* blk_idx, const int* blk_size, const double* H);

/* ---- per-sequence read-outs (SURVEY.md §5 metrics) --------------------------------------------- */
igv_status igv_get_flags(igv_batch* h, int* flags_out /* B */, int clear);
igv_status igv_cov_trace(igv_batch* h, double* trace_out /* B */);

/* ---- measurement utilities (no reference counterpart; used by bench.py) --------------------------
 * Per-kernel-family CUDA-event timing on the handle's stream. While enabled every launch is
 * bracketed by an event pair; igv_profile_read synchronises and returns the accumulated device
 * milliseconds and launch counts per family. */
enum { IGV_K_PROPAGATE = 0, IGV_K_AUGMENT = 1, IGV_K_FEATURES = 2, IGV_K_QR = 3, IGV_K_EKF = 4,
       IGV_K_GNSS_ROWS = 5, IGV_K_MARG = 6, IGV_K_OTHER = 7, IGV_K_COUNT = 8 };
igv_status igv_profile_enable(igv_batch* h, int on);
igv_status igv_profile_read(igv_batch* h, double* ms_out /* IGV_K_COUNT */, long long* launches_out /* IGV_K_COUNT */,
                            int reset);
/* Sustained FP64 FMA throughput of the device (dependent-chain-free DFMA loop on every SM), in
 * TFLOP/s: the denominator for the FP64-pipe roofline of the dense kernels. */
igv_status igv_measure_fp64_peak(int device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* INGVIO_B200_H */
