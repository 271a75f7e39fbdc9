"""ctypes binding of libingvio_b200.so (the C-ABI declared in include/ingvio_b200.h).

The CUDA library is mandatory: `load()` raises if it has not been built
(`python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback in the product.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libingvio_b200.so")

IGV_OK, IGV_ERR_INVALID, IGV_ERR_CUDA, IGV_ERR_STATE, IGV_ERR_CAPACITY = range(5)
PREC_FP64, PREC_FP32_STACK, PREC_TF32_GRAM = 0, 1, 2
IGV_PTR_HOST, IGV_PTR_DEVICE = 0, 1
GPS, GLO, GAL, BDS, FS, YOF = range(6)
R_ISO, R_DIAG, R_FULL = 0, 1, 2
VIS_ALL_OBS, VIS_SELECTED = 0, 1
COMPRESS_AUTO, COMPRESS_HOUSEHOLDER, COMPRESS_GRAM = 0, 1, 2
FLAG_NEG_DIAG, FLAG_CHOL_FAIL, FLAG_GNSS_REJECTED, FLAG_TRACKS_FULL, FLAG_GATHER_CUT, FLAG_WEAK_PIVOT = 1, 2, 4, 8, 16, 32
TRK_LOST, TRK_SEEN_AT = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_ucp = C.POINTER(C.c_ubyte)


class igv_config(C.Structure):
    _fields_ = [("batch", C.c_int), ("max_dim", C.c_int), ("max_clones", C.c_int), ("max_feats", C.c_int),
                ("max_sats", C.c_int), ("stereo", C.c_int), ("device", C.c_int), ("stream", C.c_void_p),
                ("max_landmarks", C.c_int)]


class igv_params(C.Structure):
    _fields_ = [("noise_g", C.c_double), ("noise_a", C.c_double), ("noise_bg", C.c_double),
                ("noise_ba", C.c_double), ("noise_clockbias", C.c_double), ("noise_cb_rw", C.c_double),
                ("gravity", C.c_double * 3), ("T_cl2cr_R", C.c_double * 9), ("T_cl2cr_p", C.c_double * 3)]


class igv_msckf_args(C.Structure):
    _fields_ = [("mode", C.c_int), ("n_feats", C.c_int), ("pf_w", C.c_void_p), ("anchor_slot", C.c_void_p),
                ("obs", C.c_void_p), ("obs_mask", C.c_void_p), ("chi2_dof", C.c_void_p), ("obs_slots", C.c_int),
                ("noise", C.c_double), ("max_valid", C.c_int), ("dx_out", C.c_void_p),
                ("n_accepted_out", C.c_void_p), ("gamma_out", C.c_void_p), ("feat_ok", C.c_void_p)]


class igv_tri_params(C.Structure):
    _fields_ = [("trans_thres", C.c_double), ("huber_epsilon", C.c_double), ("conv_precision", C.c_double),
                ("init_damping", C.c_double), ("outer_loop_max_iter", C.c_int), ("inner_loop_max_iter", C.c_int),
                ("max_depth", C.c_double), ("min_depth", C.c_double)]


class igv_tri_args(C.Structure):
    _fields_ = [("n_feats", C.c_int), ("obs", C.c_void_p), ("obs_mask", C.c_void_p), ("obs_slots", C.c_int),
                ("anchor_slot", C.c_void_p), ("prm", igv_tri_params), ("pf_out", C.c_void_p), ("ok_out", C.c_void_p)]


class igv_gnss_args(C.Structure):
    _fields_ = [("n_sats", C.c_int), ("unit", C.c_void_p), ("res_pos", C.c_void_p), ("res_vel", C.c_void_p),
                ("sigma_psr", C.c_void_p), ("sigma_dopp", C.c_void_p), ("sys", C.c_void_p),
                ("R_enu2ecef", C.c_void_p), ("is_adjust_yof", C.c_int), ("chi2_test", C.c_int),
                ("strong_reject", C.c_int), ("dx_out", C.c_void_p)]


class igv_frame_args(C.Structure):
    _fields_ = [("n_imu", C.c_int), ("gyro", C.c_void_p), ("accel", C.c_void_p), ("dt", C.c_void_p),
                ("visual", C.c_void_p), ("n_marg", C.c_int), ("marg_slots", c_ip), ("gnss", C.c_void_p)]


class igv_gnss_new_sys_args(C.Structure):
    _fields_ = [("n_sats", C.c_int), ("gtype", C.c_int), ("value", C.c_void_p), ("unit", C.c_void_p),
                ("res_pos", C.c_void_p), ("res_vel", C.c_void_p), ("sigma_psr", C.c_void_p), ("sigma_dopp", C.c_void_p),
                ("sys", C.c_void_p), ("R_enu2ecef", C.c_void_p), ("R_ecef2enu", C.c_void_p), ("is_adjust_yof", C.c_int),
                ("chi2_mult", C.c_double), ("prior_cov_if_rejected", C.c_double), ("accepted_out", C.c_void_p),
                ("dx_out", C.c_void_p)]


class igv_lm_init_args(C.Structure):
    _fields_ = [("obs_slots", C.c_int), ("anchor_slot", C.c_int), ("pf_w", C.c_void_p), ("obs", C.c_void_p),
                ("obs_mask", C.c_void_p), ("noise", C.c_double), ("chi2_mult", C.c_double),
                ("prior_cov_if_rejected", C.c_double), ("accepted_out", C.c_void_p)]


class igv_lm_update_args(C.Structure):
    _fields_ = [("uv", C.c_void_p), ("valid", C.c_void_p), ("noise", C.c_double), ("n_accepted_out", C.c_void_p),
                ("gamma_out", C.c_void_p)]


class igv_gnss_res_args(C.Structure):
    _fields_ = [("n_sats", C.c_int), ("sat_pos", C.c_void_p), ("sat_vel", C.c_void_p), ("sat_clk", C.c_void_p),
                ("obs", C.c_void_p), ("obs_std", C.c_void_p), ("ttx", C.c_void_p), ("sys", C.c_void_p),
                ("T_enu2ecef", C.c_void_p), ("iono", C.c_void_p), ("psr_noise_amp", C.c_double),
                ("dopp_noise_amp", C.c_double), ("unit", C.c_void_p), ("res_pos", C.c_void_p), ("res_vel", C.c_void_p),
                ("sigma_psr", C.c_void_p), ("sigma_dopp", C.c_void_p), ("azel", C.c_void_p), ("atmos", C.c_void_p),
                ("clock_init", C.c_void_p)]


class igv_sat_state_args(C.Structure):
    _fields_ = [("n_sats", C.c_int), ("eph", C.c_void_p), ("t_obs_rel", C.c_void_p), ("psr", C.c_void_p),
                ("sys", C.c_void_p), ("sat_pos", C.c_void_p), ("sat_vel", C.c_void_p), ("sat_clk", C.c_void_p),
                ("ttx_rel", C.c_void_p)]


class igv_track_gather_args(C.Structure):
    _fields_ = [("rule", C.c_int), ("n_selected", C.c_int), ("selected_slots", c_ip), ("min_obs", C.c_int),
                ("dof_fixed", C.c_int), ("n_feats", C.c_int), ("obs_slots", C.c_int), ("track_entry", C.c_void_p),
                ("n_sel", C.c_void_p), ("track_id", C.c_void_p), ("obs", C.c_void_p), ("mask_all", C.c_void_p),
                ("mask_upd", C.c_void_p), ("anchor_slot", C.c_void_p), ("chi2_dof", C.c_void_p),
                ("feat_ok", C.c_void_p)]


class igv_track_dump(C.Structure):
    _fields_ = [("obs_slots", C.c_int), ("id", C.c_void_p), ("used", C.c_void_p), ("to_marg", C.c_void_p),
                ("is_tri", C.c_void_p), ("slot_mask", C.c_void_p), ("anchor_slot", C.c_void_p), ("pf", C.c_void_p),
                ("pf_fej", C.c_void_p), ("obs", C.c_void_p), ("n_tracks", C.c_void_p)]


EPH_STRIDE = 24

# every symbol include/ingvio_b200.h declares: (restype, argtypes)
_H = C.c_void_p
_VP = C.c_void_p
SIGNATURES = {
    "igv_create": (C.c_int, [C.POINTER(igv_config), C.POINTER(_H)]),
    "igv_destroy": (C.c_int, [_H]),
    "igv_last_error": (C.c_char_p, [_H]),
    "igv_set_pointer_mode": (C.c_int, [_H, C.c_int]),
    "igv_synchronize": (C.c_int, [_H]),
    "igv_state_get_async": (C.c_int, [_H, C.c_void_p]),
    "igv_cov_trace_async": (C.c_int, [_H, C.c_void_p]),
    "igv_fence_record": (C.c_int, [_H, C.c_int]),
    "igv_fence_wait": (C.c_int, [_H, C.c_int]),
    "igv_set_compression": (C.c_int, [_H, C.c_int]),
    "igv_last_visual_path": (C.c_int, [_H]),
    "igv_launch_count": (C.c_longlong, [_H]),
    "igv_set_params": (C.c_int, [_H, C.POINTER(igv_params)]),
    "igv_set_precision": (C.c_int, [_H, C.c_int]),
    "igv_last_gram_tensor": (C.c_int, [_H]),
    "igv_set_chi2_table": (C.c_int, [_H, c_dp, C.c_int]),
    "igv_chi2_quantile": (C.c_double, [C.c_double, C.c_int]),
    "igv_state_init": (C.c_int, [_H] + [_VP] * 7 + [c_dp]),
    "igv_dim": (C.c_int, [_H]),
    "igv_num_variables": (C.c_int, [_H]),
    "igv_num_clones": (C.c_int, [_H]),
    "igv_clone_idx": (C.c_int, [_H, C.c_int]),
    "igv_gnss_idx": (C.c_int, [_H, C.c_int]),
    "igv_state_size": (C.c_int, [_H]),
    "igv_state_get": (C.c_int, [_H, _VP]),
    "igv_state_set": (C.c_int, [_H, _VP]),
    "igv_cov_get": (C.c_int, [_H, _VP, C.c_int]),
    "igv_cov_set": (C.c_int, [_H, _VP, C.c_int]),
    "igv_cov_get_blocks": (C.c_int, [_H, C.c_int, c_ip, c_ip, _VP]),
    "igv_add_gnss_variable": (C.c_int, [_H, C.c_int, _VP, C.c_double]),
    "igv_marg_gnss_variable": (C.c_int, [_H, C.c_int]),
    "igv_add_variable_independent": (C.c_int, [_H, C.c_int, c_dp]),
    "igv_marginalize": (C.c_int, [_H, C.c_int]),
    "igv_marginalize_clone": (C.c_int, [_H, C.c_int]),
    "igv_propagate_cov": (C.c_int, [_H, _VP, _VP, _VP]),
    "igv_propagate_imu": (C.c_int, [_H, C.c_int, _VP, _VP, _VP]),
    "igv_augment_clone": (C.c_int, [_H]),
    "igv_augment_clone_cov": (C.c_int, [_H, _VP, _VP, _VP]),
    "igv_ekf_update": (C.c_int, [_H, C.c_int, c_ip, c_ip, C.c_int, _VP, C.c_int, _VP, _VP, C.c_int, _VP]),
    "igv_chi2_whiten": (C.c_int, [_H, C.c_int, c_ip, c_ip, C.c_int, _VP, C.c_int, _VP, _VP, C.c_int, _VP]),
    "igv_box_plus": (C.c_int, [_H, _VP]),
    "igv_msckf_update": (C.c_int, [_H, C.POINTER(igv_msckf_args)]),
    "igv_gnss_update": (C.c_int, [_H, C.POINTER(igv_gnss_args)]),
    "igv_gnss_add_new_tracked_sys": (C.c_int, [_H, C.POINTER(igv_gnss_new_sys_args)]),
    "igv_frame_step": (C.c_int, [_H, C.POINTER(igv_frame_args)]),
    "igv_graph_replays": (C.c_longlong, [_H]),
    "igv_num_landmarks": (C.c_int, [_H]),
    "igv_landmark_idx": (C.c_int, [_H, C.c_int]),
    "igv_landmark_anchor": (C.c_int, [_H, C.c_int]),
    "igv_landmark_init": (C.c_int, [_H, C.POINTER(igv_lm_init_args)]),
    "igv_landmark_update": (C.c_int, [_H, C.POINTER(igv_lm_update_args)]),
    "igv_landmark_change_anchor": (C.c_int, [_H, C.c_int, C.c_int]),
    "igv_landmark_marginalize": (C.c_int, [_H, C.c_int]),
    "igv_gnss_residuals": (C.c_int, [_H, C.POINTER(igv_gnss_res_args)]),
    "igv_sat_states": (C.c_int, [_H, C.POINTER(igv_sat_state_args)]),
    "igv_triangulate": (C.c_int, [_H, C.POINTER(igv_tri_args)]),
    "igv_add_variable_delayed": (C.c_int, [_H, C.c_int, _VP, C.c_int, c_ip, c_ip, C.c_int, _VP, _VP, _VP,
                                           C.c_double, C.c_double, C.c_int, C.c_double, _VP, _VP]),
    "igv_replace_var_linear": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, c_ip, c_ip, _VP]),
    "igv_tracks_create": (C.c_int, [_H, C.c_int]),
    "igv_tracks_reset": (C.c_int, [_H]),
    "igv_tracks_capacity": (C.c_int, [_H]),
    "igv_tracks_collect": (C.c_int, [_H, _VP, C.c_int, _VP, _VP]),
    "igv_tracks_mark_lost": (C.c_int, [_H]),
    "igv_tracks_gather": (C.c_int, [_H, C.POINTER(igv_track_gather_args)]),
    "igv_tracks_commit_tri": (C.c_int, [_H, C.c_int, _VP, _VP, _VP, _VP]),
    "igv_tracks_erase": (C.c_int, [_H, C.c_int, _VP]),
    "igv_tracks_clean_obs": (C.c_int, [_H, C.c_int, c_ip]),
    "igv_tracks_change_anchor": (C.c_int, [_H, C.c_int, c_ip, C.c_double]),
    "igv_tracks_erase_invalid": (C.c_int, [_H, C.c_double]),
    "igv_tracks_get": (C.c_int, [_H, C.POINTER(igv_track_dump)]),
    "igv_get_flags": (C.c_int, [_H, _VP, C.c_int]),
    "igv_cov_trace": (C.c_int, [_H, _VP]),
    "igv_profile_enable": (C.c_int, [_H, C.c_int]),
    "igv_profile_read": (C.c_int, [_H, c_dp, C.POINTER(C.c_longlong), C.c_int]),
    "igv_measure_fp64_peak": (C.c_int, [C.c_int, c_dp]),
}
KERNEL_FAMILIES = ["propagate", "augment", "features", "qr", "ekf", "gnss_rows", "marginalize", "other"]

_lib = None


class IgvError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"igv status {status}: {msg}")
        self.status = status


def load():
    """Loads the shared library and declares every prototype. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def header_symbols():
    """Names declared in include/ingvio_b200.h (parsed), for the export test."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "ingvio_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(igv_[a-z0-9_]+)\s*\(", txt)))
