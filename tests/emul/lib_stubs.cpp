// TEST INFRASTRUCTURE: launchers of the kernels the CPU model does not execute (the DMMA kernels: propagation, EKF update,
// per-track MSCKF kernel, compression). Calling one makes the C-ABI call that needed it fail with IGV_ERR_CUDA, loudly.
#include "../../ingvio_b200/csrc/igv_internal.h"

extern "C" void igv_emul_set_last_error(int e);
static void unavailable() { igv_emul_set_last_error((int)cudaErrorNotSupported); }

void igv_launch_propagate(igv_batch*, int, const double*, const double*, const double*, const double*, const double*) { unavailable(); }
void igv_launch_ekf(igv_batch*, const IgvEkfLaunch&) { unavailable(); }
void igv_launch_msckf_features(igv_batch*, const IgvMsckfLaunch&) { unavailable(); }
void igv_launch_qr_compress(igv_batch*, int, int) { unavailable(); }
int igv_gram_n1p(int ncols_max) { return 24 * ((ncols_max + 1 + 23) / 24) + 8; }
bool igv_gram_supported(int) { return true; }
void igv_launch_gram_compress(igv_batch*, int, int, int) { unavailable(); }
void igv_launch_gram_factor(igv_batch*, int) { unavailable(); }
extern "C" igv_status igv_measure_fp64_peak(int, double* out) { if (out) *out = 0.0; return IGV_ERR_CUDA; }
