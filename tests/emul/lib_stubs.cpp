// TEST INFRASTRUCTURE: entry points the CPU model does not provide (the FP64 throughput probe of k_peak.cu).
#include "../../ingvio_b200/csrc/igv_internal.h"

extern "C" void igv_emul_set_last_error(int e);
static void unavailable() { igv_emul_set_last_error((int)cudaErrorNotSupported); }

extern "C" igv_status igv_measure_fp64_peak(int, double* out) { if (out) *out = 0.0; return IGV_ERR_CUDA; }
