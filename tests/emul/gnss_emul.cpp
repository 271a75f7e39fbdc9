// TEST INFRASTRUCTURE: k_sat_states / k_gnss_residuals of ingvio_b200/csrc/k_gnss_res.cu (thread per satellite, plain FP64)
// executed on the CPU through cuda_emul.h, for tests/test_gnss_emulated.py. Same source as the library build.
#define IGV_EMULATE 1
#include "cuda_emul.h"

#include "../../ingvio_b200/csrc/k_gnss_res.cu"

extern "C" {
void emu_sat_states(int B, int S, const double* eph, const double* t_obs, const double* psr, const int* sys, double* pos, double* vel,
                    double* clk, double* ttx) {
  SatArgs a;
  a.B = B; a.S = S; a.eph = eph; a.t_obs = t_obs; a.psr = psr; a.sys = sys; a.pos = pos; a.vel = vel; a.clk = clk; a.ttx = ttx;
  const long n = (long)B * S;
  emul::launch((unsigned)((n + 127) / 128), 128, 0, [&] { k_sat_states(a); });
}
void emu_gnss_residuals(int B, int S, const double* X, int xsize, const int* idx_gnss, const double* sat_pos, const double* sat_vel,
                        const double* sat_clk, const double* obs, const double* obs_std, const double* ttx, const int* sys,
                        const double* T, const double* iono, double psr_amp, double dopp_amp, double* unit, double* res_pos,
                        double* res_vel, double* sig_psr, double* sig_dopp, double* azel, double* atmos) {
  ResArgs a;
  a.X = X; a.xsize = xsize; a.B = B; a.S = S;
  for (int i = 0; i < 6; ++i) a.idx_gnss[i] = idx_gnss[i];
  a.sat_pos = sat_pos; a.sat_vel = sat_vel; a.sat_clk = sat_clk; a.obs = obs; a.obs_std = obs_std; a.ttx = ttx; a.sys = sys; a.T = T;
  a.iono = iono; a.psr_amp = psr_amp; a.dopp_amp = dopp_amp;
  a.unit = unit; a.res_pos = res_pos; a.res_vel = res_vel; a.sig_psr = sig_psr; a.sig_dopp = sig_dopp; a.azel = azel; a.atmos = atmos;
  a.clock_init = nullptr;
  const long n = (long)B * S;
  emul::launch((unsigned)((n + 127) / 128), 128, 0, [&] { k_gnss_residuals(a); });
}
}
