// gnss_comm residual generator on the device: pseudo-range / Doppler residuals, receiver->satellite unit vectors,
// azimuth / elevation, ionospheric and tropospheric delays and the measurement sigmas of one GNSS epoch, for every
// satellite of every sequence (one thread each). Its outputs are exactly the inputs of igv_gnss_update, so a real
// epoch (satellite states + raw observations) can enter the device path without a host round trip
// (SURVEY.md section 8f rank 2).
//
// Reference (vendored HKUST gnss_comm): psr_res gnss_spp.cpp:99-146, dopp_res :256-282, sat_azel
// gnss_utility.cpp:762-771, ecef2geo :347-387, ecef2enu :722-735, nmf / interpc / mapf :774-839,
// calculate_trop_delay :841-863, calculate_ion_delay :865-901; receiver state assembly GnssUpdate.cpp:101-111
// (xyzt = T_enu2ecef T_w2enu(yof) p | clock biases, dopp = R_enu2ecef R_w2enu(yof) v | clock drift) and the noise
// model GnssUpdate.cpp:177-186, :246-255.
//
// Boundary: satellite states come from gnss_comm::sat_states (ephemeris evaluation, host) and the transmit time is
// given as day-of-year and GPS seconds of week (time2doy / time2gpst, host calendar code). freq <= 0 marks a
// satellite without an L1 observation: its outputs stay zero like the reference's `continue`.
#include "igv_device.cuh"

using namespace igv;

namespace {

constexpr double kC = 2.99792458e8;            // LIGHT_SPEED        gnss_constant.hpp:214
constexpr double kOmg = 7.2921151467e-5;       // EARTH_OMG_GPS      gnss_constant.hpp:208
constexpr double kE2 = 6.69437999014e-3;       // EARTH_ECCE_2       gnss_constant.hpp:203
constexpr double kA = 6378137.0;               // EARTH_SEMI_MAJOR   gnss_constant.hpp:205
constexpr double kPi = 3.14159265358979323846;
constexpr double kD2R = kPi / 180.0, kR2D = 180.0 / kPi;

struct ResArgs {
  const double* X; int xsize; int B, S;
  int idx_gnss[6];
  const double* sat_pos; const double* sat_vel; const double* sat_clk; const double* obs; const double* obs_std;
  const double* ttx; const int* sys; const double* T; const double* iono;
  double psr_amp, dopp_amp;
  double* unit; double* res_pos; double* res_vel; double* sig_psr; double* sig_dopp; double* azel; double* atmos;
};

__device__ void ecef2geo(const double* xyz, double* lla) {   // gnss_utility.cpp:347-387
  lla[0] = lla[1] = lla[2] = 0.0;
  if (xyz[0] == 0.0 && xyz[1] == 0.0) return;
  const double a2 = kA * kA, b2 = a2 * (1 - kE2), b = sqrt(b2), ep2 = (a2 - b2) / b2;
  const double p = sqrt(xyz[0] * xyz[0] + xyz[1] * xyz[1]);
  double s1 = xyz[2] * kA, s2 = p * b, h = sqrt(s1 * s1 + s2 * s2);
  const double st = s1 / h, ct = s2 / h;
  s1 = xyz[2] + ep2 * b * (st * st * st);
  s2 = p - kA * kE2 * (ct * ct * ct);
  h = sqrt(s1 * s1 + s2 * s2);
  const double sin_lat = s1 / h, cos_lat = s2 / h;
  const double N = a2 / sqrt(a2 * cos_lat * cos_lat + b2 * sin_lat * sin_lat);
  lla[0] = atan(s1 / s2) * kR2D;
  lla[1] = atan2(xyz[1], xyz[0]) * kR2D;
  lla[2] = p / cos_lat - N;
}

__device__ double interpc(const double* coef, double lat) {  // gnss_utility.cpp:774-779
  const int i = (int)(lat / 15.0);
  if (i < 1) return coef[0];
  if (i > 4) return coef[4];
  return coef[i - 1] * (1.0 - lat / 15.0 + i) + coef[i] * (lat / 15.0 - i);
}
__device__ double mapf(double el, double a, double b, double c) {  // gnss_utility.cpp:782-786
  const double s = sin(el);
  return (1.0 + a / (1.0 + b / (1.0 + c))) / (s + (a / (s + b / (s + c))));
}

__constant__ double c_nmf[9][5] = {
    {1.2769934E-3, 1.2683230E-3, 1.2465397E-3, 1.2196049E-3, 1.2045996E-3},
    {2.9153695E-3, 2.9152299E-3, 2.9288445E-3, 2.9022565E-3, 2.9024912E-3},
    {62.610505E-3, 62.837393E-3, 63.721774E-3, 63.824265E-3, 64.258455E-3},
    {0.0000000E-0, 1.2709626E-5, 2.6523662E-5, 3.4000452E-5, 4.1202191E-5},
    {0.0000000E-0, 2.1414979E-5, 3.0160779E-5, 7.2562722E-5, 11.723375E-5},
    {0.0000000E-0, 9.0128400E-5, 4.3497037E-5, 84.795348E-5, 170.37206E-5},
    {5.8021897E-4, 5.6794847E-4, 5.8118019E-4, 5.9727542E-4, 6.1641693E-4},
    {1.4275268E-3, 1.5138625E-3, 1.4572752E-3, 1.5007428E-3, 1.7599082E-3},
    {4.3472961E-2, 4.6729510E-2, 4.3908931E-2, 4.4626982E-2, 5.4736038E-2}};

__device__ double trop_delay(double doy, const double* lla, const double* azel) {  // gnss_utility.cpp:798-863
  if (lla[2] < -100.0 || 1E4 < lla[2] || azel[1] <= 0) return 0.0;
  const double hgt = lla[2] < 0.0 ? 0.0 : lla[2];
  const double pres = 1013.25 * pow(1.0 - 2.2557E-5 * hgt, 5.2568);
  const double temp = 15.0 - 6.5E-3 * hgt + 273.16;
  const double e = 6.108 * 0.7 * exp((17.15 * temp - 4684.0) / (temp - 38.45));
  const double zhd = 0.0022768 * pres / (1.0 - 0.00266 * cos(2.0 * lla[0] * kD2R) - 0.00028 * hgt / 1E3);
  const double zwd = 0.002277 * (1255.0 / temp + 0.05) * e;
  // nmf: the height correction uses the ellipsoidal height lla[2] itself (:835)
  const double el = azel[1];
  double lat = lla[0];
  const double y = (doy - 28.0) / 365.25 + (lat < 0.0 ? 0.5 : 0.0);
  const double cosy = cos(2.0 * kPi * y);
  lat = fabs(lat);
  double ah[3], aw[3];
  for (int i = 0; i < 3; ++i) {
    ah[i] = interpc(c_nmf[i], lat) - interpc(c_nmf[i + 3], lat) * cosy;
    aw[i] = interpc(c_nmf[i + 6], lat);
  }
  const double dm = (1.0 / sin(el) - mapf(el, 2.53E-5, 5.49E-3, 1.14E-3)) * lla[2] / 1E3;
  const double mapfw = mapf(el, aw[0], aw[1], aw[2]);
  const double mapfh = mapf(el, ah[0], ah[1], ah[2]) + dm;
  return mapfh * zhd + mapfw * zwd;
}

__device__ double ion_delay(double tow, const double* ion, const double* lla, const double* azel) {  // :865-901
  if (!ion) return 0.0;
  if (lla[2] < -1E3 || azel[1] <= 0) return 0.0;
  const double psi = 0.0137 / (azel[1] / kPi + 0.11) - 0.022;
  double phi = lla[0] / 180.0 + psi * cos(azel[0]);
  if (phi > 0.416) phi = 0.416; else if (phi < -0.416) phi = -0.416;
  const double lam = lla[1] / 180.0 + psi * sin(azel[0]) / cos(phi * kPi);
  phi += 0.064 * cos((lam - 1.617) * kPi);
  double tt = 43200.0 * lam + tow;
  tt -= floor(tt / 86400.0) * 86400.0;
  const double f = 1.0 + 16.0 * pow(0.53 - azel[1] / kPi, 3.0);
  double amp = ion[0] + phi * (ion[1] + phi * (ion[2] + phi * ion[3]));
  double per = ion[4] + phi * (ion[5] + phi * (ion[6] + phi * ion[7]));
  amp = amp < 0.0 ? 0.0 : amp;
  per = per < 72000.0 ? 72000.0 : per;
  const double x = 2.0 * kPi * (tt - 50400.0) / per;
  return kC * f * (fabs(x) < 1.57 ? 5E-9 + amp * (1.0 + x * x * (-0.5 + x * x / 24.0)) : 5E-9);
}

__global__ void __launch_bounds__(128) k_gnss_residuals(ResArgs a) {
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)a.B * a.S) return;
  const int b = (int)(gid / a.S);
  const double* Xb = a.X + (size_t)b * a.xsize;
  const double* T = a.T + (size_t)b * 12;
  // receiver state (GnssUpdate.cpp:101-111)
  double sy, cy;
  sincos(Xb[33 + IGV_GNSS_YOF], &sy, &cy);
  const double pe[3] = {cy * Xb[9] - sy * Xb[10], sy * Xb[9] + cy * Xb[10], Xb[11]};     // Rz(yof) p
  const double ve[3] = {cy * Xb[12] - sy * Xb[13], sy * Xb[12] + cy * Xb[13], Xb[14]};   // Rz(yof) v
  double rp[3], rv[3];
  mat3_vec(T, pe, rp);
  mat3_vec(T, ve, rv);
  for (int i = 0; i < 3; ++i) rp[i] += T[9 + i];
  const double fs = (a.idx_gnss[IGV_GNSS_FS] >= 0) ? Xb[33 + IGV_GNSS_FS] : 0.0;
  const double* sp = a.sat_pos + gid * 3;
  const double* sv = a.sat_vel + gid * 3;
  const double sdt = a.sat_clk[gid * 3], sddt = a.sat_clk[gid * 3 + 1], tgd = a.sat_clk[gid * 3 + 2];
  const double psr = a.obs[gid * 3], dopp = a.obs[gid * 3 + 1], freq = a.obs[gid * 3 + 2];
  const int k = a.sys[gid];
  const bool l1 = freq > 0.0 && k >= 0 && k < 4;
  double unit[3] = {0, 0, 0}, res_p = 0.0, res_v = 0.0, azel[2] = {0.0, 0.0}, ion_d = 0.0, tro_d = 0.0;
  if (l1) {
    const double d[3] = {sp[0] - rp[0], sp[1] - rp[1], sp[2] - rp[2]};
    const double rng = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; ++i) unit[i] = d[i] / rng;
    azel[1] = kPi / 2.0;
    if (sqrt(rp[0] * rp[0] + rp[1] * rp[1] + rp[2] * rp[2]) > 0.0) {
      double lla[3];
      ecef2geo(rp, lla);
      // sat_azel (gnss_utility.cpp:762-771) with ecef2enu (:722-735)
      double slat, clat, slon, clon;
      sincos(lla[0] * kD2R, &slat, &clat);
      sincos(lla[1] * kD2R, &slon, &clon);
      const double e_ = -slon * unit[0] + clon * unit[1];
      const double n_ = -slat * clon * unit[0] - slat * slon * unit[1] + clat * unit[2];
      const double u_ = clat * clon * unit[0] + clat * slon * unit[1] + slat * unit[2];
      azel[0] = sqrt(unit[0] * unit[0] + unit[1] * unit[1]) < 1e-12 ? 0.0 : atan2(e_, n_);
      if (azel[0] < 0) azel[0] += 2 * kPi;
      azel[1] = asin(u_);
      tro_d = trop_delay(a.ttx[gid * 2], lla, azel);
      ion_d = ion_delay(a.ttx[gid * 2 + 1], a.iono ? a.iono + (size_t)b * 8 : nullptr, lla, azel);
    }
    const double cb = (a.idx_gnss[k] >= 0) ? Xb[33 + k] : 0.0;   // GnssManager::getClockbiasVec
    const double sagnac = kOmg * (sp[0] * rp[1] - sp[1] * rp[0]) / kC;
    const double est = rng + sagnac + cb - sdt * kC + tro_d + ion_d + tgd * kC;
    res_p = est - psr;
    const double sag_v = kOmg / kC * (sv[0] * rp[1] + sp[0] * rv[1] - sv[1] * rp[0] - sp[1] * rv[0]);
    const double est_v = (sv[0] - rv[0]) * unit[0] + (sv[1] - rv[1]) * unit[1] + (sv[2] - rv[2]) * unit[2] + fs + sag_v -
                         sddt * kC;
    res_v = est_v + dopp * (kC / freq);
  }
  double sin_el = sin(azel[1]);
  if (fabs(sin_el) < 1e-6) sin_el = 1e-6;
  const double ura = a.obs_std[gid * 3], npr = a.obs_std[gid * 3 + 1];
  const double ndp = a.obs_std[gid * 3 + 2] * kC / (freq > 0.0 ? freq : 1.0);
  for (int i = 0; i < 3; ++i) a.unit[gid * 3 + i] = unit[i];
  a.res_pos[gid] = res_p;
  a.res_vel[gid] = res_v;
  a.sig_psr[gid] = a.psr_amp * sqrt(ura * npr / (sin_el * sin_el));
  a.sig_dopp[gid] = a.dopp_amp * sqrt(ura * ndp / (sin_el * sin_el));
  if (a.azel) { a.azel[gid * 2] = azel[0]; a.azel[gid * 2 + 1] = azel[1]; }
  if (a.atmos) { a.atmos[gid * 2] = ion_d; a.atmos[gid * 2 + 1] = tro_d; }
}

}  // namespace

void igv_launch_gnss_residuals(igv_batch* h, const IgvGnssResLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_GNSS_ROWS);
  ResArgs a;
  a.X = h->Xc(); a.xsize = h->xsize; a.B = h->B; a.S = l.S;
  IgvLayout L = h->layout();
  for (int i = 0; i < 6; ++i) a.idx_gnss[i] = L.idx_gnss[i];
  a.sat_pos = l.sat_pos; a.sat_vel = l.sat_vel; a.sat_clk = l.sat_clk; a.obs = l.obs; a.obs_std = l.obs_std;
  a.ttx = l.ttx; a.sys = l.sys; a.T = l.T; a.iono = l.iono; a.psr_amp = l.psr_amp; a.dopp_amp = l.dopp_amp;
  a.unit = l.unit; a.res_pos = l.res_pos; a.res_vel = l.res_vel; a.sig_psr = l.sig_psr; a.sig_dopp = l.sig_dopp;
  a.azel = l.azel; a.atmos = l.atmos;
  const long n = (long)h->B * l.S;
  k_gnss_residuals<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a);
  h->launches++;
}
