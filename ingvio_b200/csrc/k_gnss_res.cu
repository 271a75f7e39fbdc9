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
  const double* clock_init;
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
  // clock_init (optional, B x 5): initial values of receiver clock entries not yet in the state, as addNewTrackedSys
  // writes them into xyzt / dopp from the SPP solution (GnssUpdate.cpp:351-370); NaN = take the entry from the state
  const double* ci = a.clock_init ? a.clock_init + (size_t)b * 5 : nullptr;
  const double fs = (ci && !isnan(ci[4])) ? ci[4] : ((a.idx_gnss[IGV_GNSS_FS] >= 0) ? Xb[33 + IGV_GNSS_FS] : 0.0);
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
    const double cb = (ci && !isnan(ci[k])) ? ci[k] : ((a.idx_gnss[k] >= 0) ? Xb[33 + k] : 0.0);   // GnssManager::getClockbiasVec
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

// ---- ephemeris -> satellite state: gnss_comm::sat_states (gnss_spp.cpp:50-97) ---------------------------------
// One thread per satellite. Broadcast ephemerides: Kepler (GPS / GAL / BDS, eph2svdt :437-446, eph2pos :448-531,
// eph2vel :533-634, Kepler :390-405 -- which returns the previous iterate) or GLONASS (geph2svdt :684-696, RK4
// integration glo_orbit / deq :636-682 with TSTEP = 60 s, geph2pos / geph2vel :698-733). Times are seconds relative
// to the ephemeris epoch toe (the caller's time_diff(obs->time, toe)); field order = IGV_EPH_* of the header.
constexpr double kMuGps = 3.9860050000e14, kMu = 3.9860044180e14;          // gnss_constant.hpp:210-211
constexpr double kOmgGlo = 7.2921150000e-5, kOmgBds = 7.2921150000e-5;     // gnss_constant.hpp:207,209
constexpr double kJ2Glo = 1.0826257E-3, kAGlo = 6378136.0, kTstep = 60.0;  // gnss_constant.hpp:213,206,212
constexpr double kSinN5 = -0.0871557427476582, kCosN5 = 0.9961946980917456;
constexpr double kWeek = 604800.0;

struct SatArgs {
  int B, S;
  const double* eph; const double* t_obs; const double* psr; const int* sys;
  double* pos; double* vel; double* clk; double* ttx;
};

__device__ double kepler(double mk, double es) {
  double e = mk, ek = 1e6;
  for (int it = 0; it < 30 && fabs(e - ek) > 1e-14; ++it) {
    ek = e;
    e -= (e - es * sin(e) - mk) / (1.0 - es * cos(e));
  }
  return ek;   // sic (gnss_utility.cpp:404)
}
__device__ double wrap_week(double t) { return t > kWeek / 2 ? t - kWeek : (t < -kWeek / 2 ? t + kWeek : t); }

__device__ void glo_deq(const double* p, const double* v, const double* acc, double* pd, double* vd) {
  const double r2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  if (r2 <= 0.0) { for (int i = 0; i < 3; ++i) pd[i] = vd[i] = 0.0; return; }
  const double r3 = r2 * sqrt(r2), omg2 = kOmgGlo * kOmgGlo;
  const double a = 1.5 * kJ2Glo * kMu * kAGlo * kAGlo / r2 / r3;
  const double b = 5.0 * p[2] * p[2] / r2;
  const double c = -kMu / r3 - a * (1.0 - b);
  pd[0] = v[0]; pd[1] = v[1]; pd[2] = v[2];
  vd[0] = (c + omg2) * p[0] + 2.0 * kOmgGlo * v[1] + acc[0];
  vd[1] = (c + omg2) * p[1] - 2.0 * kOmgGlo * v[0] + acc[1];
  vd[2] = (c - 2.0 * a) * p[2] + acc[2];
}
__device__ void glo_orbit(double dt, double* p, double* v, const double* acc) {
  double p1[3], p2[3], p3[3], p4[3], v1[3], v2[3], v3[3], v4[3], np_[3], nv[3];
  glo_deq(p, v, acc, p1, v1);
  for (int i = 0; i < 3; ++i) { np_[i] = p[i] + 0.5 * p1[i] * dt; nv[i] = v[i] + 0.5 * v1[i] * dt; }
  glo_deq(np_, nv, acc, p2, v2);
  for (int i = 0; i < 3; ++i) { np_[i] = p[i] + 0.5 * p2[i] * dt; nv[i] = v[i] + 0.5 * v2[i] * dt; }
  glo_deq(np_, nv, acc, p3, v3);
  for (int i = 0; i < 3; ++i) { np_[i] = p[i] + p3[i] * dt; nv[i] = v[i] + v3[i] * dt; }
  glo_deq(np_, nv, acc, p4, v4);
  for (int i = 0; i < 3; ++i) {
    p[i] += (p1[i] + 2.0 * p2[i] + 2.0 * p3[i] + p4[i]) * dt / 6.0;
    v[i] += (v1[i] + 2.0 * v2[i] + 2.0 * v3[i] + v4[i]) * dt / 6.0;
  }
}

__global__ void __launch_bounds__(128) k_sat_states(SatArgs a) {
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long)a.B * a.S) return;
  const double* E = a.eph + gid * IGV_EPH_STRIDE;
  const int sys = a.sys[gid];
  const double psr = a.psr[gid];
  double pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, dts = 0.0, ddts = 0.0, tgd = 0.0, tx = 0.0;
  if (psr > 0.0 && sys >= 0 && sys < 4) {     // no L1 observation: the SatState stays all zero (gnss_spp.cpp:64-66)
    tx = a.t_obs[gid] - psr / kC;
    if (sys == IGV_GNSS_GLO) {
      const double tau_n = E[9], gamma = E[10];
      double dt = wrap_week(tx);
      for (int i = 0; i < 2; ++i) dt -= -tau_n + gamma * dt;
      tx -= -tau_n + gamma * dt;
      double acc[3] = {E[6], E[7], E[8]};
      for (int i = 0; i < 3; ++i) { pos[i] = E[i]; vel[i] = E[3 + i]; }
      dt = tx;
      dts = -tau_n + gamma * dt;
      ddts = gamma;
      for (double tt = dt < 0.0 ? -kTstep : kTstep; fabs(dt) > 1e-9; dt -= tt) {
        if (fabs(dt) < kTstep) tt = dt;
        glo_orbit(tt, pos, vel, acc);
      }
    } else {
      const double A = E[0], e = E[1], i0 = E[2], OMG0 = E[3], omg = E[4], M0 = E[5], delta_n = E[6], OMG_dot = E[7],
                   i_dot = E[8], cuc = E[9], cus = E[10], crc = E[11], crs = E[12], cic = E[13], cis = E[14], af0 = E[15],
                   af1 = E[16], af2 = E[17], toe_tow = E[18], toe_toc = E[20];
      const int prn = (int)E[21];
      tgd = E[19];
      double dt = tx + toe_toc;
      for (int i = 0; i < 2; ++i) dt -= af0 + af1 * dt + af2 * dt * dt;
      tx -= af0 + af1 * dt + af2 * dt * dt;
      const double tk = wrap_week(tx);
      const double mu = (sys == IGV_GNSS_GPS) ? kMuGps : kMu;
      const double omg_e = (sys == IGV_GNSS_BDS) ? kOmgBds : kOmg;
      const double n = sqrt(mu / (A * A * A)) + delta_n;
      const double Ek = kepler(M0 + n * tk, e);
      double sE, cE;
      sincos(Ek, &sE, &cE);
      const double Ed = n / (1 - e * cE);
      const double vd = sqrt(1 - e * e) * Ed / (1 - e * cE);
      const double vk = atan2(sqrt(1 - e * e) * sE, cE - e);
      const double phi = vk + omg;
      double s2, c2;
      sincos(2 * phi, &s2, &c2);
      const double dud = 2 * vd * (cus * c2 - cuc * s2), drd = 2 * vd * (crs * c2 - crc * s2),
                   did = 2 * vd * (cis * c2 - cic * s2);
      const double ukd = vd + dud, rkd = A * e * Ed * sE + drd, ikd = i_dot + did;
      const double uk = phi + cus * s2 + cuc * c2;
      const double rk = A * (1 - e * cE) + crs * s2 + crc * c2;
      const double ik = i0 + i_dot * tk + cis * s2 + cic * c2;
      double si, ci, su, cu;
      sincos(ik, &si, &ci);
      sincos(uk, &su, &cu);
      const double xp = rk * cu, yp = rk * su;
      const double xpd = rkd * cu - rk * ukd * su, ypd = rkd * su + rk * ukd * cu;
      if (sys == IGV_GNSS_BDS && prn <= 5) {   // BDS GEO
        const double O = OMG0 + OMG_dot * tk - omg_e * toe_tow;
        double sO, cO, so, co;
        sincos(O, &sO, &cO);
        sincos(omg_e * tk, &so, &co);
        const double Od = OMG_dot;
        const double t1 = xpd - yp * Od * ci, t2 = xp * Od + ypd * ci - yp * ikd * si;
        const double xg = xp * cO - yp * ci * sO, yg = xp * sO + yp * ci * cO, zg = yp * si;
        const double xgd = t1 * cO - t2 * sO, ygd = t1 * sO + t2 * cO, zgd = ypd * si + ypd * ikd * ci;
        const double sod = omg_e * co, cod = -omg_e * so;
        pos[0] = xg * co + yg * so * kCosN5 + zg * so * kSinN5;
        pos[1] = -xg * so + yg * co * kCosN5 + zg * co * kSinN5;
        pos[2] = -yg * kSinN5 + zg * kCosN5;
        vel[0] = xgd * co + xg * cod + ygd * so * kCosN5 + yg * sod * kCosN5 + zgd * so * kSinN5 + zg * sod * kSinN5;
        vel[1] = -xgd * so - xg * sod + ygd * co * kCosN5 + yg * cod * kCosN5 + zgd * co * kSinN5 + zg * cod * kSinN5;
        vel[2] = -ygd * kSinN5 + zgd * kCosN5;
      } else {
        const double O = OMG0 + (OMG_dot - omg_e) * tk - omg_e * toe_tow;
        double sO, cO;
        sincos(O, &sO, &cO);
        const double Od = OMG_dot - omg_e;
        const double t1 = xpd - yp * Od * ci, t2 = xp * Od + ypd * ci - yp * ikd * si;
        pos[0] = xp * cO - yp * ci * sO; pos[1] = xp * sO + yp * ci * cO; pos[2] = yp * si;
        vel[0] = t1 * cO - t2 * sO; vel[1] = t1 * sO + t2 * cO;
        vel[2] = ypd * si + ypd * ikd * ci;   // as the reference writes it (gnss_utility.cpp:624)
      }
      const double dtc = tx + toe_toc;
      dts = af0 + af1 * dtc + af2 * dtc * dtc - 2.0 * sqrt(mu * A) * e * sE / kC / kC;
      ddts = af1 + 2.0 * af2 * dtc - 2.0 * sqrt(mu * A) * e * cE * Ed / kC / kC;
    }
  }
  for (int i = 0; i < 3; ++i) { a.pos[gid * 3 + i] = pos[i]; a.vel[gid * 3 + i] = vel[i]; }
  a.clk[gid * 3] = dts; a.clk[gid * 3 + 1] = ddts; a.clk[gid * 3 + 2] = tgd;
  if (a.ttx) a.ttx[gid] = tx;
}

}  // namespace

#if !defined(IGV_EMULATE) || defined(IGV_EMULATE_LAUNCHERS)   // tests/emul: kernels only, or (full model) launchers too
void igv_launch_gnss_residuals(igv_batch* h, const IgvGnssResLaunch& l) {
  IgvProfScope prof_scope_(h, IGV_K_GNSS_ROWS);
  ResArgs a;
  a.X = h->Xc(); a.xsize = h->xsize; a.B = h->B; a.S = l.S;
  IgvLayout L = h->layout();
  for (int i = 0; i < 6; ++i) a.idx_gnss[i] = L.idx_gnss[i];
  a.sat_pos = l.sat_pos; a.sat_vel = l.sat_vel; a.sat_clk = l.sat_clk; a.obs = l.obs; a.obs_std = l.obs_std;
  a.ttx = l.ttx; a.sys = l.sys; a.T = l.T; a.iono = l.iono; a.psr_amp = l.psr_amp; a.dopp_amp = l.dopp_amp;
  a.unit = l.unit; a.res_pos = l.res_pos; a.res_vel = l.res_vel; a.sig_psr = l.sig_psr; a.sig_dopp = l.sig_dopp;
  a.azel = l.azel; a.atmos = l.atmos; a.clock_init = l.clock_init;
  const long n = (long)h->B * l.S;
  k_gnss_residuals<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a);
  h->launches++;
}

void igv_launch_sat_states(igv_batch* h, int S, const double* eph, const double* t_obs, const double* psr, const int* sys,
                           double* pos, double* vel, double* clk, double* ttx) {
  IgvProfScope prof_scope_(h, IGV_K_GNSS_ROWS);
  SatArgs a;
  a.B = h->B; a.S = S; a.eph = eph; a.t_obs = t_obs; a.psr = psr; a.sys = sys;
  a.pos = pos; a.vel = vel; a.clk = clk; a.ttx = ttx;
  const long n = (long)h->B * S;
  k_sat_states<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a);
  h->launches++;
}
#endif  // IGV_EMULATE
