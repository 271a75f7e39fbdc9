"""gnss_comm residual generator: pseudo-range / Doppler residuals, unit vectors, azimuth/elevation and
atmospheric delays from satellite states (the step before the GNSS rows).

TEST INFRASTRUCTURE (oracle). CPU restatement of the vendored HKUST library
/root/reference/gnss_comm/src/gnss_spp.cpp:99-146 (psr_res), :256-282 (dopp_res) and
/root/reference/gnss_comm/src/gnss_utility.cpp:347-387 (ecef2geo), :722-735 (ecef2enu), :762-771 (sat_azel),
:774-839 (interpc, mapf, nmf), :841-863 (calculate_trop_delay), :865-901 (calculate_ion_delay), plus the
receiver state assembly of /root/reference/ingvio_estimator/src/GnssUpdate.cpp:101-111 and the noise model
of :177-186, :246-255.

Boundary. psr_res / dopp_res take the outputs of gnss_comm::sat_states (ephemeris -> satellite position / velocity /
clock, gnss_spp.cpp:50-97, restated at the end of this module with eph2pos / eph2vel / geph2pos of
gnss_utility.cpp:390-733) and sit after the calendar bookkeeping of gtime_t: per satellite the caller supplies the signal
transmit time as day-of-year (time2doy, gnss_utility.cpp:299-306) and GPS seconds of week (time2gpst, :165-173).
The L1 selection (L1_freq, :933-962) is the caller's too: `freq` <= 0 marks "no L1 observation", which leaves the
satellite's rows zero exactly as the `continue` statements of the reference do.

Parity status: the reference has no test for gnss_comm ("parity unpinned" at the reference level); this module is
pinned by closed-form properties in tests/test_oracle_gnss_comm.py (geodetic round trip, zenith geometry,
Saastamoinen zenith delay, Klobuchar night-time floor).
"""
import math

import numpy as np

LIGHT_SPEED = 2.99792458e8          # gnss_constant.hpp:214
EARTH_OMG_GPS = 7.2921151467e-5     # gnss_constant.hpp:208
EARTH_ECCE_2 = 6.69437999014e-3     # gnss_constant.hpp:203
EARTH_SEMI_MAJOR = 6378137.0        # gnss_constant.hpp:205
D2R = math.pi / 180.0
R2D = 180.0 / math.pi


def geo2ecef(lla):
    """gnss_utility.cpp:335-345 (lat, lon in degrees, height in metres)."""
    cl, sl = math.cos(lla[0] * D2R), math.sin(lla[0] * D2R)
    N = EARTH_SEMI_MAJOR / math.sqrt(1 - EARTH_ECCE_2 * sl * sl)
    return np.array([(N + lla[2]) * cl * math.cos(lla[1] * D2R), (N + lla[2]) * cl * math.sin(lla[1] * D2R),
                     (N * (1 - EARTH_ECCE_2) + lla[2]) * sl])


def ecef2geo(xyz):
    """gnss_utility.cpp:347-387: closed-form (Bowring) geodetic latitude/longitude [deg] and height [m]."""
    x, y, z = float(xyz[0]), float(xyz[1]), float(xyz[2])
    if x == 0 and y == 0:
        return np.zeros(3)
    e2, a = EARTH_ECCE_2, EARTH_SEMI_MAJOR
    a2 = a * a
    b2 = a2 * (1 - e2)
    b = math.sqrt(b2)
    ep2 = (a2 - b2) / b2
    p = math.sqrt(x * x + y * y)
    s1, s2 = z * a, p * b
    h = math.sqrt(s1 * s1 + s2 * s2)
    sin_theta, cos_theta = s1 / h, s2 / h
    s1 = z + ep2 * b * sin_theta ** 3
    s2 = p - a * e2 * cos_theta ** 3
    h = math.sqrt(s1 * s1 + s2 * s2)
    tan_lat = s1 / s2
    sin_lat, cos_lat = s1 / h, s2 / h
    lat = math.atan(tan_lat)
    N = a2 * (a2 * cos_lat * cos_lat + b2 * sin_lat * sin_lat) ** -0.5
    alt = p / cos_lat - N
    lon = math.atan2(y, x)
    return np.array([lat * R2D, lon * R2D, alt])


def ecef2enu(ref_lla, v):
    """gnss_utility.cpp:722-735."""
    lat, lon = ref_lla[0] * D2R, ref_lla[1] * D2R
    sl, cl, so, co = math.sin(lat), math.cos(lat), math.sin(lon), math.cos(lon)
    R = np.array([[-so, co, 0.0], [-sl * co, -sl * so, cl], [cl * co, cl * so, sl]])
    return R @ np.asarray(v, float)


def sat_azel(rev_pos, sat_pos):
    """gnss_utility.cpp:762-771."""
    rev_pos, sat_pos = np.asarray(rev_pos, float), np.asarray(sat_pos, float)
    lla = ecef2geo(rev_pos)
    d = sat_pos - rev_pos
    u = d / np.linalg.norm(d)
    enu = ecef2enu(lla, u)
    az = 0.0 if math.sqrt(u[0] * u[0] + u[1] * u[1]) < 1e-12 else math.atan2(enu[0], enu[1])
    if az < 0:
        az += 2 * math.pi
    return az, math.asin(enu[2])


_NMF = np.array([
    [1.2769934E-3, 1.2683230E-3, 1.2465397E-3, 1.2196049E-3, 1.2045996E-3],
    [2.9153695E-3, 2.9152299E-3, 2.9288445E-3, 2.9022565E-3, 2.9024912E-3],
    [62.610505E-3, 62.837393E-3, 63.721774E-3, 63.824265E-3, 64.258455E-3],
    [0.0000000E-0, 1.2709626E-5, 2.6523662E-5, 3.4000452E-5, 4.1202191E-5],
    [0.0000000E-0, 2.1414979E-5, 3.0160779E-5, 7.2562722E-5, 11.723375E-5],
    [0.0000000E-0, 9.0128400E-5, 4.3497037E-5, 84.795348E-5, 170.37206E-5],
    [5.8021897E-4, 5.6794847E-4, 5.8118019E-4, 5.9727542E-4, 6.1641693E-4],
    [1.4275268E-3, 1.5138625E-3, 1.4572752E-3, 1.5007428E-3, 1.7599082E-3],
    [4.3472961E-2, 4.6729510E-2, 4.3908931E-2, 4.4626982E-2, 5.4736038E-2]])
_AHT = (2.53E-5, 5.49E-3, 1.14E-3)


def _interpc(coef, lat):
    """gnss_utility.cpp:774-779."""
    i = int(lat / 15.0)
    if i < 1:
        return coef[0]
    if i > 4:
        return coef[4]
    return coef[i - 1] * (1.0 - lat / 15.0 + i) + coef[i] * (lat / 15.0 - i)


def _mapf(el, a, b, c):
    """gnss_utility.cpp:782-786."""
    s = math.sin(el)
    return (1.0 + a / (1.0 + b / (1.0 + c))) / (s + (a / (s + b / (s + c))))


def nmf(doy, lla, azel):
    """Niell mapping functions, gnss_utility.cpp:798-839; `doy` = time2doy(ttx)."""
    el, lat, hgt = azel[1], lla[0], lla[2]
    if el <= 0.0:
        return 0.0, 0.0
    y = (doy - 28.0) / 365.25 + (0.5 if lat < 0.0 else 0.0)
    cosy = math.cos(2.0 * math.pi * y)
    lat = abs(lat)
    ah = [_interpc(_NMF[i], lat) - _interpc(_NMF[i + 3], lat) * cosy for i in range(3)]
    aw = [_interpc(_NMF[i + 6], lat) for i in range(3)]
    dm = (1.0 / math.sin(el) - _mapf(el, *_AHT)) * hgt / 1E3
    return _mapf(el, *ah) + dm, _mapf(el, *aw)


def calculate_trop_delay(doy, lla, azel):
    """Saastamoinen + standard atmosphere, gnss_utility.cpp:841-863."""
    temp0, humi = 15.0, 0.7
    if lla[2] < -100.0 or 1E4 < lla[2] or azel[1] <= 0:
        return 0.0
    hgt = 0.0 if lla[2] < 0.0 else lla[2]
    pres = 1013.25 * (1.0 - 2.2557E-5 * hgt) ** 5.2568
    temp = temp0 - 6.5E-3 * hgt + 273.16
    e = 6.108 * humi * math.exp((17.15 * temp - 4684.0) / (temp - 38.45))
    zhd = 0.0022768 * pres / (1.0 - 0.00266 * math.cos(2.0 * lla[0] * D2R) - 0.00028 * hgt / 1E3)
    zwd = 0.002277 * (1255.0 / temp + 0.05) * e
    mh, mw = nmf(doy, lla, azel)
    return mh * zhd + mw * zwd


def calculate_ion_delay(tow, ion, lla, azel):
    """Klobuchar, gnss_utility.cpp:865-901; `tow` = time2gpst(ttx) (seconds of GPS week)."""
    if ion is None or len(ion) == 0:
        return 0.0
    if lla[2] < -1E3 or azel[1] <= 0:
        return 0.0
    psi = 0.0137 / (azel[1] / math.pi + 0.11) - 0.022
    phi = lla[0] / 180.0 + psi * math.cos(azel[0])
    phi = 0.416 if phi > 0.416 else (-0.416 if phi < -0.416 else phi)
    lam = lla[1] / 180.0 + psi * math.sin(azel[0]) / math.cos(phi * math.pi)
    phi += 0.064 * math.cos((lam - 1.617) * math.pi)
    tt = 43200.0 * lam + tow
    tt -= math.floor(tt / 86400.0) * 86400.0
    f = 1.0 + 16.0 * (0.53 - azel[1] / math.pi) ** 3.0
    amp = ion[0] + phi * (ion[1] + phi * (ion[2] + phi * ion[3]))
    per = ion[4] + phi * (ion[5] + phi * (ion[6] + phi * ion[7]))
    amp = 0.0 if amp < 0.0 else amp
    per = 72000.0 if per < 72000.0 else per
    x = 2.0 * math.pi * (tt - 50400.0) / per
    return LIGHT_SPEED * f * (5E-9 + amp * (1.0 + x * x * (-0.5 + x * x / 24.0)) if abs(x) < 1.57 else 5E-9)


def psr_res(rcv_state, sat, iono):
    """gnss_spp.cpp:99-146. rcv_state = [ecef xyz, clock bias GPS, GLO, GAL, BDS]; `sat` holds per-satellite arrays
    pos (S,3), dt, tgd, sys (0..3), psr, freq, doy, tow. Returns res (S), J (S,7), atmos (S,2), azel (S,2)."""
    S = sat["pos"].shape[0]
    res, J = np.zeros(S), np.zeros((S, 7))
    atmos, azel_all = np.zeros((S, 2)), np.zeros((S, 2))
    rp = np.asarray(rcv_state[:3], float)
    for i in range(S):
        if not sat["freq"][i] > 0:     # L1_freq: l1_idx < 0 -> continue (rows and outputs stay zero)
            continue
        sv = sat["pos"][i]
        ion_d = tro_d = 0.0
        azel = (0.0, math.pi / 2.0)
        if np.linalg.norm(rp) > 0:
            azel = sat_azel(rp, sv)
            lla = ecef2geo(rp)
            tro_d = calculate_trop_delay(sat["doy"][i], lla, azel)
            ion_d = calculate_ion_delay(sat["tow"][i], iono, lla, azel)
        d = sv - rp
        rng = np.linalg.norm(d)
        unit = d / rng
        sagnac = EARTH_OMG_GPS * (sv[0] * rp[1] - sv[1] * rp[0]) / LIGHT_SPEED
        k = int(sat["sys"][i])
        est = rng + sagnac + rcv_state[3 + k] - sat["dt"][i] * LIGHT_SPEED + tro_d + ion_d + sat["tgd"][i] * LIGHT_SPEED
        J[i, :3] = -unit
        J[i, 3 + k] = 1.0
        res[i] = est - sat["psr"][i]
        atmos[i] = (ion_d, tro_d)
        azel_all[i] = azel
    return res, J, atmos, azel_all


def dopp_res(rcv_vel_state, rcv_ecef, sat):
    """gnss_spp.cpp:256-282. rcv_vel_state = [ecef velocity xyz, clock drift]."""
    S = sat["pos"].shape[0]
    res, J = np.zeros(S), np.zeros((S, 4))
    rp = np.asarray(rcv_ecef, float)
    rv = np.asarray(rcv_vel_state, float)
    for i in range(S):
        sv, vv = sat["pos"][i], sat["vel"][i]
        d = sv - rp
        unit = d / np.linalg.norm(d)
        sagnac = EARTH_OMG_GPS / LIGHT_SPEED * (vv[0] * rp[1] + sv[0] * rv[1] - vv[1] * rp[0] - sv[1] * rv[0])
        est = float((vv - rv[:3]) @ unit) + rv[3] + sagnac - sat["ddt"][i] * LIGHT_SPEED
        if not sat["freq"][i] > 0:
            continue
        wavelength = LIGHT_SPEED / sat["freq"][i]
        res[i] = est + sat["dopp"][i] * wavelength
        J[i, :3] = -unit
        J[i, 3] = 1.0
    return res, J


def receiver_states(p_w, v_w, yof, clock_bias4, fs, R_enu2ecef, t_enu2ecef):
    """GnssUpdate.cpp:101-111: xyzt (7) and the Doppler receiver state (4) from the filter mean."""
    c, s = math.cos(yof), math.sin(yof)
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    xyzt = np.zeros(7)
    xyzt[:3] = R_enu2ecef @ (Rz @ p_w) + t_enu2ecef
    xyzt[3:] = clock_bias4
    dopp = np.zeros(4)
    dopp[:3] = R_enu2ecef @ (Rz @ v_w)
    dopp[3] = fs
    return xyzt, dopp


def epoch_residuals(p_w, v_w, yof, clock_bias4, fs, R_enu2ecef, t_enu2ecef, sat, iono, psr_amp=1.0, dopp_amp=1.0):
    """Everything GnssUpdate::updateTrackedSys derives from gnss_comm for one epoch (GnssUpdate.cpp:101-122 and the
    noise model :177-186, :246-255): unit vectors, residuals, sigmas, az/el, delays."""
    xyzt, dv = receiver_states(p_w, v_w, yof, clock_bias4, fs, R_enu2ecef, t_enu2ecef)
    res_pos, Jp, atmos, azel = psr_res(xyzt, sat, iono)
    res_vel, Jv = dopp_res(dv, xyzt[:3], sat)
    sin_el = np.sin(azel[:, 1])
    sin_el = np.where(np.abs(sin_el) < 1e-6, 1e-6, sin_el)
    freq = np.where(sat["freq"] > 0, sat["freq"], 1.0)
    ndp = sat["dopp_std"] * LIGHT_SPEED / freq
    sig_psr = psr_amp * np.sqrt(sat["ura"] * sat["psr_std"] / (sin_el * sin_el))
    sig_dopp = dopp_amp * np.sqrt(sat["ura"] * ndp / (sin_el * sin_el))
    return dict(unit_psr=-Jp[:, :3], unit_dopp=-Jv[:, :3], res_pos=res_pos, res_vel=res_vel, sigma_psr=sig_psr,
                sigma_dopp=sig_dopp, azel=azel, atmos=atmos)


# ---- ephemeris -> satellite state (gnss_comm::sat_states, gnss_spp.cpp:50-97) ---------------------------------
# Times are carried as seconds RELATIVE TO THE EPHEMERIS REFERENCE EPOCH toe (time_diff(t, toe), exact in double for
# the few hours an ephemeris is valid); gtime_t calendar arithmetic stays with the caller.
MU_GPS = 3.9860050000e14            # gnss_constant.hpp:210
MU = 3.9860044180e14                # gnss_constant.hpp:211
EARTH_OMG_GLO = 7.2921150000e-5     # gnss_constant.hpp:207
EARTH_OMG_BDS = 7.2921150000e-5     # gnss_constant.hpp:209
TSTEP = 60.0                        # gnss_constant.hpp:212
J2_GLO = 1.0826257E-3               # gnss_constant.hpp:213
EARTH_SEMI_MAJOR_GLO = 6378136.0    # gnss_constant.hpp:206
WEEK_SECONDS = 604800               # gnss_constant.hpp:217
SIN_N5, COS_N5 = -0.0871557427476582, 0.9961946980917456   # gnss_constant.hpp:226-227
SYS_GPS, SYS_GLO, SYS_GAL, SYS_BDS = 0, 1, 2, 3             # sys2idx numbering (gnss_constant.hpp:264-270)
KEPLER_FIELDS = ("A", "e", "i0", "OMG0", "omg", "M0", "delta_n", "OMG_dot", "i_dot", "cuc", "cus", "crc", "crs", "cic",
                 "cis", "af0", "af1", "af2", "toe_tow", "tgd", "toe_minus_toc", "prn")
GLO_FIELDS = ("px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az", "tau_n", "gamma")


def kepler(mk, es):
    """gnss_utility.cpp:390-405 -- note that the reference returns the PREVIOUS iterate `ek`."""
    e, ek, it = mk, 1e6, 0
    while it < 30 and abs(e - ek) > 1e-14:
        ek = e
        e -= (e - es * math.sin(e) - mk) / (1.0 - es * math.cos(e))
        it += 1
    return ek


def _wrap_week(t):
    if t > WEEK_SECONDS / 2:
        return t - WEEK_SECONDS
    if t < -WEEK_SECONDS / 2:
        return t + WEEK_SECONDS
    return t


def _mu_omg(sys):
    if sys == SYS_GPS:
        return MU_GPS, EARTH_OMG_GPS
    if sys == SYS_GLO:
        return MU, EARTH_OMG_GLO
    if sys == SYS_BDS:
        return MU, EARTH_OMG_BDS
    return MU, EARTH_OMG_GPS


def eph2svdt(t_rel, eph):
    """gnss_utility.cpp:437-446; t_rel = time_diff(t, toe)."""
    dt = t_rel + eph["toe_minus_toc"]
    for _ in range(2):
        dt -= eph["af0"] + eph["af1"] * dt + eph["af2"] * dt * dt
    return eph["af0"] + eph["af1"] * dt + eph["af2"] * dt * dt


def _kepler_plane(t_rel, eph, sys):
    tk = _wrap_week(t_rel)
    mu, omg_e = _mu_omg(sys)
    n = math.sqrt(mu / eph["A"] ** 3) + eph["delta_n"]
    Ek = kepler(eph["M0"] + n * tk, eph["e"])
    return tk, mu, omg_e, n, Ek


def eph2pos(t_rel, eph, sys):
    """gnss_utility.cpp:448-531 -> (position ECEF, svdt)."""
    tk, mu, omg_e, n, Ek = _kepler_plane(t_rel, eph, sys)
    sE, cE = math.sin(Ek), math.cos(Ek)
    e = eph["e"]
    vk = math.atan2(math.sqrt(1 - e * e) * sE, cE - e)
    phi = vk + eph["omg"]
    c2, s2 = math.cos(2 * phi), math.sin(2 * phi)
    uk = phi + eph["cus"] * s2 + eph["cuc"] * c2
    rk = eph["A"] * (1 - e * cE) + eph["crs"] * s2 + eph["crc"] * c2
    ik = eph["i0"] + eph["i_dot"] * tk + eph["cis"] * s2 + eph["cic"] * c2
    si, ci = math.sin(ik), math.cos(ik)
    xp, yp = rk * math.cos(uk), rk * math.sin(uk)
    if sys == SYS_BDS and eph["prn"] <= 5:      # BDS GEO
        O = eph["OMG0"] + eph["OMG_dot"] * tk - omg_e * eph["toe_tow"]
        sO, cO = math.sin(O), math.cos(O)
        xg, yg, zg = xp * cO - yp * ci * sO, xp * sO + yp * ci * cO, yp * si
        so, co = math.sin(omg_e * tk), math.cos(omg_e * tk)
        pos = np.array([xg * co + yg * so * COS_N5 + zg * so * SIN_N5, -xg * so + yg * co * COS_N5 + zg * co * SIN_N5,
                        -yg * SIN_N5 + zg * COS_N5])
    else:
        O = eph["OMG0"] + (eph["OMG_dot"] - omg_e) * tk - omg_e * eph["toe_tow"]
        sO, cO = math.sin(O), math.cos(O)
        pos = np.array([xp * cO - yp * ci * sO, xp * sO + yp * ci * cO, yp * si])
    dt = t_rel + eph["toe_minus_toc"]
    dts = eph["af0"] + eph["af1"] * dt + eph["af2"] * dt * dt
    dts -= 2.0 * math.sqrt(mu * eph["A"]) * e * sE / LIGHT_SPEED / LIGHT_SPEED
    return pos, dts


def eph2vel(t_rel, eph, sys):
    """gnss_utility.cpp:533-634 -> (velocity ECEF, svddt)."""
    tk, mu, omg_e, n, Ek = _kepler_plane(t_rel, eph, sys)
    sE, cE = math.sin(Ek), math.cos(Ek)
    e = eph["e"]
    Ed = n / (1 - e * cE)
    vd = math.sqrt(1 - e * e) * Ed / (1 - e * cE)
    vk = math.atan2(math.sqrt(1 - e * e) * sE, cE - e)
    phi = vk + eph["omg"]
    c2, s2 = math.cos(2 * phi), math.sin(2 * phi)
    dud = 2 * vd * (eph["cus"] * c2 - eph["cuc"] * s2)
    drd = 2 * vd * (eph["crs"] * c2 - eph["crc"] * s2)
    did = 2 * vd * (eph["cis"] * c2 - eph["cic"] * s2)
    ukd, rkd, ikd = vd + dud, eph["A"] * e * Ed * sE + drd, eph["i_dot"] + did
    uk = phi + eph["cus"] * s2 + eph["cuc"] * c2
    rk = eph["A"] * (1 - e * cE) + eph["crs"] * s2 + eph["crc"] * c2
    ik = eph["i0"] + eph["i_dot"] * tk + eph["cis"] * s2 + eph["cic"] * c2
    si, ci = math.sin(ik), math.cos(ik)
    su, cu = math.sin(uk), math.cos(uk)
    xp, yp = rk * cu, rk * su
    xpd, ypd = rkd * cu - rk * ukd * su, rkd * su + rk * ukd * cu
    if sys == SYS_BDS and eph["prn"] <= 5:
        O = eph["OMG0"] + eph["OMG_dot"] * tk - omg_e * eph["toe_tow"]
        sO, cO = math.sin(O), math.cos(O)
        Od = eph["OMG_dot"]
        t1 = xpd - yp * Od * ci
        t2 = xp * Od + ypd * ci - yp * ikd * si
        xg, yg, zg = xp * cO - yp * ci * sO, xp * sO + yp * ci * cO, yp * si
        xgd, ygd = t1 * cO - t2 * sO, t1 * sO + t2 * cO
        zgd = ypd * si + ypd * ikd * ci          # as the reference writes it (:606)
        so, co = math.sin(omg_e * tk), math.cos(omg_e * tk)
        sod, cod = omg_e * co, -omg_e * so
        vel = np.array([xgd * co + xg * cod + ygd * so * COS_N5 + yg * sod * COS_N5 + zgd * so * SIN_N5 + zg * sod * SIN_N5,
                        -xgd * so - xg * sod + ygd * co * COS_N5 + yg * cod * COS_N5 + zgd * co * SIN_N5 + zg * cod * SIN_N5,
                        -ygd * SIN_N5 + zgd * COS_N5])
    else:
        O = eph["OMG0"] + (eph["OMG_dot"] - omg_e) * tk - omg_e * eph["toe_tow"]
        sO, cO = math.sin(O), math.cos(O)
        Od = eph["OMG_dot"] - omg_e
        t1 = xpd - yp * Od * ci
        t2 = xp * Od + ypd * ci - yp * ikd * si
        vel = np.array([t1 * cO - t2 * sO, t1 * sO + t2 * cO, ypd * si + ypd * ikd * ci])   # z as the reference (:624)
    dt = t_rel + eph["toe_minus_toc"]
    ddts = eph["af1"] + 2.0 * eph["af2"] * dt
    ddts -= 2.0 * math.sqrt(mu * eph["A"]) * e * cE * Ed / LIGHT_SPEED / LIGHT_SPEED
    return vel, ddts


def _glo_deq(pos, vel, acc):
    """gnss_utility.cpp:636-662."""
    r2 = float(pos @ pos)
    if r2 <= 0.0:
        return np.zeros(3), np.zeros(3)
    r3 = r2 * math.sqrt(r2)
    omg2 = EARTH_OMG_GLO * EARTH_OMG_GLO
    a = 1.5 * J2_GLO * MU * EARTH_SEMI_MAJOR_GLO * EARTH_SEMI_MAJOR_GLO / r2 / r3
    b = 5.0 * pos[2] * pos[2] / r2
    c = -MU / r3 - a * (1.0 - b)
    vd = np.array([(c + omg2) * pos[0] + 2.0 * EARTH_OMG_GLO * vel[1] + acc[0],
                   (c + omg2) * pos[1] - 2.0 * EARTH_OMG_GLO * vel[0] + acc[1],
                   (c - 2.0 * a) * pos[2] + acc[2]])
    return vel.copy(), vd


def _glo_orbit(dt, pos, vel, acc):
    """RK4 step, gnss_utility.cpp:664-682."""
    p1, v1 = _glo_deq(pos, vel, acc)
    p2, v2 = _glo_deq(pos + 0.5 * p1 * dt, vel + 0.5 * v1 * dt, acc)
    p3, v3 = _glo_deq(pos + 0.5 * p2 * dt, vel + 0.5 * v2 * dt, acc)
    p4, v4 = _glo_deq(pos + p3 * dt, vel + v3 * dt, acc)
    return pos + (p1 + 2.0 * p2 + 2.0 * p3 + p4) * dt / 6.0, vel + (v1 + 2.0 * v2 + 2.0 * v3 + v4) * dt / 6.0


def geph2svdt(t_rel, g):
    """gnss_utility.cpp:684-696."""
    dt = _wrap_week(t_rel)
    for _ in range(2):
        dt -= -g["tau_n"] + g["gamma"] * dt
    return -g["tau_n"] + g["gamma"] * dt


def geph2posvel(t_rel, g):
    """geph2pos / geph2vel (gnss_utility.cpp:698-733): the same integration, run once here."""
    pos = np.array([g["px"], g["py"], g["pz"]], float)
    vel = np.array([g["vx"], g["vy"], g["vz"]], float)
    acc = np.array([g["ax"], g["ay"], g["az"]], float)
    dt = t_rel
    dts = -g["tau_n"] + g["gamma"] * dt
    tt = -TSTEP if dt < 0.0 else TSTEP
    while abs(dt) > 1e-9:
        if abs(dt) < TSTEP:
            tt = dt
        pos, vel = _glo_orbit(tt, pos, vel, acc)
        dt -= tt
    return pos, vel, dts, g["gamma"]


def sat_state(t_obs_rel, psr, sys, eph):
    """One satellite of gnss_comm::sat_states (gnss_spp.cpp:57-94). t_obs_rel = time_diff(obs->time, toe); psr <= 0
    stands for "no L1 observation" (the reference leaves the default-constructed, all-zero SatState).
    Returns dict(pos, vel, dt, ddt, tgd, ttx_rel)."""
    out = dict(pos=np.zeros(3), vel=np.zeros(3), dt=0.0, ddt=0.0, tgd=0.0, ttx_rel=0.0)
    if not psr > 0:
        return out
    tx = t_obs_rel - psr / LIGHT_SPEED
    if sys == SYS_GLO:
        tx -= geph2svdt(tx, eph)
        pos, vel, dts, ddts = geph2posvel(tx, eph)
        tgd = 0.0
    else:
        tx -= eph2svdt(tx, eph)
        pos, dts = eph2pos(tx, eph, sys)
        vel, ddts = eph2vel(tx, eph, sys)
        tgd = eph["tgd"]
    out.update(pos=pos, vel=vel, dt=dts, ddt=ddts, tgd=tgd, ttx_rel=tx)
    return out
