"""Closed-form pins of the oracle's gnss_comm restatement (the reference has no tests for gnss_comm)."""
import math

import numpy as np
import pytest

import ingvio_oracle.gnss_comm as gc


@pytest.mark.parametrize("lla", [(22.3, 114.2, 35.0), (-33.9, 151.2, 120.0), (78.2, 15.6, 500.0), (0.01, -60.0, 3000.0)])
def test_geodetic_round_trip(lla):
    xyz = gc.geo2ecef(np.array(lla))
    back = gc.ecef2geo(xyz)
    assert abs(back[0] - lla[0]) < 1e-9 and abs(back[1] - lla[1]) < 1e-9 and abs(back[2] - lla[2]) < 1e-4


def test_ecef2geo_undefined_on_the_axis():
    assert np.all(gc.ecef2geo(np.array([0.0, 0.0, 6.4e6])) == 0.0)   # gnss_utility.cpp:350-354


def test_zenith_and_compass_geometry():
    lla = np.array([22.3, 114.2, 35.0])
    rcv = gc.geo2ecef(lla)
    up = gc.geo2ecef(lla + np.array([0, 0, 2.0e7]))
    az, el = gc.sat_azel(rcv, up)
    assert abs(el - math.pi / 2) < 1e-6
    # a point 1000 km to the geodetic north-east, 20 000 km up: azimuth in the first quadrant, elevation < 90 deg
    ne = gc.geo2ecef(lla + np.array([5.0, 5.0, 2.0e7]))
    az, el = gc.sat_azel(rcv, ne)
    assert 0.0 < az < math.pi / 2 and 0.0 < el < math.pi / 2
    # ecef2enu is a rotation
    v = np.array([0.3, -1.2, 2.5])
    assert abs(np.linalg.norm(gc.ecef2enu(lla, v)) - np.linalg.norm(v)) < 1e-12


def test_saastamoinen_zenith_delay_and_mapping():
    lla = np.array([45.0, 10.0, 0.0])
    z = gc.calculate_trop_delay(180.0, lla, (0.0, math.pi / 2))
    assert 2.3 < z < 2.7                      # hydrostatic ~2.3 m + wet ~0.1-0.3 m at sea level
    low = gc.calculate_trop_delay(180.0, lla, (0.0, math.radians(10.0)))
    assert 5.0 < low / z < 6.0                # mapping function ~ 1/sin(el) = 5.76 at 10 degrees
    assert gc.calculate_trop_delay(180.0, lla, (0.0, -0.1)) == 0.0
    assert gc.calculate_trop_delay(180.0, np.array([45.0, 10.0, 2.0e4]), (0.0, 1.0)) == 0.0
    mh, mw = gc.nmf(180.0, lla, (0.0, math.pi / 2))
    assert abs(mh - 1.0) < 1e-3 and abs(mw - 1.0) < 1e-3


def test_klobuchar_night_floor_and_day_peak():
    ion = [0.1118e-7, -0.7451e-8, -0.5961e-7, 0.1192e-6, 0.1167e6, -0.2294e6, -0.1311e6, 0.1049e7]
    lla = np.array([30.0, 0.0, 100.0])
    azel = (0.0, math.pi / 2)
    night = gc.calculate_ion_delay(2 * 3600.0, ion, lla, azel)    # 02:00 local time at lon 0
    assert abs(night - gc.LIGHT_SPEED * 5e-9 * (1.0 + 16.0 * 0.03 ** 3)) < 1e-9   # slant factor at zenith
    day = gc.calculate_ion_delay(14 * 3600.0, ion, lla, azel)     # 14:00 local: the cosine peaks
    assert day > night and day < 30.0
    assert gc.calculate_ion_delay(14 * 3600.0, [], lla, azel) == 0.0
    assert gc.calculate_ion_delay(14 * 3600.0, ion, lla, (0.0, -0.2)) == 0.0


def test_psr_and_dopp_residuals_vanish_for_consistent_measurements():
    rng = np.random.default_rng(3)
    lla = np.array([22.3, 114.2, 35.0])
    rcv = gc.geo2ecef(lla)
    S = 6
    pos = rcv + rng.standard_normal((S, 3)) * 1e6 + gc.geo2ecef(lla + np.array([0, 0, 2.0e7])) - rcv
    vel = rng.standard_normal((S, 3)) * 1e3
    sys = np.arange(S) % 4
    freq = np.full(S, 1575.42e6)
    sat = dict(pos=pos, vel=vel, dt=rng.normal(0, 1e-5, S), ddt=rng.normal(0, 1e-11, S), tgd=rng.normal(0, 1e-8, S), sys=sys,
               psr=np.zeros(S), dopp=np.zeros(S), freq=freq, doy=np.full(S, 100.0), tow=np.full(S, 3e5),
               ura=np.ones(S), psr_std=np.ones(S), dopp_std=np.ones(S))
    xyzt = np.concatenate([rcv, [10.0, -20.0, 30.0, 5.0]])
    ION = [0.1118e-7, -0.7451e-8, -0.5961e-7, 0.1192e-6, 0.1167e6, -0.2294e6, -0.1311e6, 0.1049e7]
    res, J, atmos, azel = gc.psr_res(xyzt, sat, ION)
    sat["psr"] = res.copy()                       # measured := estimated
    res2, _, _, _ = gc.psr_res(xyzt, sat, ION)
    assert np.abs(res2).max() < 1e-6
    assert np.allclose(np.linalg.norm(J[:, :3], axis=1), 1.0) and np.all(J[np.arange(S), 3 + sys] == 1.0)
    # numerical derivative of the estimated range with respect to the receiver position equals J (Sagnac and the
    # atmosphere move by ~1e-4 per metre: the hydrostatic delay follows the height)
    eps = 1.0
    for ax in range(3):
        x2 = xyzt.copy(); x2[ax] += eps
        r3, _, _, _ = gc.psr_res(x2, sat, ION)
        assert np.abs((r3 - res2) / eps - J[:, ax]).max() < 5e-4
    rv = np.array([1.0, -2.0, 0.5, 0.3])
    dres, Jv = gc.dopp_res(rv, rcv, sat)
    sat["dopp"] = -dres * freq / gc.LIGHT_SPEED
    dres2, _ = gc.dopp_res(rv, rcv, sat)
    assert np.abs(dres2).max() < 1e-9
    # satellites without L1 keep zero rows
    sat["freq"][2] = -1.0
    res4, J4, _, _ = gc.psr_res(xyzt, sat, ION)
    d4, Jv4 = gc.dopp_res(rv, rcv, sat)
    assert res4[2] == 0.0 and np.all(J4[2] == 0.0) and d4[2] == 0.0 and np.all(Jv4[2] == 0.0)


# ---- ephemeris -> satellite state ------------------------------------------------------------------------------
def _circular(A=26.56e6, sys=0):
    return dict(A=A, e=0.0, i0=0.96, OMG0=0.7, omg=0.0, M0=0.4, delta_n=0.0, OMG_dot=0.0, i_dot=0.0, cuc=0.0, cus=0.0,
                crc=0.0, crs=0.0, cic=0.0, cis=0.0, af0=0.0, af1=0.0, af2=0.0, toe_tow=1.0e5, tgd=0.0, toe_minus_toc=0.0, prn=9)


def test_kepler_returns_the_previous_iterate_and_solves_the_equation():
    for mk, es in ((0.3, 0.01), (2.5, 0.02), (-1.0, 0.3)):
        E = gc.kepler(mk, es)
        assert abs(E - es * math.sin(E) - mk) < 1e-12


def test_circular_orbit_radius_period_and_velocity():
    eph = _circular()
    mu, omg_e = gc.MU_GPS, gc.EARTH_OMG_GPS
    n = math.sqrt(mu / eph["A"] ** 3)
    p0, _ = gc.eph2pos(50.0, eph, 0)
    assert abs(np.linalg.norm(p0) - eph["A"]) < 1e-6
    T = 2 * math.pi / n
    p1, _ = gc.eph2pos(50.0 + T, eph, 0)           # one revolution later the Earth has turned by omg_e * T
    c, s = math.cos(omg_e * T), math.sin(omg_e * T)
    Rz = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])
    assert np.abs(p1 - Rz @ p0).max() < 1e-4
    v, _ = gc.eph2vel(50.0, eph, 0)
    h = 1e-3
    pm, _ = gc.eph2pos(50.0 - h, eph, 0)
    pp, _ = gc.eph2pos(50.0 + h, eph, 0)
    assert np.abs(v - (pp - pm) / (2 * h)).max() < 1e-5     # with i_dot = cis = cic = 0 the reference's z term is exact


def test_eccentric_orbit_velocity_matches_the_position_derivative_in_the_plane_terms():
    eph = _circular()
    eph.update(e=0.015, omg=0.8, delta_n=4e-9, OMG_dot=-8e-9, cuc=2e-6, cus=4e-6, crc=250.0, crs=40.0, af0=1e-4, af1=1e-11)
    v, ddt = gc.eph2vel(300.0, eph, 2)
    h = 1e-3
    pm, dm = gc.eph2pos(300.0 - h, eph, 2)
    pp, dp = gc.eph2pos(300.0 + h, eph, 2)
    assert np.abs(v - (pp - pm) / (2 * h)).max() < 1e-4
    assert abs(ddt - (dp - dm) / (2 * h)) < 1e-13
    r, _ = gc.eph2pos(300.0, eph, 2)
    assert 0.98 * eph["A"] < np.linalg.norm(r) < 1.02 * eph["A"]


def test_bds_geo_branch_is_a_rotation_of_the_inclined_frame():
    eph = _circular(A=42.164e6)
    eph.update(i0=0.09, prn=3)
    p, _ = gc.eph2pos(120.0, eph, gc.SYS_BDS)
    assert abs(np.linalg.norm(p) - eph["A"]) < 1e-6        # the -5 degree tilt and the Earth rotation preserve length
    v, _ = gc.eph2vel(120.0, eph, gc.SYS_BDS)
    h = 1e-3
    pm, _ = gc.eph2pos(120.0 - h, eph, gc.SYS_BDS)
    pp, _ = gc.eph2pos(120.0 + h, eph, gc.SYS_BDS)
    assert np.abs(v - (pp - pm) / (2 * h)).max() < 1e-4


def test_glonass_integration_round_trip_and_clock():
    g = dict(px=1.2e7, py=-1.5e7, pz=1.6e7, vx=1200.0, vy=2500.0, vz=1500.0, ax=1e-6, ay=-1e-6, az=0.0, tau_n=1e-5, gamma=1e-12)
    p1, v1, dts, ddts = gc.geph2posvel(400.0, g)
    g2 = dict(g, px=p1[0], py=p1[1], pz=p1[2], vx=v1[0], vy=v1[1], vz=v1[2])
    p0, v0, _, _ = gc.geph2posvel(-400.0, g2)
    assert np.abs(p0 - [g["px"], g["py"], g["pz"]]).max() < 1e-3 and np.abs(v0 - [g["vx"], g["vy"], g["vz"]]).max() < 1e-6
    assert abs(dts - (-1e-5 + 1e-12 * 400.0)) < 1e-18 and ddts == 1e-12
    # 60 s RK4 steps against 1 s steps
    pf, vf = np.array([g["px"], g["py"], g["pz"]]), np.array([g["vx"], g["vy"], g["vz"]])
    acc = np.array([g["ax"], g["ay"], g["az"]])
    for _ in range(400):
        pf, vf = gc._glo_orbit(1.0, pf, vf, acc)
    assert np.abs(pf - p1).max() < 1e-3


def test_sat_state_transmit_time_and_missing_l1():
    eph = _circular()
    eph.update(af0=2e-4, af1=1e-11, tgd=4e-9)
    s = gc.sat_state(100.0, 2.3e7, 0, eph)
    tof = 2.3e7 / gc.LIGHT_SPEED
    assert abs(s["ttx_rel"] - (100.0 - tof - 2e-4)) < 1e-9 and s["tgd"] == 4e-9
    p, dts = gc.eph2pos(s["ttx_rel"], eph, 0)
    assert np.all(s["pos"] == p) and s["dt"] == dts
    z = gc.sat_state(100.0, 0.0, 0, eph)
    assert np.all(z["pos"] == 0) and z["dt"] == 0.0 and z["ttx_rel"] == 0.0
