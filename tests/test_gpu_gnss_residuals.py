"""igv_gnss_residuals (gnss_comm::psr_res / dopp_res on the device) against the oracle restatement, and chained
into igv_gnss_update."""
import numpy as np
import pytest

import ingvio_oracle.gnss_comm as gc
from helpers import assert_state_close, filter_params
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF
from ingvio_b200.synth import WORKLOADS, enu2ecef_rotation, geo2ecef, raw_gnss_epoch

import test_gpu_parity as tp

pytestmark = pytest.mark.gpu


def _receiver(f, Re, t0):
    st = f.state
    e = st.extended_pose
    cb = np.array([st.gnss[g].value() if g in st.gnss else 0.0 for g in (GPS, GLO, GAL, BDS)])
    fs = st.gnss[FS].value() if FS in st.gnss else 0.0
    yof = st.gnss[YOF].value()
    return e.vec1.copy(), e.vec2.copy(), yof, cb, fs


@pytest.mark.parametrize("lat,lon", [(22.3, 114.2), (-33.9, 151.2)])
def test_residuals_match_oracle(lat, lon):
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    B, S = 3, 10
    _, st, orc, g = tp._warm(wl, B, 5, fp=fp, max_sats=12)
    rng = np.random.default_rng(11)
    Re = enu2ecef_rotation(lat, lon)
    t0 = geo2ecef(lat, lon, 40.0)
    T = np.tile(np.concatenate([Re.reshape(9), t0]), (B, 1))
    rcv, rcvv, cbs, fss = [], [], [], []
    for f in orc:
        p, v, yof, cb, fs = _receiver(f, Re, t0)
        xyzt, dv = gc.receiver_states(p, v, yof, cb, fs, Re, t0)
        rcv.append(xyzt[:3]); rcvv.append(dv[:3]); cbs.append(cb); fss.append(fs)
    raw = raw_gnss_epoch(rng, np.array(rcv), np.array(rcvv), np.array(cbs), np.array(fss), S, lat, lon,
                         no_l1=(3,), below_horizon=(5,))
    out = g.gnss_residuals(raw["sat_pos"], raw["sat_vel"], raw["sat_clk"], raw["obs"], raw["obs_std"], raw["ttx"], raw["sys"],
                           T, raw["iono"], psr_amp=1.3, dopp_amp=0.7)
    for b, f in enumerate(orc):
        p, v, yof, cb, fs = _receiver(f, Re, t0)
        sat = dict(pos=raw["sat_pos"][b], vel=raw["sat_vel"][b], dt=raw["sat_clk"][b, :, 0], ddt=raw["sat_clk"][b, :, 1],
                   tgd=raw["sat_clk"][b, :, 2], sys=raw["sys"][b], psr=raw["obs"][b, :, 0], dopp=raw["obs"][b, :, 1],
                   freq=raw["obs"][b, :, 2], doy=raw["ttx"][b, :, 0], tow=raw["ttx"][b, :, 1], ura=raw["obs_std"][b, :, 0],
                   psr_std=raw["obs_std"][b, :, 1], dopp_std=raw["obs_std"][b, :, 2])
        ref = gc.epoch_residuals(p, v, yof, cb, fs, Re, t0, sat, raw["iono"][b], psr_amp=1.3, dopp_amp=0.7)
        # ranges are 2e7 m: 1e-16 relative is 2e-9 m; the bars are absolute [m], [m/s], [rad]
        assert np.abs(out["res_pos"][b] - ref["res_pos"]).max() < 1e-6
        assert np.abs(out["res_vel"][b] - ref["res_vel"]).max() < 1e-8
        assert np.abs(out["unit"][b] - ref["unit_psr"]).max() < 1e-13
        assert np.abs(out["azel"][b] - ref["azel"]).max() < 1e-11
        assert np.abs(out["atmos"][b] - ref["atmos"]).max() < 1e-8
        assert np.allclose(out["sigma_psr"][b], ref["sigma_psr"], rtol=1e-12, atol=0)
        assert np.allclose(out["sigma_dopp"][b], ref["sigma_dopp"], rtol=1e-12, atol=0)
        # the cases the reference treats specially
        assert out["res_pos"][b, 3] == 0.0 and np.all(out["unit"][b, 3] == 0.0)            # no L1 observation
        assert out["azel"][b, 5, 1] < 0 and np.all(out["atmos"][b, 5] == 0.0)              # below the horizon
        assert np.abs(ref["res_pos"][[0, 1, 2, 4]]).max() < 40.0                            # a sane epoch: metres


def test_residuals_chain_into_gnss_update():
    """Raw epoch -> igv_gnss_residuals -> igv_gnss_update equals the oracle's update fed with the oracle's residuals."""
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    B, S = 2, 8
    _, st, orc, g = tp._warm(wl, B, 5, fp=fp, max_sats=12)
    rng = np.random.default_rng(5)
    lat, lon = 22.3, 114.2
    Re = enu2ecef_rotation(lat, lon)
    t0 = geo2ecef(lat, lon, 40.0)
    T = np.tile(np.concatenate([Re.reshape(9), t0]), (B, 1))
    rs = [_receiver(f, Re, t0) for f in orc]
    xs = [gc.receiver_states(p, v, yof, cb, fs, Re, t0) for p, v, yof, cb, fs in rs]
    raw = raw_gnss_epoch(rng, np.array([x[0][:3] for x in xs]), np.array([x[1][:3] for x in xs]),
                         np.array([r[3] for r in rs]), np.array([r[4] for r in rs]), S, lat, lon, el_range=(20.0, 85.0))
    out = g.gnss_residuals(raw["sat_pos"], raw["sat_vel"], raw["sat_clk"], raw["obs"], raw["obs_std"], raw["ttx"], raw["sys"],
                           T, raw["iono"])
    g.gnss_update(out["unit"], out["res_pos"], out["res_vel"], out["sigma_psr"], out["sigma_dopp"], raw["sys"],
                  np.tile(Re.reshape(1, 9), (B, 1)), 0, 0, 1)
    from ingvio_oracle.gnss_update import GnssEpoch
    for b, f in enumerate(orc):
        p, v, yof, cb, fs = rs[b]
        sat = dict(pos=raw["sat_pos"][b], vel=raw["sat_vel"][b], dt=raw["sat_clk"][b, :, 0], ddt=raw["sat_clk"][b, :, 1],
                   tgd=raw["sat_clk"][b, :, 2], sys=raw["sys"][b], psr=raw["obs"][b, :, 0], dopp=raw["obs"][b, :, 1],
                   freq=raw["obs"][b, :, 2], doy=raw["ttx"][b, :, 0], tow=raw["ttx"][b, :, 1], ura=raw["obs_std"][b, :, 0],
                   psr_std=raw["obs_std"][b, :, 1], dopp_std=raw["obs_std"][b, :, 2])
        ref = gc.epoch_residuals(p, v, yof, cb, fs, Re, t0, sat, raw["iono"][b])
        ep = GnssEpoch(unit=ref["unit_psr"], res_pos=ref["res_pos"], res_vel=ref["res_vel"], sys=raw["sys"][b],
                       ura=sat["ura"], psr_std=sat["psr_std"], dopp_std_mps=sat["dopp_std"] * gc.LIGHT_SPEED / sat["freq"],
                       el=ref["azel"][:, 1])
        f.gnss.update_tracked_sys(f.state, ep, Re)
    assert_state_close(g, orc, wl.sw, what="raw epoch -> residuals -> gnss update")


def test_sat_states_match_oracle():
    """igv_sat_states (gnss_comm::sat_states on the device) for Kepler (GPS, GAL, BDS MEO and GEO) and GLONASS records."""
    from ingvio_b200.synth import random_ephemerides
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    B, S = 2, 12
    _, st, orc, g = tp._warm(wl, B, 2, fp=fp, max_sats=12)
    rng = np.random.default_rng(21)
    eph, sys_, t_obs, psr = random_ephemerides(rng, B, S, geo=(7,))
    psr[:, 4] = 0.0                                   # no L1 observation
    out = g.sat_states(eph, t_obs, psr, sys_)
    for b in range(B):
        for i in range(S):
            k = int(sys_[b, i])
            rec = dict(zip(gc.GLO_FIELDS if k == gc.SYS_GLO else gc.KEPLER_FIELDS, eph[b, i]))
            ref = gc.sat_state(t_obs[b, i], psr[b, i], k, rec)
            assert np.abs(out["sat_pos"][b, i] - ref["pos"]).max() < 1e-5, (b, i, k)
            assert np.abs(out["sat_vel"][b, i] - ref["vel"]).max() < 1e-8, (b, i, k)
            assert abs(out["sat_clk"][b, i, 0] - ref["dt"]) < 1e-16 and abs(out["sat_clk"][b, i, 1] - ref["ddt"]) < 1e-20
            assert out["sat_clk"][b, i, 2] == ref["tgd"]
            assert abs(out["ttx_rel"][b, i] - ref["ttx_rel"]) < 1e-12
    assert np.all(out["sat_pos"][:, 4] == 0.0) and np.all(out["sat_clk"][:, 4] == 0.0)
    r = np.linalg.norm(out["sat_pos"][:, [0, 1, 2, 3, 7]], axis=-1)
    assert np.all(r > 2.4e7) and np.all(r < 4.3e7)
