"""k_sat_states / k_gnss_residuals (ingvio_b200/csrc/k_gnss_res.cu: gnss_comm::sat_states, psr_res, dopp_res on the device)
executed on the CPU through tests/emul against the oracle restatement -- the CPU twins of tests/test_gpu_gnss_residuals.py,
same inputs, same bars."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ingvio_oracle.gnss_comm as gc
from helpers import filter_params, make_oracles, oracle_packed_state
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF
from ingvio_b200.synth import WORKLOADS, SyntheticStream, enu2ecef_rotation, geo2ecef, random_ephemerides, raw_gnss_epoch

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul")


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(EMUL, "_build", "libgnss_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I" + cuda_inc, "-I" + EMUL,
                        os.path.join(EMUL, "gnss_emul.cpp"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    L = C.CDLL(out)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.emu_sat_states.argtypes = [i, i] + [vp] * 8
    L.emu_gnss_residuals.argtypes = [i, i, vp, i] + [vp] * 10 + [d, d] + [vp] * 7
    return L


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_sat_states_match_oracle(lib):
    B, S = 2, 12
    rng = np.random.default_rng(21)
    eph, sys_, t_obs, psr = random_ephemerides(rng, B, S, geo=(7,))
    psr[:, 4] = 0.0                                   # no L1 observation
    eph, t_obs, psr = (np.ascontiguousarray(a, np.float64) for a in (eph, t_obs, psr))
    sys_ = np.ascontiguousarray(sys_, np.int32)
    pos, vel, clk, ttx = np.zeros((B, S, 3)), np.zeros((B, S, 3)), np.zeros((B, S, 3)), np.zeros((B, S))
    lib.emu_sat_states(B, S, _p(eph), _p(t_obs), _p(psr), _p(sys_), _p(pos), _p(vel), _p(clk), _p(ttx))
    for b in range(B):
        for k in range(S):
            s = int(sys_[b, k])
            rec = dict(zip(gc.GLO_FIELDS if s == gc.SYS_GLO else gc.KEPLER_FIELDS, eph[b, k]))
            ref = gc.sat_state(t_obs[b, k], psr[b, k], s, rec)
            assert np.abs(pos[b, k] - ref["pos"]).max() < 1e-5, (b, k, s)
            assert np.abs(vel[b, k] - ref["vel"]).max() < 1e-8, (b, k, s)
            assert abs(clk[b, k, 0] - ref["dt"]) < 1e-16 and abs(clk[b, k, 1] - ref["ddt"]) < 1e-20 and clk[b, k, 2] == ref["tgd"]
            assert abs(ttx[b, k] - ref["ttx_rel"]) < 1e-12
    assert np.all(pos[:, 4] == 0.0) and np.all(clk[:, 4] == 0.0)
    r = np.linalg.norm(pos[:, [0, 1, 2, 3, 7]], axis=-1)
    assert np.all(r > 2.4e7) and np.all(r < 4.3e7)


def _receiver(f):
    st = f.state
    e = st.extended_pose
    cb = np.array([st.gnss[g].value() if g in st.gnss else 0.0 for g in (GPS, GLO, GAL, BDS)])
    fs = st.gnss[FS].value() if FS in st.gnss else 0.0
    return e.vec1.copy(), e.vec2.copy(), st.gnss[YOF].value(), cb, fs


@pytest.mark.parametrize("lat,lon", [(22.3, 114.2), (-33.9, 151.2)])
def test_residuals_match_oracle(lib, lat, lon):
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    B, S = 3, 10
    st = SyntheticStream(wl, B)
    orc = make_oracles(wl, st, fp)
    for _ in range(5):
        fr = st.next_frame()
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
    rng = np.random.default_rng(11)
    Re = enu2ecef_rotation(lat, lon)
    t0 = geo2ecef(lat, lon, 40.0)
    T = np.ascontiguousarray(np.tile(np.concatenate([Re.reshape(9), t0]), (B, 1)))
    rs = [_receiver(f) for f in orc]
    xs = [gc.receiver_states(p, v, yof, cb, fs, Re, t0) for p, v, yof, cb, fs in rs]
    raw = raw_gnss_epoch(rng, np.array([x[0][:3] for x in xs]), np.array([x[1][:3] for x in xs]), np.array([r[3] for r in rs]),
                         np.array([r[4] for r in rs]), S, lat, lon, no_l1=(3,), below_horizon=(5,))
    X = np.ascontiguousarray(np.stack([oracle_packed_state(f, wl.sw) for f in orc]))
    idx = np.array([orc[0].state.gnss[g].idx() if g in orc[0].state.gnss else -1 for g in range(6)], np.int32)
    a = {k: np.ascontiguousarray(raw[k], np.float64) for k in ("sat_pos", "sat_vel", "sat_clk", "obs", "obs_std", "ttx", "iono")}
    sys_ = np.ascontiguousarray(raw["sys"], np.int32)
    out = dict(unit=np.zeros((B, S, 3)), res_pos=np.zeros((B, S)), res_vel=np.zeros((B, S)), sigma_psr=np.zeros((B, S)),
               sigma_dopp=np.zeros((B, S)), azel=np.zeros((B, S, 2)), atmos=np.zeros((B, S, 2)))
    lib.emu_gnss_residuals(B, S, _p(X), X.shape[1], _p(idx), _p(a["sat_pos"]), _p(a["sat_vel"]), _p(a["sat_clk"]), _p(a["obs"]),
                           _p(a["obs_std"]), _p(a["ttx"]), _p(sys_), _p(T), _p(a["iono"]), 1.3, 0.7, _p(out["unit"]),
                           _p(out["res_pos"]), _p(out["res_vel"]), _p(out["sigma_psr"]), _p(out["sigma_dopp"]), _p(out["azel"]),
                           _p(out["atmos"]))
    for b, f in enumerate(orc):
        p, v, yof, cb, fs = rs[b]
        sat = dict(pos=raw["sat_pos"][b], vel=raw["sat_vel"][b], dt=raw["sat_clk"][b, :, 0], ddt=raw["sat_clk"][b, :, 1],
                   tgd=raw["sat_clk"][b, :, 2], sys=raw["sys"][b], psr=raw["obs"][b, :, 0], dopp=raw["obs"][b, :, 1],
                   freq=raw["obs"][b, :, 2], doy=raw["ttx"][b, :, 0], tow=raw["ttx"][b, :, 1], ura=raw["obs_std"][b, :, 0],
                   psr_std=raw["obs_std"][b, :, 1], dopp_std=raw["obs_std"][b, :, 2])
        ref = gc.epoch_residuals(p, v, yof, cb, fs, Re, t0, sat, raw["iono"][b], psr_amp=1.3, dopp_amp=0.7)
        assert np.abs(out["res_pos"][b] - ref["res_pos"]).max() < 1e-6          # [m] on 2e7 m ranges
        assert np.abs(out["res_vel"][b] - ref["res_vel"]).max() < 1e-8          # [m/s]
        assert np.abs(out["unit"][b] - ref["unit_psr"]).max() < 1e-13
        assert np.abs(out["azel"][b] - ref["azel"]).max() < 1e-11
        assert np.abs(out["atmos"][b] - ref["atmos"]).max() < 1e-8
        assert np.allclose(out["sigma_psr"][b], ref["sigma_psr"], rtol=1e-12, atol=0)
        assert np.allclose(out["sigma_dopp"][b], ref["sigma_dopp"], rtol=1e-12, atol=0)
        assert out["res_pos"][b, 3] == 0.0 and np.all(out["unit"][b, 3] == 0.0)            # no L1 observation
        assert out["azel"][b, 5, 1] < 0 and np.all(out["atmos"][b, 5] == 0.0)              # below the horizon
