"""Generates tests/golden/gnss_frontend.npz: seeded ephemerides / raw L1 observations and what the ORACLE's
gnss_comm restatement (oracle/ingvio_oracle/gnss_comm.py) derives from them -- satellite states (sat_states) and, at
a fixed receiver state, residuals / unit vectors / az-el / delays / sigmas (psr_res, dopp_res).

gnss_comm has no upstream tests or vectors; the fixture pins the oracle against drift (CPU,
tests/test_golden_gnss.py::test_oracle_reproduces_golden) and is the oracle-free target of the CUDA kernels (GPU,
::test_cuda_matches_golden). Regenerate with `python tests/golden/make_golden_gnss.py` only when the oracle changes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import ingvio_oracle.gnss_comm as gc  # noqa: E402
from ingvio_b200.synth import KLOBUCHAR, enu2ecef_rotation, geo2ecef, random_ephemerides  # noqa: E402

LAT, LON, H = 22.3, 114.2, 40.0
P_W = np.array([[12.0, -7.5, 1.2], [-3.0, 20.0, 0.4]])          # receiver position in the world (ENU-aligned up to yaw) frame
V_W = np.array([[4.0, 1.5, -0.1], [-2.0, 3.0, 0.2]])
YOF = np.array([0.3, -0.1])
CB = np.array([[10.0, -20.0, 30.0, 5.0], [1.0, 2.0, -3.0, 4.0]])
FS = np.array([0.2, -0.4])


def generate(S=12, seed=77):
    rng = np.random.default_rng(seed)
    B = 2
    eph, sys_, t_obs, _ = random_ephemerides(rng, B, S, geo=(7,))
    Re, t0 = enu2ecef_rotation(LAT, LON), geo2ecef(LAT, LON, H)
    out = dict(eph=eph, sys=sys_, t_obs=t_obs, iono=np.tile(KLOBUCHAR, (B, 1)),
               T=np.tile(np.concatenate([Re.reshape(9), t0]), (B, 1)), p_w=P_W, v_w=V_W, yof=YOF, cb=CB, fs=FS)
    # pseudo-ranges consistent with the geometry: two fixed-point passes over the transmit time
    psr = np.full((B, S), 2.2e7)
    for _ in range(3):
        for b in range(B):
            xyzt, _ = gc.receiver_states(P_W[b], V_W[b], YOF[b], CB[b], FS[b], Re, t0)
            for i in range(S):
                k = int(sys_[b, i])
                rec = dict(zip(gc.GLO_FIELDS if k == gc.SYS_GLO else gc.KEPLER_FIELDS, eph[b, i]))
                s = gc.sat_state(t_obs[b, i], psr[b, i], k, rec)
                psr[b, i] = np.linalg.norm(s["pos"] - xyzt[:3]) + CB[b, k] - s["dt"] * gc.LIGHT_SPEED + 7.0
    psr += rng.normal(0, 1.0, (B, S))
    psr[:, 4] = 0.0                                      # a satellite without an L1 observation
    keys = ("pos", "vel", "dt", "ddt", "tgd", "ttx_rel")
    st = {k: [] for k in keys}
    for b in range(B):
        row = {k: [] for k in keys}
        for i in range(S):
            k = int(sys_[b, i])
            rec = dict(zip(gc.GLO_FIELDS if k == gc.SYS_GLO else gc.KEPLER_FIELDS, eph[b, i]))
            s = gc.sat_state(t_obs[b, i], psr[b, i], k, rec)
            for kk in keys:
                row[kk].append(s[kk])
        for kk in keys:
            st[kk].append(np.array(row[kk]))
    sat_pos, sat_vel = np.stack(st["pos"]), np.stack(st["vel"])
    sat_clk = np.stack([np.stack(st["dt"]), np.stack(st["ddt"]), np.stack(st["tgd"])], -1)
    freq = np.where(psr > 0, np.vectorize({0: 1575.42e6, 1: 1602.0e6, 2: 1575.42e6, 3: 1561.098e6}.get)(sys_), -1.0)
    dopp = rng.normal(0, 800.0, (B, S))
    obs = np.stack([psr, dopp, freq], -1)
    obs_std = np.stack([np.full((B, S), 2.0), rng.uniform(0.5, 2.0, (B, S)), rng.uniform(0.5, 2.0, (B, S))], -1)
    ttx = np.stack([np.full((B, S), 123.0) + np.stack(st["ttx_rel"]) / 86400.0, 3.0e5 + np.stack(st["ttx_rel"])], -1)
    out.update(psr=psr, sat_pos=sat_pos, sat_vel=sat_vel, sat_clk=sat_clk, ttx_rel=np.stack(st["ttx_rel"]), obs=obs,
               obs_std=obs_std, ttx=ttx)
    res = {k: [] for k in ("unit_psr", "res_pos", "res_vel", "sigma_psr", "sigma_dopp", "azel", "atmos")}
    for b in range(B):
        sat = dict(pos=sat_pos[b], vel=sat_vel[b], dt=sat_clk[b, :, 0], ddt=sat_clk[b, :, 1], tgd=sat_clk[b, :, 2], sys=sys_[b],
                   psr=obs[b, :, 0], dopp=obs[b, :, 1], freq=obs[b, :, 2], doy=ttx[b, :, 0], tow=ttx[b, :, 1],
                   ura=obs_std[b, :, 0], psr_std=obs_std[b, :, 1], dopp_std=obs_std[b, :, 2])
        r = gc.epoch_residuals(P_W[b], V_W[b], YOF[b], CB[b], FS[b], Re, t0, sat, KLOBUCHAR, psr_amp=1.2, dopp_amp=0.8)
        for k in res:
            res[k].append(r[k])
    out.update({"res_" + k if not k.startswith("res_") else k: np.stack(v) for k, v in res.items()})
    return out


if __name__ == "__main__":
    d = generate()
    path = os.path.join(HERE, "gnss_frontend.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path) // 1024, "KiB", "max |res_pos| =", float(np.abs(d["res_pos"]).max()))
