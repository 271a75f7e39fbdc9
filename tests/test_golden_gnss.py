"""Golden fixture of the GNSS front end (tests/golden/gnss_frontend.npz, made by tests/golden/make_golden_gnss.py from
the oracle): the oracle must keep reproducing it (CPU) and the CUDA kernels must match it without the oracle (GPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gnss_frontend.npz")


def test_oracle_reproduces_golden():
    import ingvio_oracle.gnss_comm as gc
    d = np.load(GOLD)
    B, S = d["sys"].shape
    Re, t0 = d["T"][0, :9].reshape(3, 3), d["T"][0, 9:]
    for b in range(B):
        for i in range(S):
            k = int(d["sys"][b, i])
            rec = dict(zip(gc.GLO_FIELDS if k == gc.SYS_GLO else gc.KEPLER_FIELDS, d["eph"][b, i]))
            s = gc.sat_state(d["t_obs"][b, i], d["psr"][b, i], k, rec)
            assert np.abs(s["pos"] - d["sat_pos"][b, i]).max() <= 1e-7 and np.abs(s["vel"] - d["sat_vel"][b, i]).max() <= 1e-10
            assert abs(s["dt"] - d["sat_clk"][b, i, 0]) <= 1e-18 and abs(s["ttx_rel"] - d["ttx_rel"][b, i]) <= 1e-12
        sat = dict(pos=d["sat_pos"][b], vel=d["sat_vel"][b], dt=d["sat_clk"][b, :, 0], ddt=d["sat_clk"][b, :, 1],
                   tgd=d["sat_clk"][b, :, 2], sys=d["sys"][b], psr=d["obs"][b, :, 0], dopp=d["obs"][b, :, 1], freq=d["obs"][b, :, 2],
                   doy=d["ttx"][b, :, 0], tow=d["ttx"][b, :, 1], ura=d["obs_std"][b, :, 0], psr_std=d["obs_std"][b, :, 1],
                   dopp_std=d["obs_std"][b, :, 2])
        r = gc.epoch_residuals(d["p_w"][b], d["v_w"][b], d["yof"][b], d["cb"][b], d["fs"][b], Re, t0, sat, d["iono"][b],
                               psr_amp=1.2, dopp_amp=0.8)
        assert np.abs(r["res_pos"] - d["res_pos"][b]).max() <= 1e-7 and np.abs(r["res_vel"] - d["res_vel"][b]).max() <= 1e-9
        assert np.abs(r["azel"] - d["res_azel"][b]).max() <= 1e-12 and np.abs(r["atmos"] - d["res_atmos"][b]).max() <= 1e-9
        assert np.allclose(r["sigma_psr"], d["res_sigma_psr"][b], rtol=1e-13) and np.allclose(r["sigma_dopp"], d["res_sigma_dopp"][b], rtol=1e-13)


@pytest.mark.gpu
def test_cuda_matches_golden():
    from helpers import filter_params, make_gpu
    from ingvio_b200.synth import WORKLOADS, SyntheticStream
    d = np.load(GOLD)
    B, S = d["sys"].shape
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    g = make_gpu(wl, SyntheticStream(wl, B), fp, max_sats=S)
    x = g.get_state()
    x[:, 9:12], x[:, 12:15] = d["p_w"], d["v_w"]
    x[:, 33:37], x[:, 37], x[:, 38] = d["cb"], d["fs"], d["yof"]        # clock biases GPS..BDS, FS, YOF
    g.set_state(x)
    st = g.sat_states(d["eph"], d["t_obs"], d["psr"], d["sys"])
    assert np.abs(st["sat_pos"] - d["sat_pos"]).max() <= 1e-5 and np.abs(st["sat_vel"] - d["sat_vel"]).max() <= 1e-8
    assert np.abs(st["sat_clk"] - d["sat_clk"]).max() <= 1e-15 and np.abs(st["ttx_rel"] - d["ttx_rel"]).max() <= 1e-12
    out = g.gnss_residuals(st["sat_pos"], st["sat_vel"], st["sat_clk"], d["obs"], d["obs_std"], d["ttx"], d["sys"], d["T"],
                           d["iono"], psr_amp=1.2, dopp_amp=0.8)
    assert np.abs(out["res_pos"] - d["res_pos"]).max() <= 1e-5      # the device satellite positions differ by <= 1e-5 m
    assert np.abs(out["res_vel"] - d["res_vel"]).max() <= 1e-7
    assert np.abs(out["unit"] - d["res_unit_psr"]).max() <= 1e-12
    assert np.abs(out["azel"] - d["res_azel"]).max() <= 1e-11 and np.abs(out["atmos"] - d["res_atmos"]).max() <= 1e-8
    assert np.allclose(out["sigma_psr"], d["res_sigma_psr"], rtol=1e-11) and np.allclose(out["sigma_dopp"], d["res_sigma_dopp"], rtol=1e-11)
    assert np.all(out["unit"][:, 4] == 0.0) and np.all(st["sat_pos"][:, 4] == 0.0)       # no L1 observation
