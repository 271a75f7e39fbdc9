"""The C++ GnssUpdate mirror (ingvio_b200/host/ingvio_gnss.hpp: checkYofStatus / updateTrackedSys / addNewTrackedSys with
the reference's signatures, GnssUpdate.h:52-67, plus StateManager::addVariableIndependent / replaceVarLinear) driven by
tests/cpp/test_gnss_update.cpp on one raw epoch, state and covariance after every call against the oracle
(oracle gnss_comm restatement -> GnssUpdate.update_tracked_sys / add_new_tracked_sys)."""
import os
import subprocess

import numpy as np
import pytest

import ingvio_oracle as o
import ingvio_oracle.gnss_comm as gc
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF, StateManager as SM
from ingvio_oracle.gnss_update import GnssEpoch
from ingvio_b200.synth import enu2ecef_rotation, geo2ecef, raw_gnss_epoch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_gnss_update.cpp")
LIBDIR = os.path.join(ROOT, "ingvio_b200", "lib")
LIBNAME = "ingvio_b200"
if os.environ.get("IGV_TEST_LIB") == "emul":      # development aid: the CPU model of the library (tests/conftest.py)
    LIBDIR, LIBNAME = os.path.join(ROOT, "tests", "emul", "_build"), "ingvio_emul"


def _build():
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_gnss_update")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", SRC, "-o", exe, f"-L{LIBDIR}", f"-l{LIBNAME}", f"-Wl,-rpath,{LIBDIR}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libingvio_b200.so")), reason="library not built")
def test_gnss_mirror_links_against_the_library():
    _build()


def _read(path, xs):
    d = np.fromfile(path, dtype=np.float64)
    recs, pos = [], 0
    while pos < len(d):
        N = int(d[pos])
        x = d[pos + 2:pos + 2 + xs]
        P = d[pos + 2 + xs:pos + 2 + xs + N * N].reshape(N, N).T
        recs.append(dict(N=N, ng=int(d[pos + 1]), x=x, P=P))
        pos += 2 + xs + N * N
    return recs


@pytest.mark.gpu
@pytest.mark.parametrize("adjust_yof", [0, 1])
def test_gnss_mirror_vs_oracle(tmp_path, adjust_yof):
    rng = np.random.default_rng(17 + adjust_yof)
    exe = _build()
    lat, lon = 22.3, 114.2
    Re, anchor = enu2ecef_rotation(lat, lon), geo2ecef(lat, lon, 40.0)
    yaw = 0.3
    psr_amp, dopp_amp, init_cov_yof, thres = 4.0, 2.0, 0.015, 0.95
    R0 = np.eye(3)
    p0, v0 = np.array([3.0, -2.0, 1.0]), np.array([0.5, 0.2, -0.1])
    gn = [(GPS, 4.0, 4.0), (GLO, -6.0, 4.0), (FS, 0.2, 1.0)]          # GAL and BDS are seen for the first time below
    # oracle state
    fp = o.FilterParams(max_sw_clones=4, enable_gnss=1, is_adjust_yof=adjust_yof, psr_noise_amp=psr_amp, dopp_noise_amp=dopp_amp,
                        gnss_chi2_test=0, gnss_strong_reject=0, chi2_thres=thres)
    f = o.OracleFilter(fp, stereo=False, max_valid_ids=1)
    f.init(0.0, R0, p0, v0, np.zeros(3), np.zeros(3))
    for g, val, cov in gn:
        SM.add_gnss_variable(f.state, g, val, cov)
    SM.add_gnss_variable(f.state, YOF, yaw, init_cov_yof)             # checkYofStatus (cov passed as is, GnssUpdate.cpp:40-42)
    N = f.state.cov.shape[0]
    A = rng.standard_normal((N, N)) * 0.05
    P0 = A @ A.T + np.diag(np.concatenate([np.full(21, 1e-2), np.array([4.0, 4.0, 1.0, 0.015])]))
    f.state.cov[:, :] = P0
    # one raw epoch, consistent with the receiver state and with clock biases 4, -6, 9 (GAL), -3 (BDS)
    cb_true = np.array([4.0, -6.0, 9.0, -3.0])
    xyzt, dv = gc.receiver_states(p0, v0, yaw, cb_true, 0.2, Re, anchor)
    S = 12
    raw = raw_gnss_epoch(rng, xyzt[None, :3], dv[None, :3], cb_true[None], np.array([0.2]), S, lat, lon, el_range=(20.0, 85.0))
    spp_pos = np.concatenate([xyzt[:3], cb_true + rng.normal(0, 0.5, 4)])
    spp_vel = np.concatenate([dv[:3], [0.2]])
    out = [psr_amp, dopp_amp, adjust_yof, 0, 0, init_cov_yof, thres] + list(R0.reshape(9)) + list(p0) + list(v0)
    out += [len(gn)] + [x for g in gn for x in g]
    out += list(Re.reshape(9)) + list(anchor) + [yaw]
    out += list(P0.T.reshape(-1))
    out += list(raw["iono"][0]) + [S]
    for i in range(S):
        out += [int(raw["sys"][0, i])] + list(raw["sat_pos"][0, i]) + list(raw["sat_vel"][0, i]) + list(raw["sat_clk"][0, i])
        out += list(raw["obs"][0, i]) + list(raw["obs_std"][0, i]) + list(raw["ttx"][0, i])
    out += list(spp_pos) + list(spp_vel)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.asarray(out, dtype=np.float64).tofile(fin)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "GNSS DONE" in r.stdout, r.stdout + r.stderr
    recs = _read(fout, 39 + 12 * 5)

    def sat_dict():
        return dict(pos=raw["sat_pos"][0], vel=raw["sat_vel"][0], dt=raw["sat_clk"][0, :, 0], ddt=raw["sat_clk"][0, :, 1],
                    tgd=raw["sat_clk"][0, :, 2], sys=raw["sys"][0], psr=raw["obs"][0, :, 0], dopp=raw["obs"][0, :, 1],
                    freq=raw["obs"][0, :, 2], doy=raw["ttx"][0, :, 0], tow=raw["ttx"][0, :, 1], ura=raw["obs_std"][0, :, 0],
                    psr_std=raw["obs_std"][0, :, 1], dopp_std=raw["obs_std"][0, :, 2])

    def epoch(clock_override=None):
        st = f.state
        e = st.extended_pose
        cb = np.array([st.gnss[g].value() if g in st.gnss else 0.0 for g in (GPS, GLO, GAL, BDS)])
        fs = st.gnss[FS].value() if FS in st.gnss else 0.0
        if clock_override:
            for g, v in clock_override.items():
                if g == FS:
                    fs = v
                else:
                    cb[g] = v
        sat = sat_dict()
        ref = gc.epoch_residuals(e.vec1, e.vec2, st.gnss[YOF].value(), cb, fs, Re, anchor, sat, raw["iono"][0], psr_amp=1.0, dopp_amp=1.0)
        return GnssEpoch(unit=ref["unit_psr"], res_pos=ref["res_pos"], res_vel=ref["res_vel"], sys=raw["sys"][0], ura=sat["ura"],
                         psr_std=sat["psr_std"], dopp_std_mps=sat["dopp_std"] * gc.LIGHT_SPEED / sat["freq"], el=ref["azel"][:, 1])

    def check(rec, what):
        Po = f.cov()
        assert rec["P"].shape == Po.shape, (what, rec["P"].shape, Po.shape)
        err = np.linalg.norm(rec["P"] - Po) / max(1.0, np.linalg.norm(Po))
        assert err <= 1e-8, f"{what}: |dP| = {err:.3e}"
        st = f.state
        e = st.extended_pose
        assert np.abs(rec["x"][0:9] - e.rot.reshape(9)).max() <= 1e-9 and np.abs(rec["x"][9:12] - e.vec1).max() <= 1e-9
        assert np.abs(rec["x"][12:15] - e.vec2).max() <= 1e-9
        for g in range(6):
            if g in st.gnss:
                assert abs(rec["x"][33 + g] - st.gnss[g].value()) <= 1e-8 * max(1.0, abs(st.gnss[g].value())), (what, g)

    check(recs[0], "prior")
    f.gnss.update_tracked_sys(f.state, epoch(), Re)
    check(recs[1], "updateTrackedSys")
    spp = {GAL: spp_pos[3 + GAL], BDS: spp_pos[3 + BDS]}
    for g in (GAL, BDS):                                   # GNSSType order, residuals re-evaluated after each addition
        res = f.gnss.add_new_tracked_sys(f.state, epoch(spp), Re, [g], spp, R_ecef2enu=Re.T)
        assert res[g], "the consistent epoch must pass the 0.95 gate"
    check(recs[2], "addNewTrackedSys")
    assert recs[2]["ng"] == 6
    extra = o.Scalar()
    SM.add_variable_independent(f.state, extra, [[0.25]])
    SM.replace_var_linear(f.state, extra, [f.state.bg], np.array([[0.5, -1.0, 2.0]]))
    check(recs[3], "addVariableIndependent + replaceVarLinear")
