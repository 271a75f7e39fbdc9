"""The oracle's GNSS restatement is pinned to the REFERENCE ITSELF: GnssUpdate::checkYofStatus / updateTrackedSys /
addNewTrackedSys (GnssUpdate.cpp:33-476), GnssManager, StateManager::addVariableDelayed (Givens sweep, 0.95 chi^2 gate,
invertible initialisation) and gnss_comm's sat_states / eph2pos / geph2pos / psr_res / dopp_res / Saastamoinen-Niell /
Klobuchar (gnss_comm/src/gnss_spp.cpp, gnss_utility.cpp), all compiled unmodified against the stand-in headers of
oracle/ref_shim and driven by oracle/ref_shim/ref_gnss_driver.cpp on one epoch of broadcast ephemerides and raw L1
observations (rows a17, a18 and the "next" row f-2 of SURVEY.md section 8).  Live where oracle/_ref exists, and against
the committed outputs tests/golden/ref_gnss.npz (made by `PYTHONPATH=.:oracle:tests python tests/test_ref_pin_gnss.py --golden`)."""
import datetime
import os
import subprocess
import sys

import numpy as np
import pytest

import ingvio_oracle as o
import ingvio_oracle.gnss_comm as gc
from ingvio_oracle import BDS, FS, GAL, GLO, GPS, YOF, StateManager as SM
from ingvio_oracle.gnss_update import GnssEpoch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_gnss_driver")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_gnss.npz")
GPST0, BDT0 = 315964800, 1136073600
T_OBS = GPST0 + 2200 * 604800 + 345600
L1 = {0: 1575.42e6, 1: 1602.0e6, 2: 1575.42e6, 3: 1561.098e6}


def _doy(t):
    d = datetime.datetime(1970, 1, 1) + datetime.timedelta(seconds=int(np.floor(t)))
    y0 = (datetime.datetime(d.year, 1, 1) - datetime.datetime(1970, 1, 1)).total_seconds()
    return (t - y0) / 86400.0 + 1.0


def _scenario(adjust_yof, seed=5):
    from ingvio_b200.synth import enu2ecef_rotation, geo2ecef, random_ephemerides
    rng = np.random.default_rng(seed)
    lat, lon = 22.3, 114.2
    Re, anchor = enu2ecef_rotation(lat, lon), geo2ecef(lat, lon, 40.0)
    yaw = 0.3
    p0, v0 = np.array([3.0, -2.0, 1.0]), np.array([0.5, 0.2, -0.1])
    cb_true, fs_true = np.array([4.0, -6.0, 9.0, -3.0]), 0.2
    eph, sys_, _, _ = random_ephemerides(rng, 1, 160)
    eph, sys_ = eph[0], sys_[0]
    xyzt, dv = gc.receiver_states(p0, v0, yaw, cb_true, fs_true, Re, anchor)
    sats = []
    per_sys = {0: 0, 1: 0, 2: 0, 3: 0}
    for i in range(len(sys_)):
        k = int(sys_[i])
        if per_sys[k] >= 3:
            continue
        t_rel = float(rng.integers(-3000, 3000))
        toe = T_OBS - int(t_rel)
        rec = eph[i].copy()
        if k != 1:
            rec[18] = float((toe - GPST0) % 604800) if k != 3 else float((toe - 14 - BDT0) % 604800)
            rec[20] = float(rec[20])       # time_diff(toe, toc): 0, 16 or -7200 s
        recd = dict(zip(gc.GLO_FIELDS if k == gc.SYS_GLO else gc.KEPLER_FIELDS, rec))
        st = gc.sat_state(t_rel, 2.3e7, k, recd)
        az, el = gc.sat_azel(xyzt[:3], st["pos"])
        if el < np.deg2rad(20.0):
            continue
        per_sys[k] += 1
        sats.append(dict(sys=k, prn=int(rec[21]) if k != 1 else 5 + per_sys[k], toe=toe, t_rel=t_rel, rec=rec, recd=recd))
    assert all(v == 3 for v in per_sys.values()), per_sys
    S = len(sats)
    iono = np.array([0.1118e-7, -0.7451e-8, -0.5961e-7, 0.1192e-6, 0.1167e6, -0.2294e6, -0.1311e6, 0.1049e7])
    psr = np.full(S, 2.3e7)
    dopp = np.zeros(S)
    freq = np.array([L1[s["sys"]] for s in sats])
    noise_p, noise_d = rng.normal(0, 1.5, S), rng.normal(0, 0.05, S)

    def sat_arrays(psr_, dopp_):
        st = [gc.sat_state(s["t_rel"], psr_[i], s["sys"], s["recd"]) for i, s in enumerate(sats)]
        ttx = np.array([s["toe"] + x["ttx_rel"] for s, x in zip(sats, st)])
        return dict(pos=np.array([x["pos"] for x in st]), vel=np.array([x["vel"] for x in st]), dt=np.array([x["dt"] for x in st]),
                    ddt=np.array([x["ddt"] for x in st]), tgd=np.array([x["tgd"] for x in st]), sys=np.array([s["sys"] for s in sats]),
                    psr=psr_.copy(), dopp=dopp_.copy(), freq=freq, doy=np.array([_doy(t) for t in ttx]),
                    tow=np.array([(t - GPST0) % 604800 for t in ttx]), ura=np.full(S, 2.0), psr_std=np.full(S, 1.0),
                    dopp_std=np.full(S, 1.0))

    for _ in range(4):       # measured pseudo-range / Doppler consistent with the true receiver up to the noise
        sat = sat_arrays(psr, dopp)
        ref = gc.epoch_residuals(p0, v0, yaw, cb_true, fs_true, Re, anchor, sat, iono)
        psr = psr + ref["res_pos"] + noise_p - (psr - sat["psr"])
        psr = sat["psr"] + ref["res_pos"] + noise_p          # est + noise
        dopp = sat["dopp"] - (ref["res_vel"] - noise_d) * freq / gc.LIGHT_SPEED
    gn = [(GPS, cb_true[0] + 0.5, 4.0), (GLO, cb_true[1] - 0.5, 4.0), (FS, fs_true + 0.02, 1.0)]
    spp_pos = np.concatenate([xyzt[:3], cb_true + rng.normal(0, 0.5, 4)])
    spp_vel = np.concatenate([dv[:3], [fs_true]])
    N = 21 + len(gn) + 1
    A = rng.standard_normal((N, N)) * 0.05
    P0 = A @ A.T + np.diag(np.concatenate([np.full(21, 1e-2), np.array([4.0, 4.0, 1.0, 0.015])]))
    return dict(adjust_yof=adjust_yof, Re=Re, anchor=anchor, yaw=yaw, p0=p0, v0=v0, sats=sats, psr=psr, dopp=dopp, freq=freq,
                iono=iono, gn=gn, spp_pos=spp_pos, spp_vel=spp_vel, P0=P0, sat_arrays=sat_arrays, psr_amp=4.0, dopp_amp=2.0,
                add_order=[GAL, BDS])


def _run_reference(sc, tmp):
    out = [sc["psr_amp"], sc["dopp_amp"], sc["adjust_yof"], 0, 0, 0.015, 0.95] + list(np.eye(3).reshape(9)) + list(sc["p0"]) + list(sc["v0"])
    out += [len(sc["gn"])] + [x for g in sc["gn"] for x in g]
    out += list(sc["Re"].reshape(9)) + list(sc["anchor"]) + [sc["yaw"]]
    out += list(sc["P0"].T.reshape(-1)) + list(sc["iono"]) + [float(T_OBS), len(sc["sats"])]
    for i, s in enumerate(sc["sats"]):
        out += [s["sys"], s["prn"], float(s["toe"])] + list(s["rec"]) + [sc["psr"][i], sc["dopp"][i], sc["freq"][i], 2.0, 1.0, 1.0]
    out += list(sc["spp_pos"]) + list(sc["spp_vel"]) + [len(sc["add_order"])] + [float(g) for g in sc["add_order"]]
    fin, fout = os.path.join(tmp, "gin.bin"), os.path.join(tmp, "gout.bin")
    np.asarray(out, dtype=np.float64).tofile(fin)
    r = subprocess.run([DRIVER, fin, fout], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "GNSS REF DONE" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]
    d = np.fromfile(fout, dtype=np.float64)
    recs, pos = [], 0
    while pos < len(d):
        N = int(d[pos])
        gidx, x = d[pos + 2:pos + 8], d[pos + 8:pos + 47]
        P = d[pos + 47:pos + 47 + N * N].reshape(N, N).T
        recs.append(dict(N=N, ng=int(d[pos + 1]), gidx=gidx.copy(), x=x.copy(), P=P.copy()))
        pos += 47 + N * N
    return recs


def _run_oracle(sc):
    fp = o.FilterParams(max_sw_clones=4, enable_gnss=1, is_adjust_yof=sc["adjust_yof"], psr_noise_amp=sc["psr_amp"],
                        dopp_noise_amp=sc["dopp_amp"], gnss_chi2_test=0, gnss_strong_reject=0, chi2_thres=0.95)
    f = o.OracleFilter(fp, stereo=False, max_valid_ids=1)
    f.init(0.0, np.eye(3), sc["p0"], sc["v0"], np.zeros(3), np.zeros(3))
    for g, val, cov in sc["gn"]:
        SM.add_gnss_variable(f.state, g, val, cov)
    SM.add_gnss_variable(f.state, YOF, sc["yaw"], 0.015)
    f.state.cov[:, :] = sc["P0"]
    sat = sc["sat_arrays"](sc["psr"], sc["dopp"])

    def record():
        st = f.state
        x = np.zeros(39)
        e = st.extended_pose
        x[0:9], x[9:12], x[12:15] = e.rot.reshape(9), e.vec1, e.vec2
        x[15:18], x[18:21] = st.bg.value(), st.ba.value()
        gidx = np.full(6, -1.0)
        for g, v in st.gnss.items():
            x[33 + g] = v.value()
            gidx[g] = v.idx()
        return dict(N=st.cov.shape[0], gidx=gidx, x=x, P=st.cov.copy())

    def epoch(override=None):
        st = f.state
        e = st.extended_pose
        cb = np.array([st.gnss[g].value() if g in st.gnss else 0.0 for g in (GPS, GLO, GAL, BDS)])
        fs = st.gnss[FS].value() if FS in st.gnss else 0.0
        for g, v in (override or {}).items():
            if g == FS:
                fs = v
            else:
                cb[g] = v
        ref = gc.epoch_residuals(e.vec1, e.vec2, st.gnss[YOF].value(), cb, fs, sc["Re"], sc["anchor"], sat, sc["iono"])
        return GnssEpoch(unit=ref["unit_psr"], res_pos=ref["res_pos"], res_vel=ref["res_vel"], sys=sat["sys"], ura=sat["ura"],
                         psr_std=sat["psr_std"], dopp_std_mps=sat["dopp_std"] * gc.LIGHT_SPEED / sat["freq"], el=ref["azel"][:, 1])

    recs = [record()]
    f.gnss.update_tracked_sys(f.state, epoch(), sc["Re"])
    recs.append(record())
    for g in sc["add_order"]:
        val = sc["spp_pos"][3 + g] if g != FS else sc["spp_vel"][3]
        res = f.gnss.add_new_tracked_sys(f.state, epoch({g: val}), sc["Re"], [g], {g: val}, R_ecef2enu=sc["Re"].T)
        recs.append(dict(record(), accepted=bool(res[g])))
    return recs


def _compare(ref, orc, what):
    assert ref["N"] == orc["N"], (what, ref["N"], orc["N"])
    assert np.array_equal(ref["gidx"], orc["gidx"]), (what, ref["gidx"], orc["gidx"])
    eP = np.linalg.norm(ref["P"] - orc["P"]) / max(1.0, np.linalg.norm(orc["P"]))
    assert eP <= 1e-9, f"{what}: |dP|_F = {eP:.3e}"
    ex = np.max(np.abs(ref["x"] - orc["x"]) / np.maximum(1.0, np.abs(orc["x"])))
    assert ex <= 1e-9, f"{what}: state {ex:.3e}"


STAGES = ("prior", "updateTrackedSys", "addNewTrackedSys GAL", "addNewTrackedSys BDS")


@pytest.mark.parametrize("adjust_yof", [0, 1])
def test_oracle_gnss_matches_reference_golden(adjust_yof):
    z = np.load(GOLDEN)
    orc = _run_oracle(_scenario(adjust_yof))
    assert orc[-1]["N"] == 27 and orc[2]["accepted"] and orc[3]["accepted"]      # both new constellations passed the gate
    for k, what in enumerate(STAGES):
        ref = dict(N=int(z[f"y{adjust_yof}_N{k}"]), gidx=z[f"y{adjust_yof}_g{k}"], x=z[f"y{adjust_yof}_x{k}"], P=z[f"y{adjust_yof}_P{k}"])
        _compare(ref, orc[k], f"{what} (golden, adjust_yof={adjust_yof})")
    # the epoch is a sane one: the update moved the position by centimetres to metres, not kilometres
    assert 1e-4 < np.abs(orc[1]["x"][9:12] - orc[0]["x"][9:12]).max() < 5.0


@pytest.mark.parametrize("adjust_yof", [0, 1])
def test_oracle_gnss_matches_reference_live(adjust_yof, tmp_path):
    import ref_pin
    if not ref_pin.build_ref() or not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_gnss_driver not built and /root/reference absent: the golden test above is the pin")
    sc = _scenario(adjust_yof)
    ref, orc = _run_reference(sc, str(tmp_path)), _run_oracle(sc)
    assert len(ref) == len(orc) == 4
    for k, what in enumerate(STAGES):
        _compare(ref[k], orc[k], f"{what} (live, adjust_yof={adjust_yof})")


if __name__ == "__main__" and "--golden" in sys.argv:
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_pin
    assert ref_pin.build_ref() and os.path.exists(DRIVER)
    out = {}
    for y in (0, 1):
        with tempfile.TemporaryDirectory() as d:
            for k, r in enumerate(_run_reference(_scenario(y), d)):
                out[f"y{y}_N{k}"], out[f"y{y}_g{k}"], out[f"y{y}_x{k}"], out[f"y{y}_P{k}"] = r["N"], r["gidx"], r["x"], r["P"]
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
