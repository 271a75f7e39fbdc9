"""The C++ updater mirror (ingvio_b200/host/ingvio_updaters.hpp: RemoveLostUpdate / SwMargUpdate / KeyframeUpdate with the
reference's interfaces, MapServerManager, State / StateManager) driven frame by frame in the reference's call order by
tests/cpp/test_updaters_frames.cpp on a recorded stream (IMU + tracker messages).

  * CPU: the driver runs against tests/emul/igv_shim.cpp (table calls on the kernel source executed on the CPU, algebra
    stubbed): window policy, slot bookkeeping and the call chain of the mirror are exercised without a GPU;
  * GPU: the same driver against libingvio_b200.so, state and covariance after every frame compared with the oracle filter
    driven by the oracle MapServer (tests/track_frames.py) at the parity bar of the other tests.
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import cov_diag21, filter_params, make_oracles, oracle_packed_state
from ingvio_b200.synth import K_IMU, SyntheticStream, TrackerStream, Workload
from track_frames import OracleFrontEnd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_updaters_frames.cpp")
LIBDIR = os.path.join(ROOT, "ingvio_b200", "lib")
LIBNAME = "ingvio_b200"
if os.environ.get("IGV_TEST_LIB") == "emul":      # development aid: the CPU model of the library (tests/conftest.py)
    LIBDIR, LIBNAME = os.path.join(ROOT, "tests", "emul", "_build"), "ingvio_emul"
EMUL = os.path.join(ROOT, "tests", "emul")
SW, F, M, T, FRAMES = 5, 32, 32, 96, 14


def _stream(keyframe, stereo, sw=None, n_tracks=18, m=None, n_frames=None):
    """The recorded stream: window `sw` (default SW = 5), `n_tracks` persistent landmarks per message of at most `m`."""
    sw, m, n_frames = sw or SW, m or M, n_frames or FRAMES
    wl = Workload("trk", 11 + int(stereo), sw + (0 if keyframe else 1), max(F, m), 0, stereo=stereo)
    fp = filter_params(wl, max_sw_clones=sw, frame_select_interval=2)
    st = SyntheticStream(wl, 1)
    trk = TrackerStream(st, n_tracks, m)
    frames = []
    for _ in range(n_frames):
        st.n_clones = 0
        fr = st.next_frame(with_visual=False, with_gnss=False, marg_oldest=False)
        frames.append((fr,) + trk.message(fr.t))
    return wl, fp, st, frames


def _write_input(path, wl, fp, st, frames, keyframe):
    import ingvio_oracle as o
    sp = o.StateParams(fp)
    ini = st.initial_state()
    rho = wl.rho
    m_ = frames[0][2].shape[1]           # message capacity of this stream
    out = [len(frames), K_IMU, m_, rho, int(keyframe), fp.max_sw_clones, max(F, m_), fp.frame_select_interval, max(T, 4 * m_),
           sp.noise_g, sp.noise_a, sp.noise_bg, sp.noise_ba, sp.noise_clockbias, sp.noise_cb_rw, 0.0, 0.0, -fp.gravity_norm]
    out += list(np.asarray(fp.T_cl2i_R).reshape(9)) + list(fp.T_cl2i_p) + list(np.asarray(fp.T_cl2cr_R).reshape(9)) + list(fp.T_cl2cr_p)
    out += [sp.init_cov_rot, sp.init_cov_pos, sp.init_cov_vel, sp.init_cov_bg, sp.init_cov_ba, sp.init_cov_ext_rot, sp.init_cov_ext_pos]
    out += [fp.visual_noise, fp.chi2_thres]
    out += list(ini["R"][0].reshape(9)) + list(ini["p"][0]) + list(ini["v"][0]) + list(ini["bg"][0]) + list(ini["ba"][0])
    for fr, n, ids, uv in frames:
        out += [fr.t] + list(fr.gyro[0].reshape(-1)) + list(fr.accel[0].reshape(-1)) + list(fr.dt[0]) + [int(n[0])]
        out += [float(i) for i in ids[0]] + list(uv[0].reshape(-1))
    np.asarray(out, dtype=np.float64).tofile(path)


def _read_output(path, max_clones):
    d = np.fromfile(path, dtype=np.float64)
    xs = 39 + 12 * max_clones
    recs, pos = [], 0
    while pos < len(d):
        N, ncl, ntr = int(d[pos]), int(d[pos + 1]), int(d[pos + 2])
        x = d[pos + 3:pos + 3 + xs]
        P = d[pos + 3 + xs:pos + 3 + xs + N * N].reshape(N, N).T     # column-major
        recs.append(dict(N=N, ncl=ncl, ntr=ntr, x=x, P=P))
        pos += 3 + xs + N * N
    return recs


def _build(exe, libdir, libname):
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", SRC, "-o", exe, f"-L{libdir}", f"-l{libname}", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


@pytest.mark.parametrize("keyframe,stereo", [(False, False), (True, False), (False, True)])
def test_updater_mirror_wiring_on_cpu_shim(tmp_path, keyframe, stereo):
    bdir = os.path.join(EMUL, "_build")
    os.makedirs(bdir, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I" + cuda_inc, "-I" + EMUL,
                        os.path.join(EMUL, "igv_shim.cpp"), "-o", os.path.join(bdir, "libigv_shim.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = _build(os.path.join(bdir, "test_updaters_frames_cpu"), bdir, "igv_shim")
    wl, fp, st, frames = _stream(keyframe, stereo)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, wl, fp, st, frames, keyframe)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"FRAMES DONE {FRAMES}" in r.stdout, r.stdout + r.stderr
    recs = _read_output(fout, SW + 1)
    assert len(recs) == FRAMES
    # window policy of the reference: SwMargUpdate keeps SW clones after each frame once full (marginalises when SW+1 are
    # present); KeyframeUpdate drops two clones whenever SW are present
    ncl = [r_["ncl"] for r_ in recs]
    if keyframe:
        exp, n = [], 0
        for _ in range(FRAMES):
            n += 1
            if n >= SW:
                n -= 2
            exp.append(n)
    else:
        exp = [min(k + 1, SW) for k in range(FRAMES)]
    assert ncl == exp, (ncl, exp)
    assert all(r_["N"] == 21 + 6 * r_["ncl"] for r_ in recs)
    assert all(r_["ntr"] > 0 for r_ in recs)
    # lost tracks leave the table (RemoveLostUpdate erases what it selected): it cannot grow without bound
    assert max(r_["ntr"] for r_ in recs) <= 2 * 18 + 8


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libingvio_b200.so")), reason="library not built")
def test_updater_mirror_links_against_the_library():
    _build(os.path.join(ROOT, "tests", "cpp", "_build", "test_updaters_frames"), LIBDIR, LIBNAME)


@pytest.mark.gpu
@pytest.mark.parametrize("keyframe,stereo", [(False, False), (True, False), (False, True)])
def test_updater_mirror_vs_oracle(tmp_path, keyframe, stereo):
    exe = _build(os.path.join(ROOT, "tests", "cpp", "_build", "test_updaters_frames"), LIBDIR, LIBNAME)
    wl, fp, st, frames = _stream(keyframe, stereo)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, wl, fp, st, frames, keyframe)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"FRAMES DONE {FRAMES}" in r.stdout, r.stdout + r.stderr
    recs = _read_output(fout, SW + 1)
    fe = OracleFrontEnd(make_oracles(wl, st, fp, with_gnss=False)[0], keyframe)
    for k, (fr, n, ids, uv) in enumerate(frames):
        fe.frame(fr.seq(0), int(n[0]), ids[0], uv[0])
        Po = fe.f.cov()
        rec = recs[k]
        assert rec["N"] == Po.shape[0] and rec["ncl"] == len(fe.f.state.sw_camleft_poses) and rec["ntr"] == len(fe.ms), (k, rec["N"], rec["ntr"])
        err = np.linalg.norm(rec["P"] - Po) / max(1.0, np.linalg.norm(Po))
        assert err <= 1e-8, f"frame {k}: |dP|_F/max(1,|P|_F) = {err:.3e}"
        xo = oracle_packed_state(fe.f, SW + 1)
        nu = 39 + 12 * rec["ncl"]
        ex = np.max(np.abs(rec["x"][:nu] - xo[:nu]) / np.maximum(1.0, np.abs(xo[:nu])))
        assert ex <= 1e-9, f"frame {k}: state mismatch {ex:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("keyframe,stereo", [(False, False), (True, False), (False, True), (True, True)])
def test_cuda_path_vs_reference_outputs(tmp_path, keyframe, stereo):
    """CUDA path (C++ estimator mirror over the C-ABI) vs the REFERENCE's own outputs for the same recorded stream: the
    committed golden of the reference build (tests/golden/ref_frames.npz, tests/golden/make_golden_ref.py) and, where the
    prebuilt oracle/_ref/ref_driver travelled with the snapshot, a live run of it for every frame.  No oracle in between."""
    import ref_pin
    exe = _build(os.path.join(ROOT, "tests", "cpp", "_build", "test_updaters_frames"), LIBDIR, LIBNAME)
    wl, fp, st, frames = _stream(keyframe, stereo)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, wl, fp, st, frames, keyframe)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"FRAMES DONE {FRAMES}" in r.stdout, r.stdout + r.stderr
    recs = _read_output(fout, SW + 1)

    def close(rec, ref, what):
        assert rec["P"].shape == ref["P"].shape, (what, rec["P"].shape, ref["P"].shape)
        err = np.linalg.norm(rec["P"] - ref["P"]) / max(1.0, np.linalg.norm(ref["P"]))
        assert err <= 1e-8, f"{what}: |dP|_F/max(1,|P|_F) = {err:.3e}"
        nu = 39 + 12 * rec["ncl"]
        ex = np.max(np.abs(rec["x"][:nu] - ref["x"][:nu]) / np.maximum(1.0, np.abs(ref["x"][:nu])))
        assert ex <= 1e-9, f"{what}: state mismatch {ex:.3e}"

    gold = ref_pin.load_golden()[ref_pin.config_key(keyframe, stereo)]
    for f, ref in gold.items():
        close(recs[f], ref, f"frame {f} vs reference golden")
    if os.path.exists(ref_pin.REF_DRIVER):
        live = ref_pin.run_ref(keyframe, stereo)
        for k, ref in enumerate(live):
            assert recs[k]["ntr"] == ref["ntr"] and recs[k]["N"] == ref["N"], (k, recs[k]["ntr"], ref["ntr"])
            close(recs[k], ref, f"frame {k} vs reference live")


@pytest.mark.gpu
def test_cuda_path_vs_reference_baseline_size(tmp_path):
    """The same comparison on the BASELINE-sized stream (configs[1]: mono, window 11, 150 tracks per image): CUDA path
    (C++ estimator mirror over the C-ABI) against the reference build's committed outputs and, where the binary
    travelled, its live run frame by frame."""
    import ref_pin
    big = ref_pin.BIG
    exe = _build(os.path.join(ROOT, "tests", "cpp", "_build", "test_updaters_frames"), LIBDIR, LIBNAME)
    wl, fp, st, frames = _stream(False, False, **big)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_input(fin, wl, fp, st, frames, False)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"FRAMES DONE {big['n_frames']}" in r.stdout, r.stdout + r.stderr
    recs = _read_output(fout, big["sw"] + 1)

    def close(rec, ref, what):
        assert rec["P"].shape == ref["P"].shape, (what, rec["P"].shape, ref["P"].shape)
        err = np.linalg.norm(rec["P"] - ref["P"]) / max(1.0, np.linalg.norm(ref["P"]))
        assert err <= 1e-8, f"{what}: |dP|_F/max(1,|P|_F) = {err:.3e}"
        nu = 39 + 12 * rec["ncl"]
        ex = np.max(np.abs(rec["x"][:nu] - ref["x"][:nu]) / np.maximum(1.0, np.abs(ref["x"][:nu])))
        assert ex <= 1e-9, f"{what}: state mismatch {ex:.3e}"

    for f, ref in ref_pin.load_golden()["big"].items():
        assert recs[f]["ntr"] == ref["ntr"]
        close(recs[f], ref, f"baseline-size frame {f} vs reference golden")
    if os.path.exists(ref_pin.REF_DRIVER):
        for k, ref in enumerate(ref_pin.run_ref(False, False, **big)):
            assert recs[k]["ntr"] == ref["ntr"], (k, recs[k]["ntr"], ref["ntr"])
            close(recs[k], ref, f"baseline-size frame {k} vs reference live")


def test_imu_buffer_matches_oracle(tmp_path):
    """ImuPropagator::storeImu / propagateUntil of the C++ mirror vs the oracle restatement of ImuPropagator.cpp:232-292 on
    irregular stamps and awkward end times: the steps handed to the device, the state time and the buffer must agree
    exactly (same double arithmetic)."""
    import ingvio_oracle as o
    rng = np.random.default_rng(4)
    bdir = os.path.join(EMUL, "_build")
    os.makedirs(bdir, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I" + cuda_inc, "-I" + EMUL,
                        os.path.join(EMUL, "igv_shim.cpp"), "-o", os.path.join(bdir, "libigv_shim.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(bdir, "test_imu_buffer")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_imu_buffer.cpp"), "-o", exe,
                        f"-L{bdir}", "-ligv_shim", f"-Wl,-rpath,{bdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # events: irregular sample times, end times before the first sample, between samples, on a sample, within 1e-6 of the
    # state time, beyond the last sample, and with an empty buffer
    t0 = 10.0
    events, t = [("init", t0)], t0
    stamps = []
    for _ in range(60):
        t += rng.uniform(0.002, 0.008)
        stamps.append(t)
    cursor = 0

    def store(n):
        nonlocal cursor
        for s in stamps[cursor:cursor + n]:
            events.append(("imu", s, rng.standard_normal(3), rng.standard_normal(3)))
        cursor += n

    events.append(("until", t0 + 0.01))                 # empty buffer
    store(5)
    events.append(("until", t0 - 1.0))                  # t_end <= state time
    events.append(("until", stamps[0] - 1e-4))          # before the first sample: nothing happens (:240)
    events.append(("until", 0.5 * (stamps[2] + stamps[3])))   # between samples: last partial step with sample 2
    events.append(("until", stamps[3] + 5e-7))          # the next sample is closer than 1e-6 past ... and t_end too
    store(10)
    events.append(("until", stamps[9]))                 # exactly on a sample
    events.append(("until", stamps[14] + 0.02))         # beyond the last stored sample: partial step with the last one
    store(20)                                           # these include samples OLDER than the state time now
    events.append(("until", stamps[30]))
    store(25)
    events.append(("until", stamps[59] - 1e-7))
    events.append(("until", stamps[59] + 1e-3))
    flat = []
    for e in events:
        if e[0] == "init":
            flat += [e[1]]
        elif e[0] == "imu":
            flat += [0.0, e[1]] + list(e[2]) + list(e[3])
        else:
            flat += [1.0, e[1]]
    fin, fout = str(tmp_path / "ev.bin"), str(tmp_path / "out.bin")
    np.asarray(flat, np.float64).tofile(fin)
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "EVENTS DONE" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(fout, np.float64)
    # the oracle, with the transition replaced by a recorder (the loop logic is what is compared)
    st = o.State(o.FilterParams(max_sw_clones=3, enable_gnss=0))
    st.timestamp = t0
    prop = o.ImuPropagator(9.8)
    steps = []

    def record(state, ctrl, dt, is_analytic=True):
        steps.append((ctrl.gyro_raw.copy(), ctrl.accel_raw.copy(), dt))
        state.timestamp += dt
        return np.eye(15), np.zeros((15, 12))
    prop.state_and_cov_transition = record
    from ingvio_oracle import StateManager
    orig = StateManager.propagate_state_cov
    StateManager.propagate_state_cov = staticmethod(lambda *a, **k: None)
    try:
        pos, n_until, n_steps_total = 0, 0, 0
        for e in events[1:]:
            if e[0] == "imu":
                prop.store_imu(o.ImuCtrl(e[1], e[2], e[3]))        # (timestamp, gyro, accel)
                continue
            steps.clear()
            prop.propagate_until(st, e[1])
            ts, nbuf, n = got[pos], int(got[pos + 1]), int(got[pos + 2])
            assert ts == st.timestamp, (e, ts, st.timestamp)
            assert nbuf == len(prop.imu_ctrl_buffer), (e, nbuf, len(prop.imu_ctrl_buffer))
            assert n == len(steps), (e, n, len(steps))
            for s in range(n):
                rec = got[pos + 3 + 7 * s:pos + 3 + 7 * (s + 1)]
                assert np.array_equal(rec[0:3], steps[s][0]) and np.array_equal(rec[3:6], steps[s][1]) and rec[6] == steps[s][2], (e, s)
            pos += 3 + 7 * n
            n_until += 1
            n_steps_total += n
        assert pos == len(got) and n_until == 10 and n_steps_total > 40
    finally:
        StateManager.propagate_state_cov = orig

