"""TEST INFRASTRUCTURE: runs the reference's own code (oracle/_ref/ref_driver: the unmodified translation units of
/root/reference/ingvio_estimator/src compiled against the stand-in headers of oracle/ref_shim) on a recorded stream and
reads back state + covariance after every frame.  `oracle/_ref/` is built by `make -C oracle/ref_shim` wherever
/root/reference exists; elsewhere (the GPU box) the prebuilt binary travels with the snapshot, and the committed outputs
tests/golden/ref_frames.npz (made by tests/golden/make_golden_ref.py) stand in when even that is missing."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF_SRC = "/root/reference/ingvio_estimator/src"
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_frames.npz")
CONFIGS = [(False, False), (True, False), (False, True), (True, True)]   # (keyframe, stereo) of test_cpp_updaters._stream
GOLDEN_FRAMES = (0, 4, 8, 13)
LM_CONFIGS = [(False, 6), (True, 6)]     # (keyframe, max_lm_feats): SLAM landmarks kept in the state (mono)
# BASELINE-sized stream (configs[1]: mono, SW = 11, 150 tracked features per image): window 11, 150 persistent tracks
BIG = dict(sw=11, n_tracks=150, m=160, n_frames=20)
BIG_FRAMES = (10, 15, 19)


def build_ref():
    """Build oracle/_ref if the reference sources are here; returns True when the driver exists afterwards."""
    if os.path.isdir(REF_SRC):
        r = subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle", "ref_shim")], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + (r.stdout + r.stderr)[-3000:])
    return os.path.exists(REF_DRIVER)


def config_key(keyframe, stereo, max_lm=0):
    return f"{'kf' if keyframe else 'swmarg'}_{'stereo' if stereo else 'mono'}" + (f"_lm{max_lm}" if max_lm else "")


def _read_lm_output(path, max_clones):
    """ref_driver records with the landmark block: (id, covariance index, world xyz) per landmark in the state."""
    d = np.fromfile(path, dtype=np.float64)
    xs = 39 + 12 * max_clones
    recs, pos = [], 0
    while pos < len(d):
        N, ncl, ntr = int(d[pos]), int(d[pos + 1]), int(d[pos + 2])
        x = d[pos + 3:pos + 3 + xs]
        P = d[pos + 3 + xs:pos + 3 + xs + N * N].reshape(N, N).T
        pos += 3 + xs + N * N
        nl = int(d[pos])
        lms = d[pos + 1:pos + 1 + 5 * nl].reshape(nl, 5).copy()
        pos += 1 + 5 * nl
        recs.append(dict(N=N, ncl=ncl, ntr=ntr, x=x, P=P, lms=lms))
    return recs


def run_ref(keyframe, stereo, max_lm=0, **stream_kw):
    """The recorded stream of tests/test_cpp_updaters.py through the reference build; one record per frame.
    max_lm > 0 turns the SLAM-landmark branch of the frame callback on (LandmarkUpdate, mono)."""
    from test_cpp_updaters import SW, _read_output, _stream, _write_input
    wl, fp, st, frames = _stream(keyframe, stereo, **stream_kw)
    SW = stream_kw.get("sw") or SW
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        _write_input(fin, wl, fp, st, frames, keyframe)
        env = dict(os.environ, IGV_REF_MAX_LM=str(int(max_lm)))
        r = subprocess.run([REF_DRIVER, fin, fout], capture_output=True, text=True, timeout=600, env=env)
        if r.returncode != 0 or "FRAMES DONE" not in r.stdout:
            raise RuntimeError("ref_driver failed: " + r.stdout[-500:] + r.stderr[-2000:])
        return _read_lm_output(fout, SW + 1) if max_lm else _read_output(fout, SW + 1)


def load_golden():
    z = np.load(GOLDEN)
    out = {}
    for keyframe, stereo in CONFIGS:
        k = config_key(keyframe, stereo)
        out[k] = {int(f): dict(x=z[f"{k}_x{f}"], P=z[f"{k}_P{f}"]) for f in GOLDEN_FRAMES}
    for keyframe, max_lm in LM_CONFIGS:
        k = config_key(keyframe, False, max_lm)
        out[k] = {int(f): dict(x=z[f"{k}_x{f}"], P=z[f"{k}_P{f}"], lms=z[f"{k}_lms{f}"]) for f in GOLDEN_FRAMES}
    out["big"] = {int(f): dict(x=z[f"big_x{f}"], P=z[f"big_P{f}"], ntr=int(z[f"big_ntr{f}"])) for f in BIG_FRAMES}
    return out
