"""Golden fixtures of the track table (tests/golden/tracks_*.pkl.gz, made by tests/golden/make_golden_tracks.py from the
oracle MapServer): the oracle is pinned against drift, the kernel source executed on the CPU (tests/emul) and -- with
-m gpu -- the CUDA library are compared with the committed vectors without running the oracle."""
import gzip
import os
import pickle
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))
sys.path.insert(0, os.path.join(HERE, "golden"))

from track_scenario import replay  # noqa: E402


def _load(name):
    with gzip.open(os.path.join(HERE, "golden", f"tracks_{name}.pkl.gz"), "rb") as f:
        return pickle.load(f)


def _deep_equal(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys(), path
        for k in a:
            _deep_equal(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _deep_equal(x, y, f"{path}[{i}]")
    elif isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), path
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("name", ["mono", "stereo"])
def test_oracle_reproduces_golden(name):
    """Oracle drift pin: re-recording the scenario gives the committed events, bit for bit."""
    import make_golden_tracks
    _deep_equal(make_golden_tracks.make(name), _load(name))


@pytest.mark.parametrize("name", ["mono", "stereo"])
def test_emulated_kernels_vs_golden(name):
    from trk_emul import EmulatedTrackTable
    G = _load(name)
    tab = EmulatedTrackTable(G["B"], G["cap"], G["F"], G["T"], G["case"]["stereo"])
    try:
        assert replay(tab, tab.augment, tab.marg, G["events"]) > 40
    finally:
        tab.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mono", "stereo"])
def test_cuda_vs_golden(name):
    from ingvio_b200.filter import BatchFilter
    G = _load(name)
    B = G["B"]
    g = BatchFilter(B, G["cap"], G["F"], 1, stereo=G["case"]["stereo"])
    eye = np.tile(np.eye(3).reshape(1, 9), (B, 1))
    z = np.zeros((B, 3))
    g.init_state_and_cov(eye, z, z, z, z, eye, z, np.full(21, 1e-2))
    g.create_map_server(G["T"])

    def augment(R, p):
        g.augment_sliding_window_pose_cov(eye, np.asarray(R).reshape(B, 9), np.asarray(p).reshape(B, 3))

    assert replay(g, augment, g.marg_sliding_window_pose, G["events"]) > 40
    g.close()
