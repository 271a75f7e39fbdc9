"""Generates tests/golden/tracks_{mono,stereo}.pkl.gz: a recorded track-table scenario -- every input (clone poses, tracker
messages, triangulation results) and every expectation of the ORACLE MapServer (oracle/ingvio_oracle/map_server.py: table
snapshots after each call, the gathered selections in std::map order) -- so that the CUDA path is compared with committed
vectors without running the oracle, and the oracle is pinned against drift. The recording backend is the CPU execution of
the kernel source (tests/emul), whose agreement with the oracle is asserted while recording.

    python tests/golden/make_golden_tracks.py
"""
import gzip
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, os.path.join(ROOT, "oracle"), TESTS, os.path.join(TESTS, "emul")):
    sys.path.insert(0, p)

from trk_emul import EmulatedTrackTable  # noqa: E402
from track_scenario import run_scenario  # noqa: E402

CASES = {"mono": dict(mode="sw_marg", stereo=False, SW=4, seed=101), "stereo": dict(mode="keyframe", stereo=True, SW=5, seed=102)}
B, F, T, FRAMES = 2, 16, 32, 9


def make(name):
    c = CASES[name]
    cap = c["SW"] + 1 if c["mode"] == "sw_marg" else c["SW"]
    tab = EmulatedTrackTable(B, cap, F, T, c["stereo"])
    rec = []
    try:
        run_scenario(tab, tab.augment, tab.marg, tab.clone_poses, c["mode"], B, c["SW"], c["stereo"], frames=FRAMES,
                     seed=c["seed"], F=F, meas_target=10, meas_stride=16, rec=rec)
    finally:
        tab.close()
    return dict(case=c, B=B, F=F, T=T, cap=cap, events=rec)


if __name__ == "__main__":
    for name in CASES:
        out = os.path.join(HERE, f"tracks_{name}.pkl.gz")
        with gzip.GzipFile(out, "wb", mtime=0) as f:
            pickle.dump(make(name), f, protocol=4)
        print(out, os.path.getsize(out))
