"""Generates tests/golden/ref_frames.npz: state and covariance written by the REFERENCE's own code (oracle/_ref/ref_driver,
built from /root/reference by oracle/ref_shim/Makefile) after frames 0, 4, 8 and 13 of the recorded streams of
tests/test_cpp_updaters._stream (SwMarg / keyframe x mono / stereo).  Run in the container that holds /root/reference:
    python tests/golden/make_golden_ref.py
These are reference outputs, not oracle outputs: tests/test_ref_pin.py holds the numpy oracle (and, on the GPU, the CUDA
path) to them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import ref_pin  # noqa: E402


def main():
    assert ref_pin.build_ref(), "needs /root/reference (or a prebuilt oracle/_ref/ref_driver)"
    out = {}
    for keyframe, stereo in ref_pin.CONFIGS:
        recs = ref_pin.run_ref(keyframe, stereo)
        k = ref_pin.config_key(keyframe, stereo)
        for f in ref_pin.GOLDEN_FRAMES:
            out[f"{k}_x{f}"] = recs[f]["x"]
            out[f"{k}_P{f}"] = recs[f]["P"]
    for keyframe, max_lm in ref_pin.LM_CONFIGS:
        recs = ref_pin.run_ref(keyframe, False, max_lm)
        k = ref_pin.config_key(keyframe, False, max_lm)
        for f in ref_pin.GOLDEN_FRAMES:
            out[f"{k}_x{f}"] = recs[f]["x"]
            out[f"{k}_P{f}"] = recs[f]["P"]
            out[f"{k}_lms{f}"] = recs[f]["lms"]
    recs = ref_pin.run_ref(False, False, **ref_pin.BIG)      # BASELINE-sized window and track count
    for f in ref_pin.BIG_FRAMES:
        out[f"big_x{f}"], out[f"big_P{f}"], out[f"big_ntr{f}"] = recs[f]["x"], recs[f]["P"], recs[f]["ntr"]
    np.savez_compressed(ref_pin.GOLDEN, **out)
    print("wrote", ref_pin.GOLDEN, os.path.getsize(ref_pin.GOLDEN), "bytes")


if __name__ == "__main__":
    main()
