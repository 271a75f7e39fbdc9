"""tests/precision_sweep.py: its FP64 variant must BE the oracle (so the reduced-precision variants measure precision and
nothing else), and the reduced variants must stay inside the bounds profiles/r01_fp32_sweep.md quotes for them."""
import numpy as np

from helpers import filter_params, make_oracles
from ingvio_b200.synth import WORKLOADS, SyntheticStream
from precision_sweep import SweepFilter, sweep, tf32_round


def test_fp64_variant_is_the_oracle():
    wl = WORKLOADS["tiny"]
    fp = filter_params(wl)
    sa, sb = SyntheticStream(wl, 1), SyntheticStream(wl, 1)
    ref = make_oracles(wl, sa, fp)[0]
    sf = SweepFilter(make_oracles(wl, sb, fp)[0], "fp64", fp, wl.feats)
    accepted = 0
    for k in range(10):
        ref.step(sa.next_frame().seq(0))
        sf.step(sb.next_frame().seq(0))
        P0, P1 = ref.cov(), sf.f.cov()
        assert np.linalg.norm(P1 - P0) <= 1e-10 * max(1.0, np.linalg.norm(P0)), k
        assert np.linalg.norm(sf.f.state.extended_pose.vec1 - ref.state.extended_pose.vec1) <= 1e-10, k
        accepted += [g[3] for g in ref.last.get("gammas", [])].count(True)
    assert accepted > 20


def test_tf32_rounding():
    x = np.array([1.0, 1.0 + 2.0 ** -10, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, -3.14159274], np.float32)
    y = tf32_round(x)
    assert y[0] == 1.0 and y[1] == np.float32(1.0 + 2.0 ** -10)
    assert y[2] == 1.0                                   # tie -> even
    assert y[3] == np.float32(1.0 + 2.0 ** -9)           # tie -> even (upwards)
    assert abs(y[4] - x[4]) <= 2.0 ** -10 * 2.0          # half an ulp of a 10-bit mantissa at |x| in [2,4)


def test_reduced_precision_bounds_tiny():
    r = sweep("tiny", 8)["results"]
    # an FP32 stack feeding the FP64 update: ~1e-6 on the pose, no gate flips expected at this size
    assert not r["stack32"]["diverged"] and r["stack32"]["pos"] < 1e-4 and r["stack32"]["P_rel"] < 1e-3
    assert not r["stack_tf32"]["diverged"] and r["stack_tf32"]["pos"] < 1e-2
    # the parity bar of the FP64 path (1e-8) is out of reach for every reduced variant: that is the point of the sweep
    assert r["stack32"]["P_rel"] > 1e-10
