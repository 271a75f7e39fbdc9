"""BASELINE configs[4] ("FP32 vs FP64 tolerance sweep") on the device: IGV_PREC_FP32_STACK stores the projected per-track
blocks [H | r] in single precision (the gate and the whole EKF update stay double) and must stay within the tolerances the
oracle-level sweep of round 1 predicted for "FP32 stack -> FP64 update" (profiles/r01_fp32_sweep.md): position 1e-4 m,
covariance 1e-6 relative, and no chi^2 decision differing from the FP64 run (here: at most 1 per 10^4 tracks, caused only by
the state drifting apart by the FP32 rounding of earlier frames)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from helpers import filter_params, gstep, make_gpu, make_oracles
from ingvio_b200 import capi
from ingvio_b200.synth import WORKLOADS, SyntheticStream

TOL_POS, TOL_P = 1e-4, 1e-6


def _run(wname, B, frames, with_oracle, mode=None):
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st64, st32 = SyntheticStream(wl, B), SyntheticStream(wl, B)
    g64, g32 = make_gpu(wl, st64, fp), make_gpu(wl, st32, fp)
    mode = capi.PREC_FP32_STACK if mode is None else mode
    g32.set_precision(mode)
    orc = make_oracles(wl, SyntheticStream(wl, B), fp) if with_oracle else None
    n_tracks = flips = tc_frames = 0
    worst_p = worst_P = pre_p = pre_P = 0.0
    for i in range(frames):
        fr64, fr32 = st64.next_frame(), st32.next_frame()
        o64 = g64.step(fr64, noise=fp.visual_noise, want=True)
        o32 = g32.step(fr32, noise=fp.visual_noise, want=True)
        if orc is not None:
            for b, f in enumerate(orc):
                f.step(fr64.seq(b))
        if "visual" in o64:
            assert g32.last_visual_path() == 1          # the materialised stack (that is what the mode is about)
            n1 = 6 * int(fr32.obs_mask[0, 0].sum()) + 1     # columns of the stack: the tensor-core kernel serves 129..192
            want_tc = 1 if (mode == capi.PREC_TF32_GRAM and 128 < n1 <= 192) else 0
            assert g32.last_gram_tensor() == want_tc
            tc_frames += want_tc
            gm64, gm32 = o64["visual"]["gamma"], o32["visual"]["gamma"]
            thr = np.array([g64_thr(fr64, fp, b) for b in range(B)])
            ok64 = gm64 < thr
            ok32 = gm32 < thr
            both = np.isfinite(gm64) & np.isfinite(gm32)
            n_tracks += int(both.sum())
            flips += int((ok64 != ok32)[both].sum())
        x64, x32 = g64.get_state(), g32.get_state()
        P64, P32 = g64.get_full_cov(), g32.get_full_cov()
        worst_p = max(worst_p, float(np.abs(x64[:, 9:12] - x32[:, 9:12]).max()))
        worst_P = max(worst_P, max(np.linalg.norm(P64[b] - P32[b]) / max(1.0, np.linalg.norm(P64[b])) for b in range(B)))
        if flips == 0:      # deviation from arithmetic alone: once a gate decision differs the runs hold different tracks
            pre_p, pre_P = worst_p, worst_P
    res = dict(workload=wname, frames=frames, tracks=n_tracks, flips=flips, pos=worst_p, P=worst_P, tc_frames=tc_frames,
               pos_before_first_flip=pre_p, P_before_first_flip=pre_P)
    if orc is not None:
        xo = np.array([f.pose()[1] for f in orc])
        res["pos_vs_oracle"] = float(np.abs(g32.get_state()[:, 9:12] - xo).max())
        res["P_vs_oracle"] = max(np.linalg.norm(g32.get_full_cov()[b] - f.cov()) / max(1.0, np.linalg.norm(f.cov())) for b, f in enumerate(orc))
    return res


def g64_thr(fr, fp, b):
    from scipy.stats import chi2
    dof = np.asarray(fr.obs_total[b], dtype=np.int64) - 1
    return chi2.ppf(fp.chi2_thres, np.maximum(dof, 1))


@pytest.mark.parametrize("wname,B,frames,with_oracle", [("c2", 2, 24, True), ("c5", 1, 8, False)])
def test_fp32_stack_tolerance(wname, B, frames, with_oracle):
    r = _run(wname, B, frames, with_oracle)
    print("precision sweep:", r)
    assert r["tracks"] > 0
    assert r["pos"] <= TOL_POS, r
    assert r["P"] <= TOL_P, r
    assert r["flips"] <= max(1, r["tracks"] // 10000), r
    if with_oracle:
        assert r["pos_vs_oracle"] <= TOL_POS and r["P_vs_oracle"] <= TOL_P, r


TOL_POS_TC, TOL_P_TC = 5e-2, 2e-3


@pytest.mark.parametrize("wname,B,frames", [("c5", 1, 40), ("c5", 2, 32)])
def test_tf32_gram_tolerance(wname, B, frames):
    """IGV_PREC_TF32_GRAM at the window it is meant for (c5, SW = 30: up to 181 columns): the Gram matrix of the float stack on
    the tcgen05 tensor cores (k_gram_tc.cuh: 3 x TF32 split, FP32 accumulation over 128 rows in TMEM, FP64 sums). The gate is
    still evaluated in double, so decisions only move through the state. The unit's FP32 accumulation truncates, which leaves
    ~1.5e-6 relative in the Gram matrix (tests/cuda/gram_tc_harness.cu); with 15 000+ rows per update that moves the posterior
    by up to 2.3 cm and 5e-4 relative in P within 32-40 frames (profiles/r02_tcgen05_eval.md) -- three orders of magnitude
    outside the bars of the FP32 stack. The mode is a throughput / accuracy trade-off, NOT a parity mode; the bars asserted
    here (5 cm, 2e-3, 2 gate decisions per 10^4 tracks) only pin that measured behaviour against regressions."""
    r = _run(wname, B, frames, False, mode=capi.PREC_TF32_GRAM)
    print("tf32 gram sweep:", r)
    assert r["tracks"] > 0 and r["tc_frames"] >= 8, r
    assert r["pos"] <= TOL_POS_TC, r
    assert r["P"] <= TOL_P_TC, r
    assert r["flips"] <= max(2, r["tracks"] // 5000), r


def test_tf32_gram_narrow_stack_falls_back():
    """Stacks of <= 128 columns (c2: 67) keep the FP64 accumulation of the float stack: same tolerances as the FP32 stack."""
    wl = WORKLOADS["c2"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    g = make_gpu(wl, st, fp)
    g.set_precision(capi.PREC_TF32_GRAM)
    for _ in range(5):
        g.step(st.next_frame(), noise=fp.visual_noise)
    assert g.last_visual_path() == 1 and g.last_gram_tensor() == 0
