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


def _run(wname, B, frames, with_oracle):
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st64, st32 = SyntheticStream(wl, B), SyntheticStream(wl, B)
    g64, g32 = make_gpu(wl, st64, fp), make_gpu(wl, st32, fp)
    g32.set_precision(capi.PREC_FP32_STACK)
    orc = make_oracles(wl, SyntheticStream(wl, B), fp) if with_oracle else None
    n_tracks = flips = 0
    worst_p = worst_P = 0.0
    for i in range(frames):
        fr64, fr32 = st64.next_frame(), st32.next_frame()
        o64 = g64.step(fr64, noise=fp.visual_noise, want=True)
        o32 = g32.step(fr32, noise=fp.visual_noise, want=True)
        if orc is not None:
            for b, f in enumerate(orc):
                f.step(fr64.seq(b))
        if "visual" in o64:
            assert g32.last_visual_path() == 1          # the materialised stack (that is what the mode is about)
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
    res = dict(workload=wname, frames=frames, tracks=n_tracks, flips=flips, pos=worst_p, P=worst_P)
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
