"""Ill-conditioned MSCKF stacks: the default compression forms [R | Q^T r] from the Gram matrix of the stack (pivot threshold
1e-13 relative, k_gram.cu), the reference uses a Householder QR (SPQR, RemoveLostUpdate.cpp:138-160). On stacks with weakly
observable directions -- every feature far away, so that the translation columns are almost a multiple of the rotation
columns, or a far / near mix over five decades of depth -- the Gram path, the Householder kernels and the oracle's LAPACK QR
must still give the same posterior at the parity bars: the update consumes R only through R^T R = H^T H, whose backward
error is eps ||H||^2 either way.  IGV_FLAG_WEAK_PIVOT reports that the Gram form met such a column."""
import numpy as np
import pytest

from helpers import assert_state_close, filter_params, make_gpu, make_oracles
from ingvio_b200 import capi
from ingvio_b200.synth import CAM_RATE, WORKLOADS, SyntheticStream

pytestmark = pytest.mark.gpu


def _rebuild_visual(st, fr, rng, lo, hi):
    """Replace the frame's tracks by ones at log-uniform depth in [lo, hi] metres (truth re-projected into every clone of
    the window, the same observation noise, triangulation error proportional to depth)."""
    B, F = fr.pf_w.shape[:2]
    ncl = int(fr.obs_mask[0, 0].sum())
    frames = [st.frame_idx - (ncl - 1) + s for s in range(ncl)]
    Rm, pm = st.cam_pose(frames[ncl // 2] / CAM_RATE)
    depth = np.exp(rng.uniform(np.log(lo), np.log(hi), (B, F)))
    xy = rng.uniform(-0.45, 0.45, (B, F, 2))
    pc = np.concatenate([xy * depth[..., None], depth[..., None]], -1)
    truth = np.einsum("bij,bfj->bfi", Rm, pc) + pm[:, None, :]
    fr.pf_w = truth + rng.normal(0.0, 1.0, (B, F, 3)) * (0.002 * depth[..., None])
    for s, k in enumerate(frames):
        Rc, pcw = st.cam_pose(k / CAM_RATE)
        q = np.einsum("bji,bfj->bfi", Rc, truth - pcw[:, None, :])
        fr.obs[:, :, s, 0:2] = q[..., 0:2] / q[..., 2:3] + rng.normal(0.0, st.obs_noise, (B, F, 2))


@pytest.mark.parametrize("depths", [(1e3, 1e5), (0.8, 1e5), (1e6, 1e8)], ids=["far", "mixed", "beyond_threshold"])
@pytest.mark.parametrize("wname", ["tiny", "c2"])
def test_gram_householder_oracle_agree(monkeypatch, wname, depths):
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    nfr = 7 if wname == "tiny" else 13
    runs = {}
    for name, cfg in (("gram", "30"), ("householder", "8")):
        monkeypatch.setenv("IGV_QR_CFG", cfg)
        rng = np.random.default_rng(77)
        st = SyntheticStream(wl, 2)
        orc = make_oracles(wl, st, fp) if name == "gram" else None
        g = make_gpu(wl, st, fp)
        for i in range(nfr):
            fr = st.next_frame()
            if fr.visual_mode is not None:
                _rebuild_visual(st, fr, rng, *depths)
            g.step(fr, noise=fp.visual_noise)
            if orc is not None:
                for b, f in enumerate(orc):
                    f.step(fr.seq(b))
                assert_state_close(g, orc, wl.sw, what=f"{wname} {depths} frame {i} (Gram path vs oracle)")
        fl = g.flags()
        assert np.all((fl & (capi.FLAG_NEG_DIAG | capi.FLAG_CHOL_FAIL)) == 0)
        runs[name] = (g.get_state().copy(), g.get_full_cov().copy(), fl.copy())
        g.close()
    xg, Pg, flg = runs["gram"]
    xh, Ph, flh = runs["householder"]
    assert np.max(np.abs(xg - xh)) <= 1e-9 * max(1.0, np.max(np.abs(xh)))
    for b in range(Pg.shape[0]):
        assert np.linalg.norm(Pg[b] - Ph[b]) <= 1e-8 * max(1.0, np.linalg.norm(Ph[b]))
    assert np.all((flh & capi.FLAG_WEAK_PIVOT) == 0), "the Householder path never raises the Gram flag"
    print(f"{wname} {depths}: weak-pivot flag on the Gram path: {(flg & capi.FLAG_WEAK_PIVOT) != 0}")
