"""The default compression is the Gram / Cholesky path (k_gram.cu, IGV_QR_CFG=30 forces it). The Householder QR compression has three kernels (multi-warp CTA per row range, single-warp DFMA streams, single-warp
DMMA panel streams) and a row-split + merge path. The dispatcher picks by batch size, so the parity suite's
small batches would only ever reach one of them: re-run the visual-update parity cases with each kernel
forced through the IGV_QR_CFG / IGV_QR_SPLIT knobs (k_qr.cu launch_qr)."""
import pytest

import test_gpu_parity as tp

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[("8", None), ("8", "3"), ("9", "2"), ("20", None), ("20", "3"), ("30", None), ("30", "3")],
                ids=["stream", "stream_split3", "cta_split2", "mma", "mma_split3", "gram", "gram_split3"])
def qr_variant(request, monkeypatch):
    cfg, split = request.param
    monkeypatch.setenv("IGV_QR_CFG", cfg)
    if split:
        monkeypatch.setenv("IGV_QR_SPLIT", split)
    return request.param


@pytest.mark.parametrize("wname", ["tiny", "tiny_stereo"])
def test_all_obs_frames(qr_variant, wname):
    tp.test_msckf_all_obs_frames(wname)


def test_ragged_outliers_and_cap(qr_variant):
    tp.test_msckf_ragged_outliers_and_cap()


@pytest.mark.parametrize("mode,stereo", [("keyframe", False), ("sw_marg", False), ("keyframe", True)])
def test_selected_modes(qr_variant, mode, stereo):
    tp.test_msckf_selected_modes(mode, stereo)


def test_c2_frames(qr_variant):
    tp.test_c2_frames_against_oracle()


@pytest.mark.parametrize("cfg", ["8", "20", "30"])
def test_c1_frames(monkeypatch, cfg):
    """c1 (SW=5: n+1 = 31 columns: one column slot / four column tiles) through the stream and DMMA kernels."""
    import numpy as np
    from helpers import assert_state_close, filter_params, gstep, make_gpu, make_oracles
    from ingvio_b200.synth import WORKLOADS, SyntheticStream
    monkeypatch.setenv("IGV_QR_CFG", cfg)
    wl = WORKLOADS["c1"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(9):
        fr = st.next_frame()
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
        assert_state_close(g, orc, wl.sw, what=f"c1 frame {i}")
    assert np.all((g.flags() & 3) == 0)


# ---- the fused kernel (Gram matrix accumulated inside k_msckf_features, no stack in HBM) ------------------------
# The dispatcher only fuses for batches that fill the chip (B >= 296); IGV_FUSE=1 forces it for the small parity
# batches. Every visual update of these runs must really have taken path 2 (igv_last_visual_path).
@pytest.fixture
def fused(monkeypatch):
    from ingvio_b200.filter import BatchFilter
    monkeypatch.setenv("IGV_FUSE", "1")
    monkeypatch.delenv("IGV_QR_CFG", raising=False)
    seen = []
    orig = BatchFilter.msckf_update

    def wrapped(self, *a, **k):
        out = orig(self, *a, **k)
        seen.append(self.last_visual_path())
        return out

    monkeypatch.setattr(BatchFilter, "msckf_update", wrapped)
    yield seen
    assert seen and all(p == 2 for p in seen), f"fused path not taken: {sorted(set(seen))}"


@pytest.mark.parametrize("wname", ["tiny", "tiny_stereo"])
def test_fused_all_obs_frames(fused, wname):
    tp.test_msckf_all_obs_frames(wname)


def test_fused_ragged_outliers_and_cap(fused):
    tp.test_msckf_ragged_outliers_and_cap()


@pytest.mark.parametrize("mode,stereo", [("keyframe", False), ("sw_marg", False), ("keyframe", True)])
def test_fused_selected_modes(fused, mode, stereo):
    tp.test_msckf_selected_modes(mode, stereo)


def test_fused_c2_frames(fused):
    tp.test_c2_frames_against_oracle()


def test_fused_batch_equals_singles(fused):
    tp.test_batch_equals_singles()


def test_fused_c1_frames(fused):
    _c1_frames()


def _c1_frames():
    import numpy as np
    from helpers import assert_state_close, filter_params, gstep, make_gpu, make_oracles
    from ingvio_b200.synth import WORKLOADS, SyntheticStream
    wl = WORKLOADS["c1"]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 2)
    orc = make_oracles(wl, st, fp)
    g = make_gpu(wl, st, fp)
    for i in range(9):
        fr = st.next_frame()
        gstep(g, fr, fp)
        for b, f in enumerate(orc):
            f.step(fr.seq(b))
        assert_state_close(g, orc, wl.sw, what=f"c1 frame {i}")
    assert np.all((g.flags() & 3) == 0)


def test_column_factor_kernel_c2(monkeypatch):
    """k_gram_factor (column by column, packed storage: the fallback of wide windows) on the c2 frames; the default
    for n <= 160 is the blocked k_gram_factor_blocked."""
    monkeypatch.setenv("IGV_FACTOR_CFG", "1")
    tp.test_c2_frames_against_oracle()
