"""CPU: the oracle filter driven by tracker messages through the oracle MapServer stays a healthy filter (the frame
driver the GPU parity test tests/test_gpu_tracks.py compares against)."""
import numpy as np
import pytest

from helpers import filter_params, make_oracles
from ingvio_b200.synth import SyntheticStream, TrackerStream, Workload
from track_frames import OracleFrontEnd


@pytest.mark.parametrize("keyframe,stereo", [(False, False), (True, False), (False, True)])
def test_oracle_front_end(keyframe, stereo):
    SW = 5
    wl = Workload("trk", 11 + int(stereo), SW + (0 if keyframe else 1), 32, 0, stereo=stereo)
    fp = filter_params(wl, max_sw_clones=SW, frame_select_interval=2)
    st = SyntheticStream(wl, 2)
    trk = TrackerStream(st, 18, 32)
    fes = [OracleFrontEnd(f, keyframe) for f in make_oracles(wl, st, fp, with_gnss=False)]
    for k in range(14):
        st.n_clones = 0
        fr = st.next_frame(with_visual=False, with_gnss=False, marg_oldest=False)
        n, ids, uv = trk.message(fr.t)
        for b, fe in enumerate(fes):
            fe.frame(fr.seq(b), int(n[b]), ids[b], uv[b])
    for b, fe in enumerate(fes):
        P = fe.f.cov()
        assert np.allclose(P, P.T, atol=1e-12) and np.linalg.eigvalsh(P).min() > -1e-12
        assert len(fe.f.state.sw_camleft_poses) <= wl.sw
        assert fe.counts["sel"] > 0 and (keyframe or fe.counts["lost"] > 0), fe.counts   # (two clones leave per
        # keyframe step, so lost tracks rarely keep the 4 observations RemoveLostUpdate asks for)
        # position error stays bounded (the filter is fed consistent measurements)
        tt = np.zeros(st.B) + fe.f.state.timestamp
        assert np.linalg.norm(fe.f.state.extended_pose.vec1 - st.traj.pos(tt)[b]) < 0.05
