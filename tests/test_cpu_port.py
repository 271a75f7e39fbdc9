"""The oracle's C++ CPU port (bench.py's timed CPU baseline) against the numpy oracle. CPU only."""
import numpy as np
import pytest

import ingvio_oracle as o
from cpu_port.port import CpuPortFilter
from helpers import GNSS_INIT, cov_diag21, filter_params, make_oracles, oracle_packed_state
from ingvio_b200.filter import chi2_table
from ingvio_b200.synth import WORKLOADS, SyntheticStream


def make_port(wl, st, fp, b=0):
    sp = o.StateParams(fp)
    f = CpuPortFilter([sp.noise_g, sp.noise_a, sp.noise_bg, sp.noise_ba, sp.noise_clockbias, sp.noise_cb_rw],
                      [0, 0, -fp.gravity_norm], (fp.T_cl2cr_R, fp.T_cl2cr_p), wl.stereo, chi2_table(160, fp.chi2_thres))
    ini = st.initial_state()
    f.init(ini["R"][b], ini["p"][b], ini["v"][b], ini["bg"][b], ini["ba"][b], fp.T_cl2i_R, fp.T_cl2i_p, cov_diag21(fp))
    if wl.sats > 0:
        for g, val, cov in GNSS_INIT:
            f.add_gnss(g, val, cov)
    return f


@pytest.mark.parametrize("wname,frames", [("tiny", 9), ("tiny_stereo", 8), ("c1", 8)])
def test_port_matches_numpy_oracle(wname, frames):
    wl = WORKLOADS[wname]
    fp = filter_params(wl)
    st = SyntheticStream(wl, 1)
    orc = make_oracles(wl, st, fp)[0]
    port = make_port(wl, st, fp)
    for i in range(frames):
        fr = st.next_frame()
        orc.step(fr.seq(0))
        port.step(fr.seq(0), fp.visual_noise)
        Po, Pp = orc.cov(), port.cov()
        assert Po.shape == Pp.shape
        assert np.linalg.norm(Po - Pp) <= 1e-8 * max(1, np.linalg.norm(Po)), (wname, i)
        xo = oracle_packed_state(orc, wl.sw)[:39 + 12 * len(orc.state.sw_camleft_poses)]
        assert np.max(np.abs(xo - port.state()) / np.maximum(1, np.abs(xo))) <= 1e-9, (wname, i)
        if fr.visual_mode is not None:
            assert port.n_accepted() == sum(1 for x in orc.last["gammas"] if x[3])
