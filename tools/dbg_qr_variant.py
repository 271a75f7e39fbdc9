import os, sys
os.environ["IGV_QR_CFG"] = sys.argv[1]
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle"); sys.path.insert(0, ".")
import numpy as np
from helpers import filter_params, make_gpu
from ingvio_b200.synth import WORKLOADS, SyntheticStream
wl = WORKLOADS[sys.argv[2]]
fp = filter_params(wl)
st = SyntheticStream(wl, 2)
g = make_gpu(wl, st, fp)
res = []
for i in range(8):
    fr = st.next_frame()
    g.step(fr, noise=fp.visual_noise)
    res.append((g.get_state().copy(), g.get_full_cov().copy()))
np.save("gpurun_out/dbg_%s_%s.npy" % (sys.argv[1], sys.argv[2]), np.array([np.concatenate([x.ravel(), p.ravel()]) for x, p in res], dtype=object), allow_pickle=True)
