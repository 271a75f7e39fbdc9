"""Pins the oracle's Triangulator with the reference's own test (test/TestTriangulator.cpp:133-177):
triangulating (1,2,3) from 10 views with sigma=0.02 pixel noise within 0.05 m, and from the 6 views
looking along +-x/+-y/+-z within 0.15 m, mono and stereo."""
import numpy as np

import ingvio_oracle as o

PF = np.array([1.0, 2.0, 3.0])
T_CL2CR = (np.eye(3), np.array([0.001, -0.12, 0.003]))


def rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def from_two_vectors(a, b):
    """Eigen::Quaterniond::FromTwoVectors(a, b).toRotationMatrix() for unit vectors."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    c = float(a @ b)
    if c < -1 + 1e-12:   # Eigen: opposite vectors -> rotation by pi about a singular vector of [a;b]
        _, _, vt = np.linalg.svd(np.vstack([a, b]))
        ax = vt[2]
        return 2 * np.outer(ax, ax) - np.eye(3)
    v = np.cross(a, b)
    K = o.skew(v)
    return np.eye(3) + K + K @ K / (1 + c)


def views(rng, poses, stereo):
    obs = []
    for R, p in poses:
        body = R.T @ (PF - p)
        z = [body[0] / body[2] + rng.normal(0, 0.02), body[1] / body[2] + rng.normal(0, 0.02)]
        if stereo:
            br = T_CL2CR[0] @ body + T_CL2CR[1]
            z += [br[0] / br[2] + rng.normal(0, 0.02), br[1] / br[2] + rng.normal(0, 0.02)]
        obs.append(np.array(z))
    return obs


def poses1(rng):
    return [(rotz(rng.normal(0, 0.1)), np.array([2 * i - 9.0, 2 * i - 9.0, 0.0])) for i in range(10)]


def poses2():
    z = [0, 0, 1]
    spec = [([0, 0, 1], [0, 0, 0]), ([0, -1, 0], [0, 5, 0]), ([0, 0, -1], [0, 0, 8]), ([0, 1, 0], [0, -6, 0]),
            ([-1, 0, 0], [5.5, 0, 0]), ([1, 0, 0], [-10, 0, 0])]
    return [(from_two_vectors(z, d), np.array(p, float)) for d, p in spec]


def test_mono_triangulate():
    tri = o.Triangulator()
    for seed in range(5):
        rng = np.random.default_rng(seed)
        p1 = poses1(rng)
        ok, res = tri.triangulate_mono(views(rng, p1, False), p1)
        assert ok and np.linalg.norm(res - PF) < 0.05
        p2 = poses2()
        ok, res = tri.triangulate_mono(views(rng, p2, False), p2)
        assert ok and np.linalg.norm(res - PF) < 0.15


def test_stereo_triangulate():
    """The Huber-weighted LM converges linearly; with the default 10 outer iterations and conv_precision
    5e-7 some noise draws do not converge and the reference returns false for them (Triangulator.cpp:281-282).
    Like upstream (one fixed rand() sequence) this checks accuracy on draws that converge and requires most to."""
    tri = o.Triangulator()
    n_ok = 0
    for seed in range(10):
        rng = np.random.default_rng(100 + seed)
        p1 = poses1(rng)
        ok, res = tri.triangulate_stereo(views(rng, p1, True), p1, T_CL2CR)
        if ok:
            n_ok += 1
            assert np.linalg.norm(res - PF) < 0.05
        p2 = poses2()
        ok, res = tri.triangulate_stereo(views(rng, p2, True), p2, T_CL2CR)
        if ok:
            n_ok += 1
            assert np.linalg.norm(res - PF) < 0.15
    assert n_ok >= 12
    # with more iterations every draw converges to the same accuracy
    slow = o.Triangulator(o.TriParams(outer_loop_max_iter=40))
    rng = np.random.default_rng(101)
    p1 = poses1(rng)
    ok, res = slow.triangulate_stereo(views(rng, p1, True), p1, T_CL2CR)
    assert ok and np.linalg.norm(res - PF) < 0.05


def test_rejections():
    tri = o.Triangulator()
    rng = np.random.default_rng(7)
    p1 = poses1(rng)
    obs = views(rng, p1, False)
    assert not tri.triangulate_mono(obs[:4], p1[:4])[0]                       # <= 4 views (:183)
    still = [(np.eye(3), np.array([0.0, 0.0, 0.0]) + 1e-3 * i) for i in range(6)]
    assert not tri.triangulate_mono(views(rng, still, False), still)[0]       # baseline below trans_thres (:192)
    far = o.Triangulator(o.TriParams(max_depth=2.0))
    assert not far.triangulate_mono(obs, p1)[0]                               # depth gate (:306)
