"""so(3) helpers and the Gamma / Psi integrals of the invariant EKF.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/AuxGammaFunc.cpp:28-226.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math

import numpy as np


def skew(v):
    """AuxGammaFunc.cpp:28-35."""
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def vee(m):
    """AuxGammaFunc.cpp:37-44."""
    return np.array([m[2, 1], m[0, 2], m[1, 0]])


def gamma_func(vec, m):
    """Gamma_m(phi) = sum_k phi^k/(k+m)!  (AuxGammaFunc.cpp:46-113; small-angle cut 1e-6)."""
    assert 0 <= m <= 3
    vec = np.asarray(vec, dtype=np.float64).reshape(3)
    theta = float(np.linalg.norm(vec))
    if abs(theta) < 1e-6:
        factor = {3: 1.0 / 6.0, 2: 0.5}.get(m, 1.0)
        return factor * np.eye(3)
    n = vec / theta
    nx = skew(n)
    nx2 = nx @ nx
    s, c = math.sin(theta), math.cos(theta)
    if m == 1:
        f0, f1, f2 = 1.0, (1.0 - c) / theta, (theta - s) / theta
    elif m == 2:
        f0 = 0.5
        f1 = (theta - s) / theta ** 2
        f2 = (theta ** 2 + 2.0 * c - 2.0) / (2.0 * theta ** 2)
    elif m == 3:
        f0 = 1.0 / 6.0
        t3 = theta ** 3
        f1 = (theta ** 2 + 2.0 * c - 2.0) / (2.0 * t3)
        f2 = (t3 - 6.0 * theta + 6.0 * s) / (6.0 * t3)
    else:
        f0, f1, f2 = 1.0, s, 1.0 - c
    return f0 * np.eye(3) + f1 * nx + f2 * nx2


def _w_products(w, a):
    W, A = skew(w), skew(a)
    WA = W @ A
    WAW = WA @ W
    WAW2 = WAW @ W
    W2A = W @ WA
    W2AW = W2A @ W
    W2AW2 = W2AW @ W
    return WA, WAW, WAW2, W2A, W2AW, W2AW2


def psi1_func(w, a, dt):
    """AuxGammaFunc.cpp:115-166 (cut-off 1e-8 on |w dt|)."""
    w = np.asarray(w, dtype=np.float64).reshape(3)
    a = np.asarray(a, dtype=np.float64).reshape(3)
    if np.linalg.norm(w * dt) < 1e-8:
        return np.zeros((3, 3))
    M1 = skew(a) @ gamma_func(-w * dt, 2) * dt ** 2
    WA, WAW, WAW2, W2A, W2AW, W2AW2 = _w_products(w, a)
    eta = float(np.linalg.norm(w))
    xi = eta * dt
    xi2 = xi ** 2
    s1, c1_ = math.sin(xi), math.cos(xi)
    s2, c2_ = math.sin(2 * xi), math.cos(2 * xi)
    eta3 = eta ** 3
    eta4 = eta * eta3
    eta5 = eta * eta4
    eta6 = eta * eta5
    c1 = (s1 - xi * c1_) / eta3
    c2 = (c2_ - 4 * c1_ + 3) / (4 * eta4)
    c3 = (4 * s1 + s2 - 4 * xi * c1_ - 2 * xi) / (4 * eta5)
    c4 = (xi2 - 2 * xi * s1 - 2 * c1_ + 2) / (2 * eta4)
    c5 = (6 * xi - 8 * s1 + s2) / (4 * eta5)
    c6 = (2 * xi2 - 4 * xi * s1 - c2_ + 1) / (4 * eta6)
    return M1 @ (c1 * WA + c2 * WAW + c3 * WAW2 + c4 * W2A + c5 * W2AW + c6 * W2AW2)


def psi2_func(w, a, dt):
    """AuxGammaFunc.cpp:168-225 (cut-off 1e-7 on |w dt|)."""
    w = np.asarray(w, dtype=np.float64).reshape(3)
    a = np.asarray(a, dtype=np.float64).reshape(3)
    if np.linalg.norm(w * dt) < 1e-7:
        return np.zeros((3, 3))
    M1 = skew(a) @ gamma_func(-w * dt, 3) * dt ** 3
    WA, WAW, WAW2, W2A, W2AW, W2AW2 = _w_products(w, a)
    eta = float(np.linalg.norm(w))
    xi = eta * dt
    xi2 = xi ** 2
    xi3 = xi * xi2
    s1, c1_ = math.sin(xi), math.cos(xi)
    s2, c2_ = math.sin(2 * xi), math.cos(2 * xi)
    eta3 = eta ** 3
    eta4 = eta * eta3
    eta5 = eta * eta4
    eta6 = eta * eta5
    eta7 = eta * eta6
    c1 = (xi * s1 + 2 * c1_ - 2) / eta4
    c2 = (6 * xi - 8 * s1 + s2) / (8 * eta5)
    c3 = (2 * xi2 + 8 * xi * s1 + 16 * c1_ + c2_ - 17) / (8 * eta6)
    c4 = (xi3 + 6 * xi - 12 * s1 + 6 * xi * c1_) / (6 * eta5)
    c5 = (6 * xi2 + 16 * c1_ - c2_ - 15) / (8 * eta6)
    c6 = (4 * xi3 + 6 * xi - 24 * s1 - 3 * s2 + 24 * xi * c1_) / (24 * eta7)
    return M1 @ (c1 * WA + c2 * WAW + c3 * WAW2 + c4 * W2A + c5 * W2AW + c6 * W2AW2)
