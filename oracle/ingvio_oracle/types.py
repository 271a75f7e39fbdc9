"""Type-indexed state variables with the left-invariant retraction.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/VecState.h:32-134, VecState.cpp:25-62,
PoseState.cpp:25-30,79-88,174-186 and AnchoredLandmark.cpp:227-243 (XYZ world
position only; the body-frame representations are out of scope).
Quaternion copies kept by the reference are redundant with the rotation matrix and
are not restated.
"""
import numpy as np

from .lie import gamma_func


class Type:
    """VecState.h:32-54: idx() = first row/col in the covariance, size() = error dim."""

    def __init__(self, size):
        self._size = size
        self._idx = -1

    def idx(self):
        return self._idx

    def size(self):
        return self._size

    def set_cov_idx(self, i):
        self._idx = i

    def update(self, dx):  # pragma: no cover - abstract
        raise NotImplementedError


class Vec3(Type):
    def __init__(self):
        super().__init__(3)
        self.vec = np.zeros(3)

    def value(self):
        return self.vec

    def set_value(self, v):
        self.vec = np.array(v, dtype=np.float64).reshape(3)

    def update(self, dx):
        """VecState.cpp:27-31."""
        self.vec = self.vec + dx[self._idx:self._idx + 3]


class Scalar(Type):
    def __init__(self):
        super().__init__(1)
        self.scalar = 0.0

    def value(self):
        return self.scalar

    def set_value(self, v):
        self.scalar = float(v)

    def update(self, dx):
        """VecState.cpp:43-47."""
        self.scalar = self.scalar + float(dx[self._idx])


class SO3(Type):
    def __init__(self):
        super().__init__(3)
        self.rot = np.eye(3)

    def update(self, dx):
        """PoseState.cpp:25-30: R <- Gamma0(dtheta) R."""
        self.rot = gamma_func(dx[self._idx:self._idx + 3], 0) @ self.rot


class SE3(Type):
    def __init__(self):
        super().__init__(6)
        self.rot = np.eye(3)
        self.vec = np.zeros(3)

    def value_linear(self):
        return self.rot

    def value_trans(self):
        return self.vec

    def set_value(self, R, p):
        self.rot = np.array(R, dtype=np.float64).reshape(3, 3)
        self.vec = np.array(p, dtype=np.float64).reshape(3)

    def update(self, dx):
        """PoseState.cpp:79-88."""
        th = dx[self._idx:self._idx + 3]
        G0 = gamma_func(th, 0)
        self.rot = G0 @ self.rot
        self.vec = G0 @ self.vec + gamma_func(th, 1) @ dx[self._idx + 3:self._idx + 6]


class SE23(Type):
    def __init__(self):
        super().__init__(9)
        self.rot = np.eye(3)
        self.vec1 = np.zeros(3)  # position
        self.vec2 = np.zeros(3)  # velocity

    def value_linear(self):
        return self.rot

    def value_trans1(self):
        return self.vec1

    def value_trans2(self):
        return self.vec2

    def update(self, dx):
        """PoseState.cpp:174-186."""
        th = dx[self._idx:self._idx + 3]
        G0 = gamma_func(th, 0)
        G1 = gamma_func(th, 1)
        self.rot = G0 @ self.rot
        self.vec1 = G0 @ self.vec1 + G1 @ dx[self._idx + 3:self._idx + 6]
        self.vec2 = G0 @ self.vec2 + G1 @ dx[self._idx + 6:self._idx + 9]


class AnchoredLandmark(Type):
    """World-xyz landmark tied to an anchor clone (AnchoredLandmark.cpp:84-100,227-243)."""

    def __init__(self):
        super().__init__(3)
        self.pos_xyz = np.zeros(3)
        self.anchored_pose = None

    def reset_anchored_pose(self, pose):
        self.anchored_pose = pose

    def get_anchored_pose(self):
        return self.anchored_pose

    def value_pos_xyz(self):
        return self.pos_xyz

    def set_value_pos_xyz(self, xyz):
        self.pos_xyz = np.array(xyz, dtype=np.float64).reshape(3)

    def update(self, dx):
        dp = dx[self._idx:self._idx + 3]
        if self.anchored_pose is not None:
            th = dx[self.anchored_pose.idx():self.anchored_pose.idx() + 3]
            self.pos_xyz = gamma_func(th, 0) @ self.pos_xyz + gamma_func(th, 1) @ dp
        else:
            self.pos_xyz = self.pos_xyz + dp
