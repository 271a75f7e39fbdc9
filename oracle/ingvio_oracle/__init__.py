"""CPU oracle for the InGVIO invariant-EKF hot path.

TEST INFRASTRUCTURE ONLY. This package restates, in numpy FP64, the algorithm of
/root/reference/ingvio_estimator/src/{AuxGammaFunc,PoseState,VecState,State,StateManager,
ImuPropagator,Update,RemoveLostUpdate,KeyframeUpdate,SwMargUpdate,GnssUpdate,GnssManager}.cpp.
It may be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product package `ingvio_b200`.

Parity pin status (see oracle/README.md): the reference ships no golden vectors; the oracle is pinned
by the reference's own property tests (tests/test_oracle_reference_properties.py restates
TestStateManager.cpp / TestPropagator.cpp 1:1). The null-space projection, the SPQR compression, the
chi^2 gate, the three visual updaters and the GNSS update are exercised by NO reference test:
for those the parity is UNPINNED at the reference level.
"""
from .lie import gamma_func, psi1_func, psi2_func, skew, vee  # noqa: F401
from .state import BDS, FS, GAL, GLO, GPS, YOF, FilterParams, State, StateParams  # noqa: F401
from .state_manager import StateManager  # noqa: F401
from .types import SE3, SE23, SO3, AnchoredLandmark, Scalar, Type, Vec3  # noqa: F401
from .imu_propagator import ImuCtrl, ImuPropagator  # noqa: F401
from .visual_update import (FeatureInfo, KeyframeUpdate, RemoveLostUpdate, SwMargUpdate,  # noqa: F401
                            UpdateBase)
from .gnss_update import GnssEpoch, GnssUpdate, calc_R_w2enu, dot_R_w2enu  # noqa: F401
from .triangulator import TriParams, Triangulator  # noqa: F401
from .frame import OracleFilter  # noqa: F401
from .landmark_update import LandmarkUpdate  # noqa: F401,E402
