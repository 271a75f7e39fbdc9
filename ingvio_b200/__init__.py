"""ingvio_b200: B200-native (sm_100a CUDA) invariant-EKF hot path of InGVIO behind a C-ABI.

Product package. It never imports anything from `oracle/`; the CUDA library is mandatory
(`ingvio_b200.capi.load()` raises if libingvio_b200.so is missing -- there is no CPU fallback).
"""
__version__ = "0.1.0"
