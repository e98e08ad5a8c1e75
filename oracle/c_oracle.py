"""ctypes wrapper of oracle/mbar_oracle.c (C (pthreads) restatement) — TEST INFRASTRUCTURE, NOT PRODUCT."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libmbar_oracle.so")
_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        if build and (not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "mbar_oracle.c"))):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        for name in ("mbar_oracle_self_consistent_update", "mbar_oracle_gradient"):
            getattr(_lib, name).argtypes = [dp, C.c_int64, C.c_int64, dp, dp, dp]
            getattr(_lib, name).restype = None
        _lib.mbar_oracle_threads.restype = C.c_int
    return _lib


def _call(name, u_kn, N_k, f_k):
    u = np.ascontiguousarray(u_kn, dtype=np.float64)
    N = np.ascontiguousarray(N_k, dtype=np.float64)
    f = np.ascontiguousarray(f_k, dtype=np.float64)
    out = np.empty(u.shape[0])
    dp = C.POINTER(C.c_double)
    getattr(load(), name)(u.ctypes.data_as(dp), u.shape[0], u.shape[1], N.ctypes.data_as(dp), f.ctypes.data_as(dp),
                          out.ctypes.data_as(dp))
    return out


def self_consistent_update(u_kn, N_k, f_k):
    return _call("mbar_oracle_self_consistent_update", u_kn, N_k, f_k)


def mbar_gradient(u_kn, N_k, f_k):
    return _call("mbar_oracle_gradient", u_kn, N_k, f_k)


def threads():
    return int(load().mbar_oracle_threads())
