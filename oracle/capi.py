"""ctypes access to oracle/_build/libtapir_oracle.so (quadpack_port.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtapir_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    src = os.path.join(_HERE, "quadpack_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_townsend_pi.restype = ctypes.c_double
        L.oracle_townsend_pi.argtypes = [ctypes.c_double, ctypes.c_double]
        L.oracle_qk21.restype = None
        L.oracle_qk21.argtypes = [ctypes.c_double] * 3 + [_dp] * 4
        L.oracle_qags.restype = ctypes.c_int
        L.oracle_qags.argtypes = [ctypes.c_double] * 5 + [ctypes.c_int, _dp, _dp, _ip, _ip]
        L.oracle_profile.restype = None
        L.oracle_profile.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        L.oracle_epochs.restype = None
        L.oracle_epochs.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_epoch_sites.restype = None
        L.oracle_epoch_sites.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_batch.restype = None
        L.oracle_batch.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]
        _lib = L
    return _lib


def qags(a, b, rate, epsabs=1.49e-8, epsrel=1.49e-8, limit=50):
    """-> (result, abserr, neval, last, ier) exactly as scipy.integrate.quad(full_output=1) reports."""
    res, err = ctypes.c_double(), ctypes.c_double()
    neval, last = ctypes.c_int(), ctypes.c_int()
    ier = lib().oracle_qags(a, b, rate, epsabs, epsrel, limit, res, err, neval, last)
    return res.value, err.value, neval.value, last.value, ier


def qk21(a, b, rate):
    out = [ctypes.c_double() for _ in range(4)]
    lib().oracle_qk21(a, b, rate, *out)
    return tuple(o.value for o in out)  # result, abserr, resabs, resasc


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def profile(rates, T):
    rates = _c(rates, np.float64)
    out = np.empty(T, dtype=np.float64)
    lib().oracle_profile(rates.ctypes.data, rates.size, T, out.ctypes.data)
    return out


def epochs(rates, epoch_list):
    rates = _c(rates, np.float64)
    ab = _c(np.asarray(epoch_list, dtype=np.float64).reshape(-1), np.float64)
    E = ab.size // 2
    si, se = np.empty(E), np.empty(E)
    nm = np.empty(E, dtype=np.int64)
    lib().oracle_epochs(rates.ctypes.data, rates.size, ab.ctypes.data, E, si.ctypes.data, se.ctypes.data,
                        nm.ctypes.data)
    return si, se, nm


def epoch_sites(rates, a, b):
    rates = _c(rates, np.float64)
    integ, err = np.empty(rates.size), np.empty(rates.size)
    last = np.empty(rates.size, dtype=np.int32)
    lib().oracle_epoch_sites(rates.ctypes.data, rates.size, float(a), float(b), integ.ctypes.data,
                             err.ctypes.data, last.ctypes.data)
    return integ, err, last


def batch(offsets, rates, T, epoch_list):
    """CSR batch -> (pi_net[L,T], integral[L,E], error[L,E]); single thread."""
    offsets = _c(offsets, np.int64)
    rates = _c(rates, np.float64)
    ab = _c(np.asarray(epoch_list, dtype=np.float64).reshape(-1), np.float64)
    L, E = offsets.size - 1, ab.size // 2
    pi_net = np.empty((L, T)); si = np.empty((L, E)); se = np.empty((L, E))
    lib().oracle_batch(L, offsets.ctypes.data, rates.ctypes.data, T, ab.ctypes.data, E, pi_net.ctypes.data,
                       si.ctypes.data, se.ctypes.data)
    return pi_net, si, se
