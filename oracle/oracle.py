"""ctypes binding of the CPU oracle (oracle/cosmo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The oracle shares nothing with the product except the `cl_spec` struct layout of include/cosmolike.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cosmology_model_fit_b200.spec import ClSpec, LikelihoodSpec, OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcosmo_oracle.so")
_dp = C.POINTER(C.c_double)


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "cosmo_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_LIB_PATH)
        sp = C.POINTER(ClSpec)
        i64 = C.c_int64
        lib.oracle_eval.argtypes = [sp, _dp, i64, i64, C.c_int, _dp, _dp, C.c_int, C.c_int]
        lib.oracle_distances.argtypes = [sp, _dp, i64, i64, _dp, i64, _dp, _dp]
        lib.oracle_bao_theory.argtypes = [sp, _dp, i64, i64, _dp]
        lib.oracle_cmb.argtypes = [sp, _dp, i64, i64, _dp]
        lib.oracle_sn_residuals.argtypes = [sp, _dp, i64, i64, _dp]
        lib.oracle_interp_hermite.argtypes = [_dp, C.c_int, _dp, _dp, _dp, C.c_int, _dp]
        lib.oracle_interp_pchip.argtypes = [_dp, C.c_int, _dp, _dp, C.c_int, _dp]
        lib.oracle_pchip_slopes.argtypes = [_dp, _dp, C.c_int, _dp]
        lib.oracle_solve_triangular.argtypes = [_dp, _dp, C.c_int, _dp]
        lib.oracle_solve_triangular.restype = C.c_double
        lib.oracle_solve_triangular_ld.argtypes = [_dp, _dp, i64, C.c_int, _dp]
        lib.oracle_max_threads.restype = C.c_int
        for f in (lib.oracle_interp_hermite, lib.oracle_interp_pchip, lib.oracle_pchip_slopes):
            f.restype = None
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _theta(theta, ndim):
    t = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
    if t.shape[1] != ndim:
        raise ValueError(f"theta must have {ndim} columns")
    return t


def max_threads():
    return int(_load().oracle_max_threads())


class Oracle:
    """Scalar CPU evaluation of one LikelihoodSpec, method names as in the reference scripts."""

    def __init__(self, spec: LikelihoodSpec, grid_builds: int = 1):
        self.spec = spec
        self.c = spec.c_spec()
        self.lib = _load()
        self.grid_builds = grid_builds

    def _eval(self, theta, what, nthreads=1, comps=False):
        t = _theta(theta, self.spec.ndim)
        out = np.empty(t.shape[0])
        cm = np.empty((t.shape[0], 4)) if comps else None
        rc = self.lib.oracle_eval(C.byref(self.c), _p(t), t.shape[0], t.shape[1], what, _p(out),
                                  _p(cm) if comps else None, nthreads, self.grid_builds)
        if rc != 0:
            raise RuntimeError(f"oracle_eval failed: {rc}")
        return (out, cm) if comps else out

    def chi_squared(self, theta, nthreads=1):
        return self._eval(theta, OUT_CHI2, nthreads)

    def log_likelihood(self, theta, nthreads=1):
        return self._eval(theta, OUT_LOGLIKE, nthreads)

    def log_probability(self, theta, nthreads=1):
        return self._eval(theta, OUT_LOGPROB, nthreads)

    def components(self, theta):
        return self._eval(theta, OUT_CHI2, 1, comps=True)[1]

    def distances(self, theta, zq):
        t = _theta(theta, self.spec.ndim)
        zq = np.ascontiguousarray(zq, dtype=np.float64)
        dm = np.empty((t.shape[0], zq.size)); dh = np.empty((t.shape[0], zq.size))
        self.lib.oracle_distances(C.byref(self.c), _p(t), t.shape[0], t.shape[1], _p(zq), zq.size, _p(dm), _p(dh))
        return dm, dh

    def bao_theory(self, theta):
        t = _theta(theta, self.spec.ndim)
        out = np.empty((t.shape[0], self.c.n_bao))
        self.lib.oracle_bao_theory(C.byref(self.c), _p(t), t.shape[0], t.shape[1], _p(out))
        return out

    def cmb(self, theta):
        t = _theta(theta, self.spec.ndim)
        out = np.empty((t.shape[0], 8))
        self.lib.oracle_cmb(C.byref(self.c), _p(t), t.shape[0], t.shape[1], _p(out))
        return out

    def sn_residuals(self, theta):
        t = _theta(theta, self.spec.ndim)
        out = np.empty((t.shape[0], self.c.n_sn))
        self.lib.oracle_sn_residuals(C.byref(self.c), _p(t), t.shape[0], t.shape[1], _p(out))
        return out


def interp_hermite(xq, x, y, yp):
    xq, x, y, yp = (np.ascontiguousarray(a, dtype=np.float64) for a in (xq, x, y, yp))
    out = np.empty_like(xq)
    _load().oracle_interp_hermite(_p(xq), xq.size, _p(x), _p(y), _p(yp), x.size, _p(out))
    return out


def interp_pchip(xq, x, y):
    xq, x, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (xq, x, y))
    out = np.empty_like(xq)
    _load().oracle_interp_pchip(_p(xq), xq.size, _p(x), _p(y), x.size, _p(out))
    return out


def pchip_slopes(x, y):
    x, y = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y))
    d = np.empty_like(x)
    _load().oracle_pchip_slopes(_p(x), _p(y), x.size, _p(d))
    return d


def solve_triangular(L, b):
    L, b = np.ascontiguousarray(L, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    y = np.empty_like(b)
    return float(_load().oracle_solve_triangular(_p(L), _p(b), b.size, _p(y)))


def solve_triangular_ld(L, R):
    """|L^-1 R[b]|^2 per row by forward substitution in 80-bit extended precision (the yardstick of the conditioning tests)."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    R = np.ascontiguousarray(np.atleast_2d(R), dtype=np.float64)
    out = np.empty(R.shape[0])
    if _load().oracle_solve_triangular_ld(_p(L), _p(R), R.shape[0], L.shape[0], _p(out)) != 0:
        raise MemoryError
    return out


def solve_triangular_batch(L, R):
    """The reference's FP64 forward substitution (solve_triangular.py:5-14) for every row of R."""
    return np.array([solve_triangular(L, r) for r in np.atleast_2d(R)])
