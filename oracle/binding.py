"""TEST INFRASTRUCTURE (oracle) -- ctypes binding of oracle/_ref/liboracle.so.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``parpe_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from parpe_b200.tables import PB200_NSTATS, HierDesc, ProblemDesc, ProblemTables

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle.so")
ISM_SIMULTANEOUS, ISM_STAGGERED = 1, 2
LINSOL_DENSE, LINSOL_KLU = 1, 9

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle library missing: run `make -C oracle` (needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ora_last_error.restype = C.c_char_p
        L.ora_has_model.argtypes = [C.c_char_p]
        L.ora_create.restype = C.c_void_p
        L.ora_create.argtypes = [C.POINTER(ProblemDesc), C.c_int, C.c_int]
        L.ora_destroy.argtypes = [C.c_void_p]
        L.ora_evaluate.argtypes = [C.c_void_p, dp, ip, C.c_int, dp, dp, C.c_int, C.c_int, dp, dp, ip,
                                   C.POINTER(C.c_int64), dp]
        L.ora_get_model_outputs.argtypes = [C.c_void_p, dp, dp, dp, C.c_int]
        L.ora_debug_numsteps.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, ip, ip]
        L.ora_map_parameters.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.ora_plist.argtypes = [C.c_void_p, C.c_int, ip]
        for f in (L.ora_get_unscaled, L.ora_get_scaled):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_int]
        L.ora_apply_chain_rule.restype = C.c_double
        L.ora_apply_chain_rule.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.ora_hier_attach.argtypes = [C.c_void_p, C.POINTER(HierDesc)]
        L.ora_hier_num_outer.argtypes = [C.c_void_p]
        L.ora_hier_evaluate.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleProblem:
    """Reference AMICI/CVODES path + restated parPE glue on the CPU."""

    def __init__(self, tables: ProblemTables, ism=ISM_STAGGERED, linsol=LINSOL_KLU, model_name=None):
        self.tables = tables
        self._desc = tables.to_desc()
        if model_name is not None:
            self._mn = model_name.encode()
            self._desc.model_name = self._mn
        self._h = lib().ora_create(C.byref(self._desc), ism, linsol)
        if not self._h:
            raise RuntimeError("ora_create failed: %s" % lib().ora_last_error().decode())
        if tables.hier is not None:
            self._hd = tables.hier.to_desc()
            if lib().ora_hier_attach(self._h, C.byref(self._hd)) != 0:
                raise RuntimeError(lib().ora_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_destroy(self._h)
            self._h = None

    def evaluate(self, theta, idx=None, want_grad=True, nthreads=0, collect_stats=False):
        T = self.tables
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        idx = np.arange(T.n_conditions, dtype=np.int32) if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
        n = len(idx)
        fval = C.c_double()
        grad = np.zeros(T.n_opt) if want_grad else None
        llh = np.zeros(n)
        sllh = np.zeros((n, T.np_))
        status = np.zeros(n, dtype=np.int32)
        stats = np.zeros((n, PB200_NSTATS), dtype=np.int64)
        secs = np.zeros(n)
        rc = lib().ora_evaluate(self._h, _dp(theta), idx.ctypes.data_as(C.POINTER(C.c_int)), n, C.byref(fval),
                                _dp(grad), nthreads, int(collect_stats), _dp(llh), _dp(sllh),
                                status.ctypes.data_as(C.POINTER(C.c_int)),
                                stats.ctypes.data_as(C.POINTER(C.c_int64)), _dp(secs))
        if rc < 0:
            raise RuntimeError(lib().ora_last_error().decode())
        return dict(status=rc, fval=fval.value, grad=grad, llh=llh, sllh=sllh, cond_status=status, stats=stats,
                    seconds=secs)

    def get_model_outputs(self, theta, nthreads=0):
        T = self.tables
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        y = np.zeros_like(T.y)
        s = np.zeros_like(T.y)
        rc = lib().ora_get_model_outputs(self._h, _dp(theta), _dp(y), _dp(s), nthreads)
        if rc < 0:
            raise RuntimeError(lib().ora_last_error().decode())
        return rc, y, s

    def debug_numsteps(self, theta, c=0, want_grad=True):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n = self.tables.nt(c)
        ns, od = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        r = lib().ora_debug_numsteps(self._h, _dp(theta), c, int(want_grad), ns.ctypes.data_as(C.POINTER(C.c_int)),
                                     od.ctypes.data_as(C.POINTER(C.c_int)))
        return ns[:max(r, 0)], od[:max(r, 0)]

    def map_parameters(self, c, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        p = np.zeros(self.tables.np_)
        lib().ora_map_parameters(self._h, c, _dp(theta), _dp(p))
        return p

    def plist(self, c):
        pl = np.zeros(self.tables.np_, dtype=np.int32)
        n = lib().ora_plist(self._h, c, pl.ctypes.data_as(C.POINTER(C.c_int)))
        return pl[:n]

    def hier_num_outer(self):
        return lib().ora_hier_num_outer(self._h)

    def hier_evaluate(self, theta_outer, want_grad=True, nthreads=0):
        theta_outer = np.ascontiguousarray(theta_outer, dtype=np.float64)
        fval = C.c_double()
        grad = np.zeros(len(theta_outer)) if want_grad else None
        full = np.zeros(self.tables.n_opt)
        rc = lib().ora_hier_evaluate(self._h, _dp(theta_outer), C.byref(fval), _dp(grad), _dp(full), nthreads)
        if rc < 0:
            raise RuntimeError(lib().ora_last_error().decode())
        return dict(status=rc, fval=fval.value, grad=grad, full_parameters=full)
