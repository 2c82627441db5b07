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
        L.ora_set_preequilibration.argtypes = [C.c_void_p, C.POINTER(C.c_int32), dp, C.POINTER(C.c_int64),
                                               C.POINTER(C.c_int32)]
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
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.ora_hier_compute_one.restype = C.c_double
        L.ora_hier_compute_one.argtypes = [C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_int, i64p, dp, dp, C.c_int]
        L.ora_hier_apply_one.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, i32p, C.c_int, C.c_int, i64p, dp, C.c_int]
        L.ora_hier_nllh.restype = C.c_double
        L.ora_hier_nllh.argtypes = [C.c_int, i64p, dp, dp, dp]
        L.ora_hier_splice.argtypes = [dp, C.c_int, i32p, C.c_int, i32p, C.c_int, i32p, C.c_int, dp, dp, dp, dp]
        L.ora_hier_filter.argtypes = [dp, C.c_int, i32p, C.c_int, dp]
        for f in (L.ora_hier_default_scaling, L.ora_hier_default_offset):
            f.restype = C.c_double
            f.argtypes = [C.c_int]
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
        pre = tables.preeq_args()
        if pre is not None:
            flag, kp, off, flat = pre
            lib().ora_set_preequilibration(self._h, flag.ctypes.data_as(C.POINTER(C.c_int32)), _dp(kp),
                                           off.ctypes.data_as(C.POINTER(C.c_int64)),
                                           flat.ctypes.data_as(C.POINTER(C.c_int32)))

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


# ---- hierarchical free functions (restated reference functions) on plain lists ------------
def _flatten(mat):
    off = np.concatenate([[0], np.cumsum([len(r) for r in mat])]).astype(np.int64)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64) for r in mat]))
    return flat, off


def _rows(rows):
    return np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 3)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def hier_compute(kind, param_idx, rows, outputs, measurements, ny):
    """kind: 'scaling' | 'offset' | 'sigma' (computeAnalytical{Scalings,Offsets,Sigmas})."""
    r = _rows(rows)
    o, off = _flatten(outputs)
    m, _ = _flatten(measurements)
    n_param = int(r[:, 0].max()) + 1
    k = {"scaling": 0, "offset": 1, "sigma": 2}[kind]
    return lib().ora_hier_compute_one(k, param_idx, n_param, r.ctypes.data_as(C.POINTER(C.c_int32)), len(r),
                                      len(outputs), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(o), _dp(m), ny)


def hier_apply(mode, param_idx, value, rows, outputs, ny):
    """mode: 'scaling' | 'offset' (applyOptimalScaling / applyOptimalOffset); returns new outputs."""
    r = _rows(rows)
    o, off = _flatten(outputs)
    n_param = int(r[:, 0].max()) + 1
    lib().ora_hier_apply_one(0 if mode == "scaling" else 1, param_idx, float(value), n_param,
                             r.ctypes.data_as(C.POINTER(C.c_int32)), len(r), len(outputs),
                             off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(o), ny)
    return [o[off[c]:off[c + 1]].tolist() for c in range(len(outputs))]


def hier_nllh(measurements, outputs, sigmas):
    m, off = _flatten(measurements)
    o, _ = _flatten(outputs)
    s, _ = _flatten(sigmas)
    return lib().ora_hier_nllh(len(measurements), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(m), _dp(o), _dp(s))


def hier_splice(reduced, scaling_idx, offset_idx, sigma_idx, scalings, offsets, sigmas):
    red = np.ascontiguousarray(reduced, dtype=np.float64)
    si, oi, gi = _i32(scaling_idx), _i32(offset_idx), _i32(sigma_idx)
    sv, ov, gv = (np.ascontiguousarray(v, dtype=np.float64) for v in (scalings, offsets, sigmas))
    full = np.zeros(len(red) + len(si) + len(oi) + len(gi))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    lib().ora_hier_splice(_dp(red), len(red), P(si), len(si), P(oi), len(oi), P(gi), len(gi), _dp(sv), _dp(ov),
                          _dp(gv), _dp(full))
    return full


def hier_filter(values, sorted_excl):
    v = np.ascontiguousarray(values, dtype=np.float64)
    e = _i32(sorted_excl)
    out = np.zeros(len(v) - len(e))
    lib().ora_hier_filter(_dp(v), len(v), e.ctypes.data_as(C.POINTER(C.c_int32)), len(e), _dp(out))
    return out
