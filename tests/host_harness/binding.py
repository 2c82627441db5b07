"""TEST INFRASTRUCTURE -- ctypes binding of the host build of the integrator core."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libhost_sim.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "host_sim.cpp")
    deps = [src, os.path.join(_HERE, "..", "..", "parpe_b200", "csrc", "bdf_core.cuh")]
    mdir = os.path.join(_HERE, "..", "..", "parpe_b200", "csrc", "models")
    deps += [os.path.join(mdir, f) for f in os.listdir(mdir) if not os.path.isdir(os.path.join(mdir, f))]
    idir = os.path.join(mdir, "imported")
    if os.path.isdir(idir):
        deps += [os.path.join(idir, f) for f in os.listdir(idir)]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.hs_simulate.argtypes = [C.c_char_p, dp, dp, ip, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double,
                                  C.c_double, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int64), dp, dp, ip]
        L.hs_simulate_ss.argtypes = [C.c_char_p, dp, dp, dp, C.c_uint, ip, C.c_int, C.c_int, dp, dp, dp, C.c_double,
                                     C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int64),
                                     dp, dp, ip, dp, C.c_int]
        _lib = L
    return _lib


def simulate_ss(model, p_unscaled, k, plist, t, my, sd, k_preeq=None, reinit_idx=(), t0=0.0, rtol=1e-8, atol=1e-12,
                maxsteps=10000, stldet=1, errconS=1, want_outputs=False, nx=None, newton_maxsteps=0):
    """The steady-state instantiation of the core: t = inf output times (post-equilibration) and, with
    ``k_preeq``, pre-equilibration + re-initialisation of the states ``reinit_idx``."""
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    p_unscaled, k, t, my, sd = d(p_unscaled), d(k), d(t), d(my).reshape(-1), d(sd).reshape(-1)
    kp = d(k_preeq) if k_preeq is not None else None
    plist = np.ascontiguousarray(plist, dtype=np.int32)
    ns = len(plist)
    llh = C.c_double()
    sllh = np.zeros(max(ns, 1))
    stats = np.zeros(20, dtype=np.int64)
    y = np.zeros(my.size) if want_outputs else None
    s = np.zeros(my.size) if want_outputs else None
    trace = np.zeros(2 * len(t), dtype=np.int32)
    xpre = np.zeros(64 * 64)
    mask = 0
    for i in reinit_idx:
        mask |= 1 << int(i)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
    st = lib().hs_simulate_ss(model.encode(), P(p_unscaled), P(k), P(kp), mask,
                              plist.ctypes.data_as(C.POINTER(C.c_int)), ns, len(t), P(t), P(my), P(sd), t0, rtol, atol,
                              maxsteps, stldet, errconS, C.byref(llh), P(sllh),
                              stats.ctypes.data_as(C.POINTER(C.c_int64)), P(y), P(s),
                              trace.ctypes.data_as(C.POINTER(C.c_int)), P(xpre), newton_maxsteps)
    return dict(numsteps=trace[0::2], order=trace[1::2], status=st, llh=llh.value, sllh=sllh[:ns], stats=stats, y=y,
                sigmay=s, x_preeq=xpre, nst_preeq=int(stats[16]))


def simulate(model, p_unscaled, k, plist, t, my, sd, t0=0.0, rtol=1e-8, atol=1e-12, maxsteps=10000, stldet=1,
             errconS=1, want_outputs=False):
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    p_unscaled, k, t, my, sd = d(p_unscaled), d(k), d(t), d(my).reshape(-1), d(sd).reshape(-1)
    plist = np.ascontiguousarray(plist, dtype=np.int32)
    ns = len(plist)
    llh = C.c_double()
    sllh = np.zeros(max(ns, 1))
    stats = np.zeros(20, dtype=np.int64)
    y = np.zeros(my.size) if want_outputs else None
    s = np.zeros(my.size) if want_outputs else None
    trace = np.zeros(2 * len(t), dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
    st = lib().hs_simulate(model.encode(), P(p_unscaled), P(k), plist.ctypes.data_as(C.POINTER(C.c_int)), ns, len(t),
                           P(t), P(my), P(sd), t0, rtol, atol, maxsteps, stldet, errconS, C.byref(llh), P(sllh),
                           stats.ctypes.data_as(C.POINTER(C.c_int64)), P(y), P(s), trace.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(numsteps=trace[0::2], order=trace[1::2], status=st, llh=llh.value, sllh=sllh[:ns], stats=stats, y=y, sigmay=s)


# ---- large-model core (bdf_team.cuh with HostTeam) ------------------------------------------
LIB_TEAM = os.path.join(_HERE, "libhost_team.so")
_lib_team = None


def build_team(force=False):
    src = os.path.join(_HERE, "host_sim_team.cpp")
    cs = os.path.join(_HERE, "..", "..", "parpe_b200", "csrc")
    deps = [src] + [os.path.join(cs, f) for f in ("bdf_team.cuh", "bdf_core.cuh", "net_model.cuh", "team.cuh", "pb_hd.h",
                                                  "sparse_lu.h")]
    mdir = os.path.join(cs, "models")
    deps += [os.path.join(mdir, f) for f in os.listdir(mdir) if f.startswith("net_")]
    if force or not os.path.exists(LIB_TEAM) or any(os.path.getmtime(d) > os.path.getmtime(LIB_TEAM) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB_TEAM, src])


def lib_team():
    global _lib_team
    if _lib_team is None:
        build_team()
        L = C.CDLL(LIB_TEAM)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.hs_team_simulate.argtypes = [C.c_char_p, dp, dp, ip, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int64),
                                       dp, dp, ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hs_team_simulate_ss.argtypes = L.hs_team_simulate.argtypes + [dp, ip, C.c_int]
        L.hs_team_solve.argtypes = [C.c_char_p, dp, C.c_double, dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, ip]
        _lib_team = L
    return _lib_team


def simulate_team(model, p_unscaled, k, plist, t, my, sd, t0=0.0, rtol=1e-8, atol=1e-12, maxsteps=10000, stldet=1,
                  errconS=1, want_outputs=False, panel_width=0, sparse=False, deg_max=0, stage_doubles=0,
                  fuse_history=True, fuse_newton=True, steady=False, k_preeq=None, reinit_idx=()):
    """``steady``: the steady-state instantiation of the core (t = inf output times; with ``k_preeq`` a
    pre-equilibration followed by the re-initialisation of the states ``reinit_idx``)."""
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    p_unscaled, k, t, my, sd = d(p_unscaled), d(k), d(t), d(my).reshape(-1), d(sd).reshape(-1)
    plist = np.ascontiguousarray(plist, dtype=np.int32)
    ns = len(plist)
    llh = C.c_double()
    sllh = np.zeros(max(ns, 1))
    stats = np.zeros(20, dtype=np.int64)
    y = np.zeros(my.size) if want_outputs else None
    s = np.zeros(my.size) if want_outputs else None
    trace = np.zeros(2 * len(t), dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
    args = [model.encode(), P(p_unscaled), P(k), plist.ctypes.data_as(C.POINTER(C.c_int)), ns,
            len(t), P(t), P(my), P(sd), t0, rtol, atol, maxsteps, stldet, errconS, panel_width,
            C.byref(llh), P(sllh), stats.ctypes.data_as(C.POINTER(C.c_int64)), P(y), P(s),
            trace.ctypes.data_as(C.POINTER(C.c_int)), int(sparse), deg_max, stage_doubles,
            int(fuse_history) + 2 * int(fuse_newton)]
    if steady or k_preeq is not None:
        kp = d(k_preeq) if k_preeq is not None else None
        ri = np.ascontiguousarray(list(reinit_idx) or [0], dtype=np.int32)
        st = lib_team().hs_team_simulate_ss(*args, P(kp), ri.ctypes.data_as(C.POINTER(C.c_int)), len(reinit_idx))
    else:
        st = lib_team().hs_team_simulate(*args)
    return dict(numsteps=trace[0::2], order=trace[1::2], status=st, llh=llh.value, sllh=sllh[:ns], stats=stats, y=y, sigmay=s)


def team_solve(model, Jcsr, gamma, B, panel_width=0, sparse=False, deg_max=0, stage_doubles=0, want_info=False):
    """(I - gamma*J)^-1 B through the blocked Gauss-Jordan inversion and the solve GEMM -- or, with
    ``sparse``, through the sparse LU of the low-degree pivots + the inverse of the dense Schur block;
    B is [ncols, nx]."""
    Jcsr = np.ascontiguousarray(Jcsr, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    X = np.zeros_like(B)
    info = np.zeros(8, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib_team().hs_team_solve(model.encode(), P(Jcsr), gamma, P(B), B.shape[0], panel_width, P(X), int(sparse),
                                  deg_max, stage_doubles, info.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise RuntimeError("hs_team_solve failed: %d" % rc)
    if want_info:
        return X, dict(zip(["n1", "m", "nnzL", "nnzU", "pairs", "factor_levels", "fwd_levels", "bwd_levels"], info.tolist()))
    return X
