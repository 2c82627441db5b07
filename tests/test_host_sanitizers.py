"""AddressSanitizer + UndefinedBehaviorSanitizer over the HOST build of both integrator cores.

compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer.log keeps its answer), so memcheck /
racecheck logs of the kernels cannot be produced.  What can be checked without it: the device cores compile for
the host from the SAME source (one-lane group / one-thread team), so every index computation of the algorithm --
history arrays, ELL Jacobian, blocked inversion, sparse factor tables and level lists, staging-area layouts of the
fused passes, column batches, the steady-state paths (pre-/post-equilibration, Newton's method and its linear
solves in the group's scratch arrays) -- runs here under ASan/UBSan with heap-allocated buffers of exactly the sizes the
kernels carve from HBM / shared memory; an out-of-bounds index in that logic aborts the run.  (Races between the
threads of a block are outside what a one-thread build can see.)"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HH = os.path.join(ROOT, "tests", "host_harness")


def _asan_lib():
    p = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


TEAM_BODY = """
    from tests.host_harness import binding as hh
    hh.LIB_TEAM = %r
    hh.build_team = lambda force=False: None
    import numpy as np
    from parpe_b200 import synth
    from parpe_b200.models import get_spec
    spec = get_spec("net32")
    tab = synth.make_problem("net32", 1, seed=7, timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:5])
    p = np.asarray(spec.p_nominal) * 1.3
    args = ("net32", p, tab.k[0], list(range(spec.np)), tab.t, tab.y_of(0), tab.sigma_of(0))
    for kw in (dict(), dict(sparse=True), dict(sparse=True, deg_max=3, stage_doubles=150),
               dict(sparse=True, deg_max=100), dict(sparse=True, fuse_history=False, fuse_newton=False),
               dict(panel_width=4),
               # the simultaneous corrector of the steady-state instantiation (SolverOptions::errconS bit 2)
               dict(errconS=5, steady=True), dict(errconS=5, steady=True, sparse=True),
               dict(errconS=5, steady=True, sparse=True, deg_max=3, stage_doubles=150)):
        r = hh.simulate_team(*args, **kw)
        assert r["status"] == 0, (kw, r["status"])
    rng = np.random.default_rng(0)
    sp = get_spec("chen")
    Jv = rng.standard_normal(sp.nnz)
    rows = np.repeat(np.arange(sp.nx), np.diff(sp.j_ptr))
    Jv[rows == sp.j_col] = -np.abs(Jv[rows == sp.j_col]) - 3.0
    hh.team_solve("chen", Jv, 0.3, rng.standard_normal((51, sp.nx)), sparse=True)
    hh.team_solve("chen", Jv, 0.3, rng.standard_normal((51, sp.nx)), sparse=True, stage_doubles=6000)
    hh.team_solve("net31", rng.standard_normal(get_spec("net31").nnz) * 0.1, 0.3, rng.standard_normal((6, 31)))
    # steady-state instantiation: pre-equilibration + re-initialisation + a t = inf output time, both linear solvers
    tp = np.array([0.1, 0.3, np.inf])
    tab = synth.make_problem("net32", 3, seed=7, timepoints=tp, preeq=True)
    for c in (0, 1, 2):
        pre = bool(tab.has_preeq[c])
        for kw in (dict(), dict(sparse=True)):
            r = hh.simulate_team("net32", p, tab.k[c], list(range(spec.np)), tp, tab.y_of(c), tab.sigma_of(c), steady=True,
                                 k_preeq=tab.k_preeq[c] if pre else None, reinit_idx=tab.reinit_idx[c] if pre else (), **kw)
            assert r["status"] == 0, (c, kw, r["status"])
    print("ASAN-OK")
"""

SMALL_BODY = """
    from tests.host_harness import binding as hh
    hh.LIB = %r
    hh.build = lambda force=False: None
    import numpy as np
    from parpe_b200 import synth
    from parpe_b200.models import get_spec
    for model in ("steadystate", "boehm", "fujita"):
        spec = get_spec(model)
        tab = synth.make_problem(model, 1, seed=3)
        p = np.asarray(spec.p_nominal, dtype=float)
        r = hh.simulate(model, p, tab.k[0], list(range(spec.np)), tab.t, tab.y_of(0), tab.sigma_of(0))
        assert r["status"] == 0, model
        r = hh.simulate(model, p, tab.k[0], [0, 2], tab.t, tab.y_of(0), tab.sigma_of(0))
        assert r["status"] == 0, model
        # steady-state instantiation: pre-equilibration + re-initialisation + t = inf output times, by simulation and
        # with Newton's method (first attempt, failed attempt + simulation, retry after a failed simulation, Newton only)
        tp = np.concatenate([np.asarray(tab.t[:4], dtype=float), [np.inf, np.inf]])
        ts = synth.make_problem(model, 3, seed=3, timepoints=tp, preeq=True)
        for c in (0, 1, 2):
            pre = bool(ts.has_preeq[c])
            for kw in (dict(), dict(newton_maxsteps=40), dict(newton_maxsteps=2), dict(newton_maxsteps=3, maxsteps=150),
                       dict(newton_maxsteps=-40), dict(newton_maxsteps=-2),  # (negative = Newton's method only)
                       dict(errconS=3), dict(errconS=3, newton_maxsteps=2), dict(errconS=5), dict(errconS=7)):  # (bit 1: steady-state
                # sensitivities by the linear solve; bit 2: simultaneous corrector)
                r = hh.simulate_ss(model, p, ts.k[c], list(range(spec.np)), tp, ts.y_of(c), ts.sigma_of(c),
                                   k_preeq=ts.k_preeq[c] if pre else None, reinit_idx=ts.reinit_idx[c] if pre else (), **kw)
                assert r["status"] in (0, -99, -1), (model, c, kw, r["status"])
    # the imported model with a spline input (csrc/amici_functions.cuh: coefficient arrays on the stack), if built
    import json, os
    from tests.util import jakstat_tables, condition_inputs
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(hh.__file__))))
    fx = os.path.join(root, "tests", "golden", "imported", "jakstat_sensiforward.json")
    if os.path.exists(os.path.join(root, "parpe_b200", "csrc", "models", "imported", "amici_jakstat_adjoint.cuh")):
        doc = json.load(open(fx))
        tj = jakstat_tables(doc)
        for c in range(tj.n_conditions):
            pj, plist, _ = condition_inputs(tj, np.array(doc["theta"]), c)
            lo, hi = tj.t_offsets[c], tj.t_offsets[c + 1]
            for ec in (1, 5):
                r = hh.simulate_ss("amici_jakstat_adjoint", pj, tj.k[c], plist, tj.t[lo:hi], tj.y_of(c), tj.sigma_of(c), errconS=ec)
                assert r["status"] == 0, (c, ec, r["status"])
    print("ASAN-OK")
"""


@pytest.mark.parametrize("which", ["team", "small"])
def test_host_cores_under_asan_ubsan(tmp_path, which):
    asan = _asan_lib()
    if asan is None:
        pytest.skip("libasan not available")
    src = os.path.join(HH, "host_sim_team.cpp" if which == "team" else "host_sim.cpp")
    lib = str(tmp_path / ("libhost_%s_asan.so" % which))
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address,undefined",
                           "-fno-sanitize-recover=undefined", "-o", lib, src])
    body = (TEAM_BODY if which == "team" else SMALL_BODY) % lib
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(body)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "ASAN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
