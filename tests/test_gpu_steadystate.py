"""Steady states on the GPU: post-equilibration (output times t = inf) and pre-equilibration
with re-initialisation of states through the C ABI (pb200_set_preequilibration) against the reference's own
SteadystateProblem (deps/AMICI/src/steadystateproblem.cpp, oracle/_ref) on the same seeded inputs.  The committed
fixtures steadystate_posteq / steadystate_scaled_preeq / boehm_preeq_posteq are covered by
tests/test_golden_fixtures.py::test_cuda_path_against_fixture."""
import ctypes as C

import numpy as np
import pytest

from parpe_b200 import capi, synth
from tests.util import assert_grad_close, assert_llh_close, assert_outputs_close

pytestmark = pytest.mark.gpu
SIM_ERROR = -99


def _oracle_sim(tab, theta):
    from oracle.binding import OracleProblem
    rc, y, _ = OracleProblem(tab).get_model_outputs(theta)
    assert rc == 0
    return y


def _problem(model, C, preeq, posteq, n_finite=6, seed=3, **kw):
    tp = np.asarray(synth.DEFAULT_TIMEPOINTS[model], dtype=float)[:n_finite + posteq]
    return synth.make_problem(model, C, seed=seed, timepoints=tp, preeq=preeq, posteq=posteq,
                              simulate=_oracle_sim, **kw)


def _both(tab):
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    return B200SummedGradientFunction(tab), OracleProblem(tab)


CASES = [
    ("steadystate", 4, False, 2, dict(scale="lin")),
    ("steadystate", 8, True, 0, dict(scale="log10", sigma_mode="data")),
    ("steadystate_scaled", 6, True, 2, dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.2, ragged=True)),
    ("boehm", 5, True, 1, dict(scale="log10")),
    ("boehm", 4, False, 2, dict(scale="mixed", sigma_mode="mixed")),
]


@pytest.mark.parametrize("model,C,preeq,posteq,kw", CASES,
                         ids=["%s-%d-%s%s" % (m, c, "preeq" if p else "", "+posteq" if q else "") for m, c, p, q, _ in CASES])
def test_steady_state_evaluate_matches_oracle(oracle_available, model, C, preeq, posteq, kw):
    tab = _problem(model, C, preeq, posteq, **kw)
    gpu, ora = _both(tab)
    assert gpu.last_launch() is not None
    for seed, width in ((0, 0.0), (1, 0.3)):
        theta = synth.random_theta(tab, seed=seed, width=width)
        ro = ora.evaluate(theta)
        st, fval, grad, _ = gpu.evaluate(theta)
        assert st == ro["status"] == 0
        assert_llh_close(fval, ro["fval"], "nllh")
        assert_grad_close(grad, ro["grad"], "gradient")
        res = gpu.condition_results()
        assert np.array_equal(res["status"], ro["cond_status"])
        assert_llh_close(res["llh"], ro["llh"], "per-condition llh")
        for c in range(C):
            assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
        # cost only: the pre-equilibration then integrates no sensitivities either
        st, fval0, g0, _ = gpu.evaluate(theta, want_gradient=False)
        ro0 = ora.evaluate(theta, want_grad=False)
        assert st == 0 and g0 is None
        assert_llh_close(fval0, ro0["fval"], "nllh (cost only)")
    # a minibatch with duplicates and the model outputs (t = inf rows: the observables at the steady state)
    sub = [C - 1, 0, 0, 1]
    st, fs, gs, _ = gpu.evaluate(theta, sub)
    rs = ora.evaluate(theta, idx=sub)
    assert_llh_close(fs, rs["fval"])
    assert_grad_close(gs, rs["grad"])
    rc, y, s = gpu.getModelOutputsAndSigmas(theta)
    rco, yo, so = ora.get_model_outputs(theta)
    assert rc == rco == 0
    assert_outputs_close(y, yo, tab)
    np.testing.assert_allclose(s, so, rtol=1e-12)


def test_fujita_shape_within_the_reference_variants_spread(oracle_available):
    """See tests/test_host_steadystate.py: where the steady-state stopping rule fires depends on the step sequence;
    on the Fujita shape the reference's KLU and dense variants differ by ~2e-5 at the t = inf points."""
    from oracle.binding import LINSOL_DENSE, LINSOL_KLU, OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    tab = _problem("fujita", 6, True, 2)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    a = OracleProblem(tab, linsol=LINSOL_KLU).evaluate(theta)
    b = OracleProblem(tab, linsol=LINSOL_DENSE).evaluate(theta)
    spread = abs(a["fval"] - b["fval"]) / abs(a["fval"])
    gs = np.max(np.abs(a["grad"]))
    gspread = np.max(np.abs(a["grad"] - b["grad"]) / np.maximum(np.abs(a["grad"]), 1e-4 * gs))
    st, fval, grad, _ = B200SummedGradientFunction(tab).evaluate(theta)
    assert st == a["status"] == 0
    assert abs(fval - a["fval"]) / abs(a["fval"]) <= max(3.0 * spread, 1e-6)
    assert np.max(np.abs(grad - a["grad"]) / np.maximum(np.abs(a["grad"]), 1e-4 * gs)) <= max(3.0 * gspread, 1e-4)


def test_many_conditions_through_the_work_queue(oracle_available):
    """More conditions than lane groups of a CTA wave would need is not required: 600 conditions, three of four
    pre-equilibrated, pulled from the in-kernel queue by the steady-state kernel; all results against the oracle."""
    tab = _problem("steadystate", 600, True, 1, scale="log10")
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=2, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    res = gpu.condition_results()
    assert_llh_close(res["llh"], ro["llh"])
    # run-to-run reproducible (fixed-order reduction; per-condition results independent of the schedule)
    st2, fval2, grad2, _ = gpu.evaluate(theta)
    assert fval2 == fval and np.array_equal(grad, grad2)


def test_equilibration_failure_maps_to_inf(oracle_available):
    """Too few steps to equilibrate: AMICI_ERROR per condition (steadystateproblem.cpp:397-425), retried with relaxed
    tolerances, then fval = +inf / functionEvaluationFailure (multiConditionProblem.cpp:669-672)."""
    tab = _problem("steadystate", 4, True, 1)
    tab.maxsteps = 150
    tab.num_trials = 2
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, _, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 1 and np.isinf(fval) and np.isinf(ro["fval"])
    res = gpu.condition_results()
    assert np.array_equal(res["status"], ro["cond_status"])
    assert np.all(res["status"] == SIM_ERROR)
    assert np.all(res["stats"][:, 12] == 2)  # both trials used


def test_relaxed_tolerance_retry_repeats_the_pre_equilibration(oracle_available):
    """maxsteps between what the first and the relaxed trial need: trial 1 fails in the equilibration, trial 2
    (tolerances x 100, pre-equilibration included) succeeds -- same trial count and result as the reference."""
    tab = _problem("steadystate", 4, True, 0)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    tab2 = _problem("steadystate", 4, True, 0)
    tab2.maxsteps = 450
    gpu2, ora2 = _both(tab2)
    ro = ora2.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu2.evaluate(theta)
    res = gpu2.condition_results()
    assert np.array_equal(res["status"], ro["cond_status"])
    assert np.array_equal(res["stats"][:, 12], ro["stats"][:, 12]), (res["stats"][:, 12], ro["stats"][:, 12])
    assert np.any(res["stats"][:, 12] > 1)
    if ro["status"] == 0:
        assert st == 0
        assert abs(fval - ro["fval"]) <= 1e-3 * abs(ro["fval"])  # (relaxed tolerances: 1e-6 / 1e-10)


def test_switching_pre_equilibration_off_restores_the_plain_problem():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("boehm", 6, seed=4, preeq=True)
    plain = synth.make_problem("boehm", 6, seed=4)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    f = B200SummedGradientFunction(tab)
    st, with_pre, g_pre, _ = f.evaluate(theta)
    assert st == 0
    capi.check(capi.lib().pb200_set_preequilibration(f._h, None, None, None, None), "off")
    st, off, g_off, _ = f.evaluate(theta)
    st, ref, g_ref, _ = B200SummedGradientFunction(plain).evaluate(theta)
    assert off == ref and np.array_equal(g_off, g_ref)
    assert with_pre != ref


TIGHT = dict(rtol=1e-11, atol=1e-15, maxsteps=100000)


@pytest.mark.parametrize("linsol", [capi.LINSOL_DENSE, capi.LINSOL_SPARSE], ids=["dense", "sparse"])
@pytest.mark.parametrize("model", ["net32", "net31"])
def test_large_model_kernel_steady_states(oracle_available, model, linsol):
    """The large-model (team) kernel's steady-state instantiation: pre-equilibration with re-initialisation and
    t = inf output times.  The generated networks equilibrate slowly (rate constants over six decades), so the
    reference's own KLU / dense variants only agree to ~1e-6 / 2e-4 (llh / sllh) even at rtol 1e-11 / atol 1e-15
    (tests/test_host_steadystate.py): the CUDA path has to stay within a small multiple of that spread."""
    from oracle.binding import LINSOL_DENSE, LINSOL_KLU, OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    tp = np.asarray(synth.DEFAULT_TIMEPOINTS[model], dtype=float)[:4]
    tab = synth.make_problem(model, 6, seed=0, timepoints=tp, preeq=True, posteq=1, simulate=_oracle_sim, **TIGHT)
    theta = synth.random_theta(tab, seed=1, width=0.2)
    a = OracleProblem(tab, linsol=LINSOL_KLU).evaluate(theta, collect_stats=True)
    b = OracleProblem(tab, linsol=LINSOL_DENSE).evaluate(theta)
    assert a["status"] == b["status"] == 0
    sl = np.max(np.abs(a["llh"] - b["llh"]) / np.abs(a["llh"]))
    gs = np.max(np.abs(a["sllh"]), axis=1)[:, None]
    sg = np.max(np.abs(a["sllh"] - b["sllh"]) / np.maximum(np.abs(a["sllh"]), 1e-4 * gs))
    assert sl < 1e-5 and sg < 2e-3, (sl, sg)
    gpu = B200SummedGradientFunction(tab)
    gpu.set_option(capi.OPT_LINSOL, linsol)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == 0
    res = gpu.condition_results()
    assert np.array_equal(res["status"], a["cond_status"])
    assert np.max(np.abs(res["llh"] - a["llh"]) / np.abs(a["llh"])) <= max(5.0 * sl, 3e-6)
    e = np.abs(res["sllh"] - a["sllh"]) / np.maximum(np.abs(a["sllh"]), 1e-4 * gs)
    assert np.max(e) <= max(5.0 * sg, 1e-3), (np.max(e), sg)
    assert abs(fval - a["fval"]) / abs(a["fval"]) <= max(5.0 * sl, 3e-6)
    # cost only: without sensitivities the stopping rule tests xdot alone and fires earlier -- a weaker criterion
    # with a spread of its own (net32: 4e-4 per condition between the reference's KLU and dense variants; net31: 1e-8)
    st, f0, _, _ = gpu.evaluate(theta, want_gradient=False)
    a0 = OracleProblem(tab, linsol=LINSOL_KLU).evaluate(theta, want_grad=False)
    b0 = OracleProblem(tab, linsol=LINSOL_DENSE).evaluate(theta, want_grad=False)
    sl0 = np.max(np.abs(a0["llh"] - b0["llh"]) / np.abs(a0["llh"]))
    res0 = gpu.condition_results()
    assert st == 0 and np.all(res0["status"] == 0)
    assert np.max(np.abs(res0["llh"] - a0["llh"]) / np.abs(a0["llh"])) <= max(5.0 * sl0, 3e-6)
    assert abs(f0 - a0["fval"]) / abs(a0["fval"]) <= max(5.0 * sl0, 3e-6)


def test_large_model_kernel_equilibration_failure(oracle_available):
    tab = synth.make_problem("net32", 4, seed=0, timepoints=np.array([0.1, 0.3, np.inf]), preeq=True)
    tab.maxsteps = 500
    tab.num_trials = 2
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.2)
    ro = ora.evaluate(theta)
    st, fval, _, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 1 and np.isinf(fval)
    res = gpu.condition_results()
    assert np.array_equal(res["status"], ro["cond_status"]) and np.all(res["status"] == SIM_ERROR)


def test_bad_reinitialisation_index_is_an_error():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("boehm", 2, seed=0, preeq=True)
    tab.reinit_idx = [[0], [8]]  # nx = 8
    with pytest.raises(capi.B200LibraryError, match="index out of range"):
        B200SummedGradientFunction(tab)


@pytest.mark.parametrize("newton_maxsteps,maxsteps", [(40, 10000), (2, 10000), (3, 150)], ids=["newton", "newton-fails-then-simulation", "retry"])
def test_newtons_method_for_steady_states(oracle_available, newton_maxsteps, maxsteps):
    """Newton's method first (the model's steady-state modes at integrateIfNewtonFails, newton_maxsteps > 0), the
    simulation from its last trial point if it fails, Newton once more if the simulation fails:
    deps/AMICI/src/steadystateproblem.cpp:120-166,628-715; newton_solver.cpp:43-125 (tests/test_host_steadystate.py has
    the details per path)."""
    tab = _problem("steadystate", 8, True, 1 if maxsteps < 1000 else 2)
    tab.newton_maxsteps = newton_maxsteps
    tab.maxsteps = maxsteps
    tab.num_trials = 1
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert np.array_equal(res["status"], ro["cond_status"])
    ok = ro["cond_status"] == 0
    assert np.any(ok)
    assert_llh_close(res["llh"][ok], ro["llh"][ok])
    for c in np.nonzero(ok)[0]:
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    assert st == ro["status"]
    if st == 0:
        assert_llh_close(fval, ro["fval"])
        assert_grad_close(grad, ro["grad"])
    st0, f0, _, _ = gpu.evaluate(theta, want_gradient=False)
    ro0 = ora.evaluate(theta, want_grad=False)
    assert st0 == ro0["status"]
    if st0 == 0:
        assert_llh_close(f0, ro0["fval"])


def test_large_model_kernel_refuses_newton():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("net32", 2, seed=0, timepoints=np.array([0.1, np.inf]))
    tab.newton_maxsteps = 10
    with pytest.raises(capi.B200LibraryError, match="Newton"):
        B200SummedGradientFunction(tab)
