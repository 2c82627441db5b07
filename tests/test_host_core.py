"""The device integrator core (parpe_b200/csrc/bdf_core.cuh), compiled for the host with the
one-lane group, against the CVODES oracle on the same inputs: same tolerances as the GPU parity
tests, so the algorithm is parity-checked on every CPU run (the GPU tests then check the CUDA
build of the same source through the C ABI)."""
import numpy as np
import pytest

from parpe_b200 import synth
from parpe_b200.tables import SCALE_LN, SCALE_LOG10
from tests.util import assert_grad_close, assert_llh_close


def _unscaled_p(tab, c, theta, oracle):
    p_sim = oracle.map_parameters(c, theta)
    out = np.empty_like(p_sim)
    for i, v in enumerate(p_sim):
        s = tab.pscale_sim[c, i]
        out[i] = 10.0 ** v if s == SCALE_LOG10 else (np.exp(v) if s == SCALE_LN else v)
    return out


CASES = [("steadystate", dict(scale="lin")), ("steadystate_scaled", dict(scale="log10", sigma_mode="mixed")),
         ("boehm", dict(scale="log10")), ("fujita", dict(scale="log10"))]


@pytest.mark.parametrize("model,kw", CASES, ids=[c[0] for c in CASES])
def test_host_core_matches_cvodes(oracle_available, model, kw):
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate
    tab = synth.make_problem(model, 3, seed=7, missing_fraction=0.1, **kw)
    ora = OracleProblem(tab, 2, 1)  # staggered, dense: the configuration the kernel implements
    theta = synth.random_theta(tab, seed=3, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p = _unscaled_p(tab, c, theta, ora)
        plist = ora.plist(c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate(model, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), rtol=tab.rtol,
                     atol=tab.atol)
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        # AMICI reports sllh w.r.t. the scaled parameter: chain-rule factor applied by the caller
        fac = np.array([p[i] * np.log(10) if tab.pscale_sim[c, i] == SCALE_LOG10 else
                        (p[i] if tab.pscale_sim[c, i] == SCALE_LN else 1.0) for i in plist])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])
        # step counts stay close to CVODES' (identical heuristics, rounding-level divergence only)
        nst_ref = ro["stats"][c, 0]
        assert abs(r["stats"][0] - nst_ref) <= 0.1 * nst_ref + 5
        # cost only (no sensitivity columns)
        r0 = simulate(model, p, tab.k[c], [], tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), rtol=tab.rtol,
                      atol=tab.atol)
        assert r0["status"] == 0
        assert_llh_close(r0["llh"], ora.evaluate(theta, idx=[c], want_grad=False)["llh"][0])


def test_first_steps_identical_to_cvodes(oracle_available):
    """Same step/order sequence as CVODES until rounding-level differences accumulate."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate
    tab = synth.make_problem("steadystate", 1, scale="lin", seed=0, sigma_mode="data")
    theta = synth.nominal_theta(tab)
    ora = OracleProblem(tab, 2, 1)
    ns_ref, q_ref = ora.debug_numsteps(theta)
    r = simulate("steadystate", theta, tab.k[0], np.arange(5), tab.t, tab.y_of(0), tab.sigma_of(0))
    assert list(r["numsteps"][:12]) == list(ns_ref[:12])
    assert list(r["order"][:12]) == list(q_ref[:12])


def test_failure_codes(oracle_available):
    from tests.host_harness.binding import simulate
    tab = synth.make_problem("steadystate", 1, scale="lin")
    theta = synth.nominal_theta(tab)
    r = simulate("steadystate", theta, tab.k[0], np.arange(5), tab.t, tab.y_of(0), tab.sigma_of(0), maxsteps=20)
    assert r["status"] in (-1, -99) and np.isnan(r["llh"])  # too much work / max steps reached
    bad_sigma = np.full_like(tab.sigma_of(0), -1.0)
    r = simulate("steadystate", theta, tab.k[0], np.arange(5), tab.t, tab.y_of(0), bad_sigma)
    assert r["status"] == -99 and np.isnan(r["llh"])  # Model::checkSigmaPositivity -> AMICI_ERROR


@pytest.mark.parametrize("model,kw", [("steadystate", dict(scale="lin", sigma_mode="data")),
                                      ("boehm", dict(scale="log10")),
                                      ("fujita", dict(scale="log10", sigma_mode="mixed", missing_fraction=0.2))],
                         ids=["steadystate", "boehm", "fujita"])
def test_simultaneous_corrector_matches_cvodes(oracle_available, model, kw):
    """pb200_problem_desc::ism = 1 (CV_SIMULTANEOUS, AMICI's InternalSensitivityMethod::simultaneous; bit 2 of
    SolverOptions::errconS, full-feature instantiation of the integrator): state and sensitivities in ONE Newton
    iteration with the joint convergence norm and ONE error test (cvodes_nls_sim.c:374-560) -- against the reference
    run with ism = simultaneous: the same step / order sequence as CVODES until rounding-level differences accumulate
    (as for the staggered corrector), llh and gradient at the path's tolerances, CVODES' own counters (no separate
    sensitivity iterations or error-test failures)."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_ss
    from tests.util import assert_grad_close, assert_llh_close, condition_inputs
    tab = synth.make_problem(model, 3, seed=0, **kw)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    tab.ism = 1
    ora = OracleProblem(tab, 2, 1)  # (the descriptor's ism wins over the argument)
    ro = ora.evaluate(theta, collect_stats=True)
    stag = OracleProblem(synth.make_problem(model, 3, seed=0, **kw), 2, 1).evaluate(theta, collect_stats=True)
    assert ro["status"] == 0
    differs = False
    for c in range(tab.n_conditions):
        p, plist, fac = condition_inputs(tab, theta, c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate_ss(tab.model_name, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), errconS=5)
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])
        nst_ref = ro["stats"][c, 0]
        assert abs(r["numsteps"][-1] - nst_ref) <= 0.1 * nst_ref + 2  # (the scatter of the staggered corrector: rounding decides)
        differs = differs or ro["stats"][c, 0] != stag["stats"][c, 0]
    assert differs  # (the reference's two correctors do take different step sequences: the case is exercised)
    # the first steps are CVODES' exactly
    tab1 = synth.make_problem("steadystate", 1, scale="lin", seed=0, sigma_mode="data")
    tab1.ism = 1
    th1 = synth.nominal_theta(tab1)
    ns_ref, q_ref = OracleProblem(tab1, 2, 1).debug_numsteps(th1)
    r = simulate_ss("steadystate", th1, tab1.k[0], np.arange(5), tab1.t, tab1.y_of(0), tab1.sigma_of(0), errconS=5)
    assert list(r["numsteps"][:10]) == list(ns_ref[:10]) and list(r["order"][:10]) == list(q_ref[:10])
