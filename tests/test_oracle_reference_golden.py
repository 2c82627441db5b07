"""Pin the oracle against the reference's own golden vectors / known-answer tests for this path
 before it is trusted as the checker.

  * examples/parpeamici/steadystate/exampleSteadystateScaledTest.h:40-48,64-79,104-108:
    steady state x = {0.456644592142075, 0.437977375496898, 0.033333333333333} (tol 1e-5) at
    p = {1, 0.5, 0.4, 2, 0.1}, k = {0.1, 0.4, 0.7, 1}; llh of matching data = -n/2 log(2 pi);
    MultiConditionProblem cost = +n/2 log(2 pi)
  * tests/parpeamici/hierarchicalOptimizationTest.cpp:428-445: parameter scaling round trips
  * the reference's pre-generated model deps/AMICI/models/model_steadystate compiled in place must
    agree with the model emitted by our generator
"""
import math

import numpy as np
import pytest

from parpe_b200.tables import SCALE_LN, SCALE_LOG10, SCALE_NONE, ProblemTables

P = [1.0, 0.5, 0.4, 2.0, 0.1]
K = [0.1, 0.4, 0.7, 1.0]
X_SS = [0.456644592142075, 0.437977375496898, 0.033333333333333]


def _tab(t, y, sigma, model="steadystate"):
    t = np.atleast_1d(np.asarray(t, dtype=float))
    return ProblemTables(model, 5, 4, 3, 5, [0, len(t)], t, y, sigma, [K], np.arange(5)[None],
                         np.zeros((1, 5), int), np.zeros(5, int))


@pytest.mark.parametrize("model", ["steadystate", "steadystate_vendored"])
def test_steady_state_golden(oracle_available, model):
    from oracle.binding import OracleProblem
    tab = _tab([1e8], X_SS, [1.0, 1.0, 1.0])
    o = OracleProblem(tab, model_name=model)
    rc, y, _ = o.get_model_outputs(P)
    assert rc == 0
    np.testing.assert_allclose(y, X_SS, atol=1e-5, rtol=0)
    # ... value reproduced in the survey container with the reference libraries
    np.testing.assert_allclose(y, [0.456644592140846, 0.437977375495366, 0.0333333333333333], rtol=1e-9)


def test_llh_of_matching_data_is_offset(oracle_available):
    """exampleSteadystateScaledTest.h:64-79,104-108: data == simulation, sigma 1 ->
    llh = -n/2 log(2 pi); the summed-gradient-function cost is its negative."""
    from oracle.binding import OracleProblem
    t = np.logspace(-5, 1, 20)
    tab0 = _tab(t, np.ones(60), np.ones(60))
    rc, ysim, _ = OracleProblem(tab0).get_model_outputs(P)
    tab = _tab(t, ysim, np.ones(60))
    r = OracleProblem(tab).evaluate(P)
    n = 60
    assert r["status"] == 0
    assert r["llh"][0] == pytest.approx(-0.5 * n * math.log(2 * math.pi), rel=1e-12)
    assert r["fval"] == pytest.approx(0.5 * n * math.log(2 * math.pi), rel=1e-12)
    np.testing.assert_allclose(r["grad"], 0.0, atol=1e-5)


def test_survey_probe_values(oracle_available):
    """SURVEY.md Appendix B probe (reference AMICI/CVODES built in place): llh and sllh at t=1e8."""
    from oracle.binding import OracleProblem
    tab = _tab([1e8], [0.45, 0.44, 0.03], [1.0, 1.0, 1.0])
    for ism in (1, 2):
        for linsol in (1, 9):
            r = OracleProblem(tab, ism, linsol).evaluate(P)
            assert r["llh"][0] == pytest.approx(-2.75684527597688, rel=1e-12)
            np.testing.assert_allclose(r["sllh"][0], [0.00138023894856, 0.000547478341967, -0.00289900735784,
                                                      -9.12463903278e-05, -0.00605486287517], rtol=1e-7)


def test_generated_model_matches_vendored_model(oracle_available):
    """Our generator's AMICI model == the reference's pre-generated model_steadystate."""
    from oracle.binding import OracleProblem
    rng = np.random.default_rng(0)
    t = np.logspace(-5, 1, 20)
    tab = _tab(t, 0.4 + 0.1 * rng.random(60), np.full(60, 0.1))
    a = OracleProblem(tab, 2, 1, model_name="steadystate").evaluate(P)
    b = OracleProblem(tab, 2, 1, model_name="steadystate_vendored").evaluate(P)
    assert a["fval"] == pytest.approx(b["fval"], rel=1e-9)
    np.testing.assert_allclose(a["grad"], b["grad"], rtol=1e-7)


def test_parameter_scaling_known_answers(oracle_available):
    """hierarchicalOptimizationTest.cpp:428-445"""
    from oracle.binding import lib
    L = lib()
    assert L.ora_get_unscaled(42.0, SCALE_NONE) == 42.0
    assert L.ora_get_scaled(42.0, SCALE_NONE) == 42.0
    assert L.ora_get_scaled(100.0, SCALE_LOG10) == 2.0
    assert L.ora_get_unscaled(2.0, SCALE_LOG10) == 100.0
    assert L.ora_get_scaled(math.exp(1.0), SCALE_LN) == pytest.approx(1.0, rel=1e-15)
    assert L.ora_get_unscaled(1.0, SCALE_LN) == pytest.approx(math.exp(1.0), rel=1e-15)
