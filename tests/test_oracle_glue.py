"""The restated parPE glue (A5/A6/A8) has no test in the reference; pin it by construction
properties and by the reference's own gradient check method (central finite differences,
src/parpeoptimization/optimizationProblem.cpp:50-131)."""
import math

import numpy as np
import pytest

from parpe_b200 import synth
from parpe_b200.tables import SCALE_LN, SCALE_LOG10, SCALE_NONE


def test_mapping_and_plist(oracle_available):
    from oracle.binding import OracleProblem
    tab = synth.make_problem("steadystate_scaled", 4, scale="mixed", condition_specific=2)
    tab.mapping[1, 3] = -1  # unmapped parameter -> override value, not in plist
    tab.overrides = np.full_like(tab.mapping, np.nan, dtype=float)
    tab.overrides[1, 3] = 0.123
    tab.__post_init__()
    o = OracleProblem(tab)
    theta = synth.random_theta(tab, seed=0)
    for c in range(4):
        p = o.map_parameters(c, theta)
        pl = o.plist(c)
        assert list(pl) == [i for i in range(tab.np_) if tab.mapping[c, i] >= 0]
        for i in range(tab.np_):
            j = tab.mapping[c, i]
            if j < 0:
                assert p[i] == tab.overrides[c, i]
                continue
            lin = {SCALE_LOG10: 10.0 ** theta[j], SCALE_LN: math.exp(theta[j]), SCALE_NONE: theta[j]}[
                int(tab.pscale_opt[j])]
            exp = {SCALE_LOG10: math.log10(lin), SCALE_LN: math.log(lin), SCALE_NONE: lin}[
                int(tab.pscale_sim[c, i])]
            assert p[i] == pytest.approx(exp, rel=1e-14)


def test_apply_chain_rule(oracle_available):
    from oracle.binding import lib
    f = lib().ora_apply_chain_rule
    assert f(3.0, 0.5, SCALE_LOG10, SCALE_LOG10) == 3.0
    # d/dlog10(p) -> d/dp: divide by p ln 10 (p given in the OLD scale: log10 p = 0.5)
    assert f(3.0, 0.5, SCALE_LOG10, SCALE_NONE) == pytest.approx(3.0 / (10 ** 0.5 * math.log(10)))
    # d/dp -> d/dlog10(p); quirk Q4: the parameter is unscaled with the old (linear) scale
    assert f(3.0, 2.0, SCALE_NONE, SCALE_LOG10) == pytest.approx(3.0 * 2.0 * math.log(10))
    assert f(3.0, 0.5, SCALE_LN, SCALE_NONE) == pytest.approx(3.0 / math.exp(0.5))


@pytest.mark.parametrize("model,kw", [("steadystate_scaled", dict(scale="mixed", condition_specific=1)),
                                      ("boehm", dict(scale="log10"))])
def test_gradient_vs_finite_differences(oracle_available, model, kw):
    from oracle.binding import OracleProblem
    tab = synth.make_problem(model, 3, seed=1, sigma_mode="mixed", **kw)
    o = OracleProblem(tab)
    theta = synth.random_theta(tab, seed=2, width=0.2)
    r = o.evaluate(theta)
    assert r["status"] == 0
    g = r["grad"]
    scale = np.max(np.abs(g))
    for j in range(tab.n_opt):
        best = np.inf
        for eps in (1e-2, 1e-3, 1e-4, 1e-5):
            tp, tm = theta.copy(), theta.copy()
            tp[j] += eps
            tm[j] -= eps
            fd = (o.evaluate(tp, want_grad=False)["fval"] - o.evaluate(tm, want_grad=False)["fval"]) / (2 * eps)
            # regularised relative error of the reference: |g - fd| / |fd + eps|; entries that vanish on
            # the gradient's scale are compared absolutely
            err = abs(fd - g[j]) / max(abs(fd + eps), 1e-3 * scale)
            best = min(best, err)
        assert best < 1e-2, (j, best, g[j])  # the reference reports an ERROR above 1e-2 (WARN above 1e-3)


def test_retry_relaxes_tolerances(oracle_available):
    """runAndLogSimulation: trial t uses tolerances * relaxation^(t-1); a run that needs more than
    maxsteps at tight tolerances succeeds at a later trial."""
    from oracle.binding import OracleProblem
    tab = synth.make_problem("steadystate", 1, scale="lin", maxsteps=150, num_trials=6)
    r = OracleProblem(tab).evaluate(synth.nominal_theta(tab), collect_stats=True)
    assert r["status"] == 0
    assert r["stats"][0, 12] > 1  # trials used
    tab1 = synth.make_problem("steadystate", 1, scale="lin", maxsteps=150, num_trials=1)
    r1 = OracleProblem(tab1).evaluate(synth.nominal_theta(tab1))
    assert r1["status"] == 1 and np.isinf(r1["fval"])


def test_duplicate_and_subset_indices(oracle_available):
    from oracle.binding import OracleProblem
    tab = synth.make_problem("steadystate", 5, seed=4)
    o = OracleProblem(tab)
    theta = synth.nominal_theta(tab)
    full = o.evaluate(theta)
    sub = o.evaluate(theta, idx=[3, 3, 1])
    assert sub["fval"] == pytest.approx(2 * (-full["llh"][3]) - full["llh"][1], rel=1e-12)
