"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle (reference
AMICI/CVODES + restated parPE glue) on the same seeded inputs.

Tolerances are the north star's: llh 1e-6 relative, gradient 1e-4 relative at
rtol=1e-8 / atol=1e-12; integer results (status, plist-compaction) exact.
"""
import numpy as np
import pytest

from parpe_b200 import synth
from tests.util import assert_grad_close, assert_llh_close

pytestmark = pytest.mark.gpu


def _oracle_sim(model):
    from oracle.binding import OracleProblem

    def simulate(tab, theta):
        rc, y, _ = OracleProblem(tab).get_model_outputs(theta)
        assert rc == 0
        return y
    return simulate


def _both(tab):
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    return B200SummedGradientFunction(tab), OracleProblem(tab)


CASES = [
    ("steadystate", 4, dict(scale="lin")),
    ("steadystate", 8, dict(scale="log10", sigma_mode="data")),
    ("steadystate_scaled", 4, dict(scale="log10", condition_specific=2)),
    ("steadystate_scaled", 6, dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.2, ragged=True)),
    ("boehm", 3, dict(scale="log10")),
    ("boehm", 5, dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.1)),
    ("fujita", 6, dict(scale="log10")),
]


@pytest.mark.parametrize("model,C,kw", CASES, ids=["%s-%d-%s" % (m, c, k.get("scale")) for m, c, k in CASES])
def test_evaluate_matches_oracle(oracle_available, model, C, kw):
    tab = synth.make_problem(model, C, simulate=_oracle_sim(model), seed=3, **kw)
    gpu, ora = _both(tab)
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
        # cost only: gradient.empty() in the reference
        st, fval0, g0, _ = gpu.evaluate(theta, want_gradient=False)
        ro0 = ora.evaluate(theta, want_grad=False)
        assert st == 0 and g0 is None
        assert_llh_close(fval0, ro0["fval"], "nllh (cost only)")


def test_minibatch_subset_and_single_index(oracle_available):
    tab = synth.make_problem("steadystate_scaled", 9, simulate=_oracle_sim("steadystate_scaled"), seed=5,
                             condition_specific=1)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=2, width=0.2)
    for idx in ([5, 2, 7], [8], [3, 3, 0], list(range(9))[::-1]):
        ro = ora.evaluate(theta, idx=idx)
        st, fval, grad, _ = gpu.evaluate(theta, idx)
        assert st == 0
        assert_llh_close(fval, ro["fval"])
        assert_grad_close(grad, ro["grad"])
    # single-int overload (multiConditionProblem.cpp:637-645)
    st, fval, grad, _ = gpu.evaluate(theta, 4)
    ro = ora.evaluate(theta, idx=[4])
    assert_llh_close(fval, ro["fval"])


def test_empty_index_list():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("steadystate", 3)
    gpu = B200SummedGradientFunction(tab)
    st, fval, grad, _ = gpu.evaluate(synth.nominal_theta(tab), [])
    assert st == 0 and fval == 0.0 and np.all(grad == 0.0)


def test_model_outputs_match_oracle(oracle_available):
    tab = synth.make_problem("boehm", 4, simulate=_oracle_sim("boehm"), seed=1, ragged=True)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=4, width=0.2)
    rc_o, y_o, s_o = ora.get_model_outputs(theta)
    rc_g, y_g, s_g = gpu.getModelOutputsAndSigmas(theta)
    assert rc_o == rc_g == 0
    np.testing.assert_allclose(y_g, y_o, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(s_g, s_o, rtol=1e-12)


def test_failure_maps_to_inf(oracle_available):
    """A condition that cannot be integrated => fval=+inf, functionEvaluationFailure
    (multiConditionProblem.cpp:669-672), on both sides."""
    tab = synth.make_problem("steadystate", 3, scale="lin", maxsteps=20, num_trials=2)
    gpu, ora = _both(tab)
    theta = synth.nominal_theta(tab)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert ro["status"] == 1 and st == 1
    assert np.isinf(fval) and np.isinf(ro["fval"])
    assert np.all(gpu.condition_results()["status"] != 0)


def test_wrong_parameter_count_is_an_error():
    from parpe_b200 import capi
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("steadystate", 2)
    gpu = B200SummedGradientFunction(tab)
    with pytest.raises(capi.B200LibraryError):
        gpu.evaluate(np.zeros(tab.n_opt + 1))


def test_partition_sums_to_whole(oracle_available):
    """Static contiguous partition: partial sums of 2 'ranks' add up to the 1-rank result."""
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("steadystate_scaled", 7, simulate=_oracle_sim("steadystate_scaled"), seed=9)
    theta = synth.random_theta(tab, seed=3, width=0.2)
    whole = B200SummedGradientFunction(tab)
    _, f, g, _ = whole.evaluate(theta)
    fs, gs = 0.0, 0.0
    for r in range(2):
        part = B200SummedGradientFunction(tab)
        part.set_partition(r, 2)
        _, fr, gr, _ = part.evaluate(theta)
        fs, gs = fs + fr, gs + gr
    assert_llh_close(fs, f)
    assert_grad_close(gs, g)


def test_reproducible_run_to_run():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("boehm", 37, seed=2)
    gpu = B200SummedGradientFunction(tab)
    theta = synth.random_theta(tab, seed=1, width=0.5)
    a = gpu.evaluate(theta)
    b = gpu.evaluate(theta)
    assert a[1] == b[1] and np.array_equal(a[2], b[2])
