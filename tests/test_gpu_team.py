"""GPU parity tests of the large-model path (one thread block per condition, table-driven
mass-action networks): CUDA path through the C ABI vs the CPU oracle (reference AMICI/CVODES
with the same tables as an amici::Model_ODE) on the same seeded inputs.

Tolerances are the north star's: llh 1e-6 relative, gradient 1e-4 relative at rtol=1e-8 /
atol=1e-12; integer results (status) exact.
"""
import numpy as np
import pytest

from parpe_b200 import synth
from tests.util import assert_grad_close, assert_llh_close

pytestmark = pytest.mark.gpu


def _both(tab):
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    return B200SummedGradientFunction(tab), OracleProblem(tab)


CASES = [
    ("net32", 5, 8, dict(scale="log10")),
    ("net32", 7, 8, dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.1, ragged=True)),
    ("net32", 300, 4, dict(scale="log10")),       # more conditions than SMs: work queue
    ("net31", 6, 6, dict(scale="log10")),         # odd nx (8-byte tile copies), one short block (nb = 31)
    ("chen", 3, 3, dict(scale="log10")),          # nx = 500, nplist = 50: rows path (tensor-core GEMMs)
    ("froehlich", 2, 1, dict(scale="log10")),     # nx = 1200, nplist = 100: generic path (panel in shared memory)
]


@pytest.mark.parametrize("model,C,ntp,kw", CASES, ids=["%s-%d-%s" % (m, c, k.get("scale")) for m, c, _, k in CASES])
def test_team_evaluate_matches_oracle(oracle_available, model, C, ntp, kw):
    tab = synth.make_problem(model, C, seed=3, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp], **kw)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"], "nllh")
    assert_grad_close(grad, ro["grad"], "gradient")
    res = gpu.condition_results()
    # a wrong inverse would still converge (Newton corrects itself) but with more iterations and
    # convergence failures: the step counts must stay close to CVODES' and failures rare
    nst, nst_ref = res["stats"][:, 0], ro["stats"][:, 0]
    assert np.all(np.abs(nst - nst_ref) <= 0.1 * nst_ref + 5), (nst, nst_ref)
    assert np.all(res["stats"][:, 9] + res["stats"][:, 10] <= 0.02 * nst + 2), res["stats"][:, 9:11]
    assert np.array_equal(res["status"], ro["cond_status"])
    assert_llh_close(res["llh"], ro["llh"], "per-condition llh")
    for c in range(min(C, 8)):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    # cost only: gradient.empty() in the reference
    st, fval0, g0, _ = gpu.evaluate(theta, want_gradient=False)
    ro0 = ora.evaluate(theta, want_grad=False)
    assert st == 0 and g0 is None
    assert_llh_close(fval0, ro0["fval"], "nllh (cost only)")


def test_team_minibatch_subset_and_outputs(oracle_available):
    tab = synth.make_problem("net32", 9, seed=5, timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:5])
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=2, width=0.2)
    for idx in ([5, 2, 7], [8], [3, 3, 0]):
        ro = ora.evaluate(theta, idx=idx)
        st, fval, grad, _ = gpu.evaluate(theta, idx)
        assert st == 0
        assert_llh_close(fval, ro["fval"])
        assert_grad_close(grad, ro["grad"])
    rc, y, s = gpu.getModelOutputsAndSigmas(theta)
    rco, yo, so = ora.get_model_outputs(theta)
    assert rc == rco == 0
    np.testing.assert_allclose(y, yo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(s, so, rtol=0, atol=0)


def test_team_reproducible_run_to_run():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("net32", 40, seed=9, timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:3])
    gpu = B200SummedGradientFunction(tab)
    theta = synth.random_theta(tab, seed=4, width=0.2)
    a = gpu.evaluate(theta)
    b = gpu.evaluate(theta)
    assert a[1] == b[1] and np.array_equal(a[2], b[2])


def test_team_failure_maps_to_inf(oracle_available):
    """A condition that cannot be integrated => fval=+inf, functionEvaluationFailure
    (multiConditionProblem.cpp:669-672), on both sides."""
    tab = synth.make_problem("net32", 3, seed=1, maxsteps=20, num_trials=2,
                             timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:2])
    gpu, ora = _both(tab)
    theta = synth.nominal_theta(tab)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert ro["status"] == 1 and st == 1
    assert np.isinf(fval) and np.isinf(ro["fval"])
    res = gpu.condition_results()
    assert np.all(res["status"] != 0)
    assert np.all(res["stats"][:, 12] == 2)  # both trials used (A8: runAndLogSimulation)


def test_team_tolerance_relaxation_retry(oracle_available):
    """The second trial runs at tolerances x relaxation factor (multiConditionProblem.cpp:369-465):
    with maxsteps too small for rtol 1e-8 but enough for the relaxed tolerances the evaluation
    succeeds on trial 2, as in the reference."""
    tab = synth.make_problem("net32", 2, seed=2, maxsteps=200, num_trials=3, relaxation_factor=1e4,
                             timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:2])
    gpu, ora = _both(tab)
    theta = synth.nominal_theta(tab)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert st == ro["status"]
    assert np.array_equal(res["stats"][:, 12], ro["stats"][:, 12])  # same number of trials
    if st == 0:
        assert np.all(res["stats"][:, 12] >= 2)
        assert abs(fval - ro["fval"]) <= 1e-3 * abs(ro["fval"])  # relaxed tolerances: looser agreement


def test_team_partition_sums_to_whole():
    """Static contiguous partition: partial sums of 2 'ranks' add up to the 1-rank result."""
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("net32", 7, seed=9, timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:3])
    theta = synth.random_theta(tab, seed=3, width=0.2)
    whole = B200SummedGradientFunction(tab)
    _, f, g, _ = whole.evaluate(theta)
    fs, gs = 0.0, 0.0
    for r in range(2):
        part = B200SummedGradientFunction(tab)
        part.set_partition(r, 2, partial_sums=True)
        _, fr, gr, _ = part.evaluate(theta)
        fs, gs = fs + fr, gs + gr
    assert abs(fs - f) <= 1e-12 * abs(f)
    np.testing.assert_allclose(gs, g, rtol=1e-11, atol=1e-12 * np.max(np.abs(g)))


@pytest.mark.parametrize("model,C,ntp", [("boehm", 200, 16), ("net32", 200, 2)])
def test_processing_order_does_not_change_results(model, C, ntp):
    """The cost-aware processing order (second evaluation of the same index list) changes which
    conditions run side by side, never a result: per-condition outputs and the sums are bit-identical."""
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem(model, C, seed=4, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp])
    theta = synth.random_theta(tab, seed=6, width=0.3)
    gpu = B200SummedGradientFunction(tab)
    a = gpu.evaluate(theta)          # natural order
    ra = gpu.condition_results()
    b = gpu.evaluate(theta)          # ordered by the first evaluation's step counts
    rb = gpu.condition_results()
    assert a[1] == b[1] and np.array_equal(a[2], b[2])
    assert np.array_equal(ra["llh"], rb["llh"]) and np.array_equal(ra["sllh"], rb["sllh"])
    assert np.array_equal(ra["stats"][:, :14], rb["stats"][:, :14])
