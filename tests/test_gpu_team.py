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
