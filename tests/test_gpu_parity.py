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
        part.set_partition(r, 2, partial_sums=True)
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


def _hier_outer(tab):
    full = synth.random_theta(tab, seed=6, width=0.2)
    return np.delete(full, tab.hier.inner_idx)


@pytest.mark.parametrize("maker,C", [(synth.steadystate_hierarchical, 8), (synth.fujita_hierarchical, 12)],
                         ids=["steadystate_scaled", "fujita"])
def test_hierarchical_matches_oracle(oracle_available, maker, C):
    """HierarchicalOptimizationWrapper::evaluate: pass 1 + device reductions + pass 2, vs the oracle's
    restatement of hierarchicalOptimization.cpp (which is pinned by the reference's known answers)."""
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    tab = maker(C, seed=2)
    # measurements from noise-free outputs at a nearby parameter vector so that the inner problem is well posed
    rc, ysim, _ = OracleProblem(tab).get_model_outputs(synth.nominal_theta(tab))
    rng = np.random.default_rng(0)
    tab.y = np.abs(ysim * (1.0 + 0.05 * rng.standard_normal(ysim.size))) + 1e-3
    tab.y[rng.random(tab.y.size) < 0.1] = np.nan
    gpu, ora = B200SummedGradientFunction(tab), OracleProblem(tab)
    outer = _hier_outer(tab)
    assert gpu.hier_num_outer() == ora.hier_num_outer() == len(outer)
    for want_grad in (False, True):
        ro = ora.hier_evaluate(outer, want_grad=want_grad)
        rg = gpu.hier_evaluate(outer, want_gradient=want_grad)
        assert ro["status"] == rg["status"] == 0
        inner = tab.hier.inner_idx
        np.testing.assert_allclose(rg["full_parameters"][inner], ro["full_parameters"][inner], rtol=1e-6)
        outer_idx = np.setdiff1d(np.arange(tab.n_opt), inner)
        assert np.array_equal(rg["full_parameters"][outer_idx], outer)  # splice: bit-exact
        assert_llh_close(rg["fval"], ro["fval"], "hierarchical nllh")
        if want_grad:
            # The outer gradient is the full gradient at the spliced vector restricted to the outer
            # parameters (hierarchicalOptimization.cpp:204-238).  It is very sensitive to the inner
            # parameters (a 1e-7 relative change of a scaling moves small entries by more than the
            # 1e-4 bound), so parity of the simulation pass is checked on IDENTICAL inputs: the
            # oracle evaluated at the spliced vector the CUDA path produced ...
            same_input = ora.evaluate(rg["full_parameters"])["grad"][outer_idx]
            assert_grad_close(rg["grad"], same_input, "hierarchical gradient (same spliced vector)")
            # ... and end to end against the oracle's own hierarchical gradient on the gradient's scale
            scale = np.max(np.abs(ro["grad"]))
            assert np.max(np.abs(rg["grad"] - ro["grad"])) <= 1e-5 * scale


def test_hierarchical_size_mismatch_is_an_error():
    from parpe_b200 import capi
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.steadystate_hierarchical(4)
    gpu = B200SummedGradientFunction(tab)
    with pytest.raises(capi.B200LibraryError):  # reference throws ParPEException (hierarchicalOptimization.cpp:155-160)
        gpu.hier_evaluate(np.zeros(tab.n_opt))


def test_full_size_properties():
    """At BASELINE's full sizes the oracle is too slow for the test suite; check size-independent
    properties instead: additivity over a split of the index list, replicated conditions contribute
    equally, run-to-run reproducibility."""
    from parpe_b200.problem import B200SummedGradientFunction
    C = 10000
    tab = synth.make_problem("boehm", C, seed=0)
    # replicate: condition c and c+5000 are identical
    tab.k[5000:] = tab.k[:5000]
    tab.y.reshape(C, -1)[5000:] = tab.y.reshape(C, -1)[:5000]
    gpu = B200SummedGradientFunction(tab)
    theta = synth.random_theta(tab, seed=1, width=0.5)
    st, f, g, _ = gpu.evaluate(theta)
    assert st == 0
    res = gpu.condition_results()
    assert np.array_equal(res["llh"][:5000], res["llh"][5000:])       # bit-exact replicas
    assert np.array_equal(res["sllh"][:5000], res["sllh"][5000:])
    st, f1, g1, _ = gpu.evaluate(theta, np.arange(0, 5000))
    st, f2, g2, _ = gpu.evaluate(theta, np.arange(5000, C))
    assert_llh_close(f1 + f2, f)
    assert_grad_close(g1 + g2, g)
    assert f1 == f2 and np.array_equal(g1, g2)


@pytest.mark.parametrize("model,lanes", [("boehm", 4), ("boehm", 8), ("boehm", 16), ("fujita", 8), ("fujita", 16),
                                         ("steadystate", 2), ("steadystate", 4), ("steadystate", 8)])
def test_every_lane_group_width_matches_oracle(oracle_available, model, lanes):
    """All instantiated lane-group widths (G/2, G, 2G) run the same algorithm: each must meet the
    parity tolerances on its own, whatever the default is."""
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem(model, 5, seed=21, lanes_per_condition=lanes)
    theta = synth.random_theta(tab, seed=8, width=0.3)
    gpu = B200SummedGradientFunction(tab)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert gpu.last_launch()["lanes"] == lanes
    ro = OracleProblem(tab).evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])


@pytest.mark.parametrize("maker,C", [(synth.steadystate_hierarchical, 10), (synth.fujita_hierarchical, 12)],
                         ids=["steadystate", "fujita"])
def test_hierarchical_partitioned_over_two_ranks(maker, C):
    """Hierarchical evaluation with the conditions split over two ranks (two handles with
    pb200_set_partition, the collective hook implemented with a thread barrier): the analytic inner
    parameters need sums over ALL conditions, so every rank must arrive at the single-rank result."""
    import threading
    from parpe_b200 import capi
    from parpe_b200.problem import B200SummedGradientFunction
    tab = maker(C, seed=2)
    full = synth.random_theta(tab, seed=5, width=0.2)
    outer = np.delete(full, tab.hier.inner_idx)
    ref = B200SummedGradientFunction(tab).hier_evaluate(outer, want_gradient=True)
    ref0 = B200SummedGradientFunction(tab).hier_evaluate(outer, want_gradient=False)
    assert ref["status"] == 0

    world = 2
    barrier = threading.Barrier(world)
    slots = [None] * world

    def make_allreduce(rank):
        def allreduce(arr, op):
            slots[rank] = arr.copy()
            barrier.wait()
            stacked = np.stack(slots)
            out = stacked.sum(axis=0) if op == capi.REDUCE_SUM else stacked.max(axis=0)
            barrier.wait()
            arr[:] = out
        return allreduce

    results = [None] * world
    results0 = [None] * world

    def run(rank):
        f = B200SummedGradientFunction(tab)
        f.set_partition(rank, world)
        f.set_allreduce(make_allreduce(rank))
        results[rank] = f.hier_evaluate(outer, want_gradient=True)
        results0[rank] = f.hier_evaluate(outer, want_gradient=False)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for r in range(world):
        assert results[r] is not None and results[r]["status"] == 0
        assert abs(results[r]["fval"] - ref["fval"]) <= 1e-10 * abs(ref["fval"])
        np.testing.assert_allclose(results[r]["grad"], ref["grad"], rtol=1e-9, atol=1e-10 * np.max(np.abs(ref["grad"])))
        np.testing.assert_allclose(results[r]["full_parameters"], ref["full_parameters"], rtol=1e-10)
        assert abs(results0[r]["fval"] - ref0["fval"]) <= 1e-10 * abs(ref0["fval"])
    # a partitioned hierarchical evaluation without the hook is refused
    f = B200SummedGradientFunction(tab)
    f.set_partition(0, 2)
    with pytest.raises(capi.B200LibraryError):
        f.hier_evaluate(outer)
