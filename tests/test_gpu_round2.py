"""GPU tests added in round 2: the sparse fixed-pattern linear solver of the large-model path, the
A5 holes (mapping = -1 + parameterOverrides, nplist < np) on both kernels, rejected t = inf, the
global result of a partitioned evaluate (host hook and in-library NCCL), launch geometry for few
conditions, several handles / models / devices in one process.

Tolerances are the north star's: llh 1e-6 relative, gradient 1e-4 relative (entries below
1e-4 max|g| on the gradient's scale) at rtol=1e-8 / atol=1e-12; integer results exact.
"""
import threading

import numpy as np
import pytest

from parpe_b200 import capi, synth
from tests.util import assert_grad_close, assert_llh_close

pytestmark = pytest.mark.gpu


def _gpu(tab, **kw):
    from parpe_b200.problem import B200SummedGradientFunction
    return B200SummedGradientFunction(tab, **kw)


def _ora(tab):
    from oracle.binding import OracleProblem
    return OracleProblem(tab)


# ---------------------------------------------------------------- sparse LU + dense Schur block
@pytest.mark.parametrize("model,C,ntp,linsol", [("net32", 6, 8, capi.LINSOL_SPARSE), ("net31", 5, 6, capi.LINSOL_SPARSE),
                                                ("chen", 3, 4, capi.LINSOL_SPARSE), ("chen", 2, 3, capi.LINSOL_DENSE),
                                                ("froehlich", 2, 1, capi.LINSOL_SPARSE)])
def test_team_linear_solvers_match_oracle(oracle_available, model, C, ntp, linsol):
    """Both linear solvers of the large-model path -- sparse LU of the low-degree pivots + dense
    inverse of the Schur block (the reference's KLU role), and the dense inverse of the whole matrix
    (its `dense` solver) -- against the reference (KLU)."""
    tab = synth.make_problem(model, C, seed=3, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp],
                             scale="mixed" if model == "net32" else "log10")
    gpu, ora = _gpu(tab), _ora(tab)
    gpu.set_option(capi.OPT_LINSOL, linsol)
    info = gpu.linsol_info()
    assert info["n_sparse"] + info["m_dense"] == info["nx"]
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"], "nllh")
    assert_grad_close(grad, ro["grad"], "gradient")
    res = gpu.condition_results()
    nst, nst_ref = res["stats"][:, 0], ro["stats"][:, 0]
    assert np.all(np.abs(nst - nst_ref) <= 0.1 * nst_ref + 5), (nst, nst_ref)
    assert np.all(res["stats"][:, 9] + res["stats"][:, 10] <= 0.02 * nst + 2)
    assert np.all(res["stats"][:, 16] == 0), "sparse factorisation fell back to the dense path"
    assert_llh_close(res["llh"], ro["llh"], "per-condition llh")
    for c in range(C):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    st, fval0, g0, _ = gpu.evaluate(theta, want_gradient=False)
    assert st == 0 and g0 is None
    assert_llh_close(fval0, ora.evaluate(theta, want_grad=False)["fval"], "nllh (cost only)")


def test_team_sparse_and_dense_agree_and_are_reproducible():
    """Same problem through both solvers: results agree far inside the tolerance, step counts nearly
    equal, and the sparse path is bit-reproducible run to run (no atomics in the numeric phase)."""
    tab = synth.make_problem("chen", 4, seed=5, timepoints=synth.DEFAULT_TIMEPOINTS["chen"][:3])
    theta = synth.random_theta(tab, seed=2, width=0.3)
    gs, gd = _gpu(tab), _gpu(tab)
    gs.set_option(capi.OPT_LINSOL, capi.LINSOL_SPARSE)
    gd.set_option(capi.OPT_LINSOL, capi.LINSOL_DENSE)
    a, a2, d = gs.evaluate(theta), gs.evaluate(theta), gd.evaluate(theta)
    assert a[0] == d[0] == 0
    assert a[1] == a2[1] and np.array_equal(a[2], a2[2])
    assert abs(a[1] - d[1]) <= 1e-8 * abs(d[1])
    assert_grad_close(a[2], d[2])
    ns_, nd_ = gs.condition_results()["stats"][:, 0], gd.condition_results()["stats"][:, 0]
    assert np.all(np.abs(ns_ - nd_) <= 0.05 * nd_ + 3)


def test_automatic_linear_solver_choice():
    """Sparse where it removes a real part of the dense work (Chen / Froehlich shapes), dense for small
    or nearly full patterns (net32)."""
    for model, kind in (("net32", 1), ("chen", 2), ("froehlich", 2)):
        tab = synth.make_problem(model, 1, seed=0, timepoints=[0.1])
        info = _gpu(tab).linsol_info()
        assert info["kind"] == kind, (model, info)
        if kind == 2:
            assert info["m_dense"] <= 0.75 * info["nx"] and info["m_dense"] % 2 == 0


# ---------------------------------------------------------------- A5 holes on the CUDA path
@pytest.mark.parametrize("model,C,kw", [("boehm", 6, dict(scale="mixed", sigma_mode="mixed")),
                                        ("steadystate_scaled", 5, dict(scale="log10", condition_specific=2)),
                                        ("net32", 5, dict(scale="log10"))])
def test_mapping_holes_and_overrides(oracle_available, model, C, kw):
    """mapping = -1 entries take their value from parameterOverrides (both scales) and leave the
    sensitivity list (nplist < np, plist compaction): the `ov_unscaled` / `nplist < np` device branches
    of both kernels against the oracle, per condition."""
    tab = synth.make_problem(model, C, seed=4, holes=2, **kw)
    assert np.any(tab.mapping < 0) and tab.overrides is not None
    gpu, ora = _gpu(tab), _ora(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    res = gpu.condition_results()
    assert_llh_close(res["llh"], ro["llh"])
    for c in range(C):
        npl = int(np.sum(tab.mapping[c] >= 0))
        assert npl < tab.np_
        assert_grad_close(res["sllh"][c][:npl], ro["sllh"][c][:npl], "sllh[%d]" % c)
        assert np.all(res["sllh"][c][npl:] == 0.0)
    # a subset with duplicates
    idx = [C - 1, 0, 0, 2]
    ro = ora.evaluate(theta, idx=idx)
    st, fval, grad, _ = gpu.evaluate(theta, idx)
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])


def test_unmapped_parameter_without_override_fails_like_the_reference(oracle_available):
    """mapping = -1 and no override table: the parameter is NaN (multiConditionDataProvider.cpp:160-166),
    the simulation fails, evaluate returns +inf / failure on both sides."""
    tab = synth.make_problem("boehm", 3, seed=4)
    tab.mapping[1, 2] = -1
    gpu, ora = _gpu(tab), _ora(tab)
    theta = synth.nominal_theta(tab)
    ro = ora.evaluate(theta)
    st, fval, _, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 1 and np.isinf(fval)
    assert gpu.condition_results()["status"][1] != 0


# ---------------------------------------------------------------- t = inf
def test_nonfinite_or_decreasing_output_times_are_rejected():
    # (t = +inf at the END of a condition's list is post-equilibration: tests/test_gpu_steadystate.py)
    for bad in ([0.1, -np.inf], [0.1, np.nan], [1.0, 0.5], [np.inf, 1.0]):
        tab = synth.make_problem("boehm", 2, seed=0, timepoints=np.array(bad))
        with pytest.raises(capi.B200LibraryError, match="output time"):
            _gpu(tab)
    tab = synth.make_problem("net32", 2, seed=0, timepoints=np.array([np.inf, 0.1]))
    with pytest.raises(capi.B200LibraryError, match="output time"):
        _gpu(tab)


# ---------------------------------------------------------------- partitioned evaluate returns the global result
def test_partitioned_evaluate_needs_a_collective():
    tab = synth.make_problem("boehm", 6, seed=1)
    f = _gpu(tab)
    f.set_partition(0, 2)
    with pytest.raises(capi.B200LibraryError, match="collective"):
        f.evaluate(synth.nominal_theta(tab))


@pytest.mark.parametrize("model,fail", [("boehm", False), ("net32", False), ("boehm", True)])
def test_partitioned_evaluate_with_host_hook_returns_global_result(model, fail):
    """Two handles (threads) with pb200_set_partition + the host all-reduce hook: evaluate returns the
    SAME global fval / gradient / status on both ranks; a failed condition on one rank makes BOTH ranks
    report +inf / failure (the reference master applies the rule after summing all workers' results)."""
    tab = synth.make_problem(model, 9, seed=2, timepoints=synth.DEFAULT_TIMEPOINTS[model][:4])
    theta = synth.random_theta(tab, seed=3, width=0.2)
    if fail:
        tab.mapping[7, 0] = -1   # unmapped, no override: NaN parameter -> that condition (rank 1's share) fails
    whole = _gpu(tab)
    ref = whole.evaluate(theta)
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

    def run(rank):
        f = _gpu(tab)
        f.set_partition(rank, world)
        f.set_allreduce(make_allreduce(rank))
        results[rank] = f.evaluate(theta)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for r in range(world):
        assert results[r] is not None
        assert results[r][0] == ref[0] == (1 if fail else 0)
        if fail:
            assert np.isinf(results[r][1])
        else:
            assert abs(results[r][1] - ref[1]) <= 1e-12 * abs(ref[1])
            np.testing.assert_allclose(results[r][2], ref[2], rtol=1e-11, atol=1e-12 * np.max(np.abs(ref[2])))
    if not fail:
        assert results[0][1] == results[1][1] and np.array_equal(results[0][2], results[1][2])


def test_nccl_in_library_two_gpus():
    """pb200_nccl_init: one handle per GPU (two threads of this process), one ncclAllReduce of
    (nllh, #errors, gradient) inside pb200_evaluate -- every rank returns the 1-GPU result."""
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    tab = synth.make_problem("boehm", 300, seed=2)
    theta = synth.random_theta(tab, seed=3, width=0.3)
    ref = _gpu(tab).evaluate(theta)
    from parpe_b200.problem import B200SummedGradientFunction
    uid = B200SummedGradientFunction.nccl_unique_id()
    world = 2
    results = [None] * world
    hier_tab = synth.fujita_hierarchical(24, seed=2)
    full = synth.random_theta(hier_tab, seed=5, width=0.2)
    outer = np.delete(full, hier_tab.hier.inner_idx)
    href = _gpu(hier_tab).hier_evaluate(outer, want_gradient=True)
    hres = [None] * world

    def run(rank):
        f = _gpu(tab, device=rank)
        f.nccl_init(uid, rank, world)
        results[rank] = f.evaluate(theta)
        results[rank] = f.evaluate(theta)  # (second call: cost-aware order in use)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for r in range(world):
        assert results[r] is not None and results[r][0] == 0
        assert abs(results[r][1] - ref[1]) <= 1e-12 * abs(ref[1])
        np.testing.assert_allclose(results[r][2], ref[2], rtol=1e-11, atol=1e-12 * np.max(np.abs(ref[2])))
    assert results[0][1] == results[1][1] and np.array_equal(results[0][2], results[1][2])

    uid2 = B200SummedGradientFunction.nccl_unique_id()

    def run_h(rank):
        f = _gpu(hier_tab, device=rank)
        f.nccl_init(uid2, rank, world)
        hres[rank] = f.hier_evaluate(outer, want_gradient=True)

    threads = [threading.Thread(target=run_h, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for r in range(world):
        assert hres[r] is not None and hres[r]["status"] == 0
        assert abs(hres[r]["fval"] - href["fval"]) <= 1e-10 * abs(href["fval"])
        np.testing.assert_allclose(hres[r]["grad"], href["grad"], rtol=1e-9, atol=1e-10 * np.max(np.abs(href["grad"])))


# ---------------------------------------------------------------- launch geometry / handles
def test_few_conditions_are_spread_over_all_sms(oracle_available):
    """n < (lane groups of a resident wave): the grid covers every SM with ceil(n / grid) busy groups per
    CTA instead of filling the first CTAs (the north-star size: 10^4 conditions / 8 GPUs = 1250 per GPU).
    Results per condition are the same numbers as in a full launch."""
    import ctypes as C
    tab = synth.make_problem("boehm", 1250, seed=6)
    theta = synth.random_theta(tab, seed=1, width=0.5)
    gpu = _gpu(tab)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == 0
    ll = gpu.last_launch()
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert ll["grid"] == min(sms, 1250), ll
    res = gpu.condition_results()
    # the same conditions inside a launch that fills the GPU: bit-identical per-condition results
    C_big = 8 * 1250
    tab_big = synth.make_problem("boehm", C_big, seed=6)
    tab_big.k[:1250] = tab.k
    tab_big.y[:tab.y.size] = tab.y
    big = _gpu(tab_big)
    big.evaluate(theta)
    rb = big.condition_results()
    assert np.array_equal(res["llh"], rb["llh"][:1250])
    assert np.array_equal(res["sllh"], rb["sllh"][:1250])
    sub = np.arange(0, 1250, 97)
    ro = _ora(tab).evaluate(theta, idx=sub)
    assert_llh_close(res["llh"][sub], ro["llh"])


def test_two_models_and_handles_in_one_process(oracle_available):
    """Handles of different models (different dynamic shared-memory sizes of the same kernels) created
    and used alternately: the opt-in shared-memory limit is set per launch, not cached per process."""
    tabs = [synth.make_problem("chen", 1, seed=1, timepoints=[0.1]),
            synth.make_problem("net32", 3, seed=1, timepoints=synth.DEFAULT_TIMEPOINTS["net32"][:3]),
            synth.make_problem("boehm", 3, seed=1), synth.make_problem("fujita", 3, seed=1)]
    gpus = [_gpu(t) for t in tabs]
    for rep in range(2):
        order = list(zip(tabs, gpus))
        if rep:
            order.reverse()
        for t, g in order:
            theta = synth.random_theta(t, seed=2, width=0.2)
            st, fval, grad, _ = g.evaluate(theta)
            assert st == 0
            if rep == 0 and t.model_name != "chen":
                ro = _ora(t).evaluate(theta)
                assert_llh_close(fval, ro["fval"])
    if capi.device_count() >= 2:  # the same kernels on a second device of the same process
        g1 = _gpu(tabs[2], device=1)
        theta = synth.random_theta(tabs[2], seed=2, width=0.2)
        a, b = g1.evaluate(theta), gpus[2].evaluate(theta)
        assert a[0] == b[0] == 0 and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_transfer_bytes_are_reported():
    tab = synth.make_problem("boehm", 64, seed=1)
    gpu = _gpu(tab)
    theta = synth.nominal_theta(tab)
    gpu.evaluate(theta)
    h2d, d2h = gpu.last_transfer_bytes()
    # first call: parameter table + index list + order refresh (step counts down, order up)
    assert h2d == 48 * tab.n_opt + 4 * 64 + 4 * 64 and d2h == 8 * (2 + tab.n_opt) + 8 * 64
    gpu.evaluate(theta)
    h2d, d2h = gpu.last_transfer_bytes()
    assert h2d == 48 * tab.n_opt and d2h == 8 * (2 + tab.n_opt)
