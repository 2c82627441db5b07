"""Known-answer tests of the reference for the hierarchical inner problem, restated against the
oracle's restatement (tests/parpeamici/hierarchicalOptimizationTest.cpp:201-223, 265-504, 580-590;
fixture tests/parpeamici/hierarchicalOptimizationTest.py:53-76)."""
import math

import numpy as np
import pytest

NAN = float("nan")
# fixture of hierarchicalOptimizationTest.py: rows (scaling idx, condition, observable)
MAP = [[0, 1, 0], [0, 2, 0], [1, 1, 1], [1, 2, 1], [1, 3, 1]]
MODEL_OUTPUT = [[1.0] * 6 for _ in range(4)]
MEASUREMENTS = [[NAN, 1.0, 1.0, 1.0, 1.0, 1.0], [2.0, 1.0, 1.0, 2.0, 1.0, NAN], [2.0, 1.0, 1.0, 2.0, NAN, 1.0],
                [1.0, 1.0, 1.0, NAN, 1.0, 1.0]]


def test_scaling_two(oracle_available):
    from oracle import binding as ob
    s = ob.hier_compute("scaling", 0, MAP, MODEL_OUTPUT, MEASUREMENTS, 3)
    assert math.log10(s) == math.log10(2.0)
    out = ob.hier_apply("scaling", 0, 2.0, MAP, MODEL_OUTPUT, 3)
    assert out[1][0] == MEASUREMENTS[1][0] and out[1][3] == MEASUREMENTS[1][3]
    assert out[2][0] == MEASUREMENTS[2][0] and out[2][3] == MEASUREMENTS[2][3]
    sig = [[1.0] * 6 for _ in range(4)]
    nllh = ob.hier_nllh(MEASUREMENTS, out, sig)
    assert nllh - 0.5 * math.log(2 * math.pi) * 20 == pytest.approx(0.0, abs=1e-10)


def test_compute_analytical_scalings(oracle_available):
    from oracle import binding as ob
    rows = [[0, 0, 0]]
    assert ob.hier_compute("scaling", 0, rows, [[1.0, 2.0, 3.0, 4.0]], [[10.0, 20.0, 30.0, 40.0]], 2) == 10.0
    # NaN measurements are skipped
    assert ob.hier_compute("scaling", 0, rows, [[2345.0, 2.0, 3.0, 4.0]], [[NAN, 20.0, 30.0, 40.0]], 2) == 10.0
    # zero denominator -> 1
    assert ob.hier_compute("scaling", 0, rows, [[0.0, 2.0, 0.0, 4.0]], [[10.0, 20.0, 30.0, 40.0]], 2) == 1.0


def test_compute_analytical_offsets(oracle_available):
    from oracle import binding as ob
    rows = [[0, 0, 0]]
    assert ob.hier_compute("offset", 0, rows, [[1.0, 2.0, 3.0, 4.0]], [[11.0, 12.0, 13.0, 14.0]], 2) == 10.0
    assert ob.hier_compute("offset", 0, rows, [[1.0, 2.0, 3.0, 4.0]], [[NAN, 12.0, NAN, 14.0]], 2) == 0.0


def test_compute_analytical_sigmas(oracle_available):
    from oracle import binding as ob
    rows = [[0, 0, 0]]
    s = ob.hier_compute("sigma", 0, rows, [[1.0, 2.0, 3.0, 4.0]], [[2.0, 2.0, 5.0, 4.0]], 2)
    assert s == pytest.approx(math.sqrt((1.0 + 4.0) / 2))
    # perfect fit: floored at max(1e-12, 0.01 * max|mes|)
    s0 = ob.hier_compute("sigma", 0, rows, [[1.0, 2.0, 3.0, 4.0]], [[1.0, 2.0, 3.0, 4.0]], 2)
    assert s0 == pytest.approx(0.03)


def test_apply_scaling_and_offset(oracle_available):
    from oracle import binding as ob
    rows = [[0, 0, 0]]
    assert ob.hier_apply("scaling", 0, 0.5, rows, [[2.0, 4.0, 6.0, 8.0]], 2) == [[1.0, 4.0, 3.0, 8.0]]
    assert ob.hier_apply("offset", 0, 5.0, rows, [[-4.0, 4.0, -2.0, 8.0]], 2) == [[1.0, 4.0, 3.0, 8.0]]


def test_splice_parameters(oracle_available):
    from oracle import binding as ob
    full = ob.hier_splice([1.0, 5.0, 8.0], [2, 3, 7], [0, 4], [6], [2.0, 3.0, 7.0], [0.0, 4.0], [6.0])
    assert list(full) == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]
    assert list(ob.hier_splice([0.0, 1.0, 2.0], [], [], [], [], [], [])) == [0.0, 1.0, 2.0]
    # defaults spliced for the pass-1 evaluation: log10 scalings -> 0
    assert list(ob.hier_splice([3.0, 2.0], [0, 1], [], [], [0.0, 0.0], [], [])) == [0.0, 0.0, 3.0, 2.0]


def test_fill_filtered_params(oracle_available):
    from oracle import binding as ob
    out = ob.hier_filter([9.0, 1.0, 2.0, 9.0, 9.0, 3.0, 4.0, 5.0, 9.0], [0, 3, 4, 8])
    assert list(out) == [1.0, 2.0, 3.0, 4.0, 5.0]


def test_likelihood_of_matching_data(oracle_available):
    from oracle import binding as ob
    data = [[1.0, 2.0, 3.0]]
    assert ob.hier_nllh(data, data, [[1.0, 1.0, 1.0]]) == 0.5 * math.log(2 * math.pi) * 3


def test_default_inner_parameters(oracle_available):
    from oracle.binding import lib
    L = lib()
    assert L.ora_hier_default_scaling(0) == 1.0 and L.ora_hier_default_scaling(2) == 0.0
    assert L.ora_hier_default_offset(0) == 0.0 and L.ora_hier_default_offset(2) == -math.inf  # quirk Q6


def test_hierarchical_evaluate_protocol(oracle_available):
    """With gradient: pass 2 at the spliced vector, gradient w.r.t. inner parameters ~ 0 at the
    analytic optimum; without: analytic nllh equals the full evaluation at the spliced vector."""
    from oracle.binding import OracleProblem
    from parpe_b200 import synth
    tab = synth.fujita_hierarchical(6, seed=0, sigma_mode="model")
    o = OracleProblem(tab)
    n_outer = o.hier_num_outer()
    assert n_outer == tab.n_opt - 6
    theta_full = synth.nominal_theta(tab)
    outer = np.delete(theta_full, tab.hier.inner_idx)
    r0 = o.hier_evaluate(outer, want_grad=False)
    r1 = o.hier_evaluate(outer, want_grad=True)
    assert r0["status"] == 0 and r1["status"] == 0
    np.testing.assert_allclose(r0["full_parameters"], r1["full_parameters"], rtol=0, atol=0)
    full_eval = o.evaluate(r1["full_parameters"])
    assert r0["fval"] == pytest.approx(full_eval["fval"], rel=1e-6)
    assert r1["fval"] == pytest.approx(full_eval["fval"], rel=1e-12)
    inner_grad = full_eval["grad"][tab.hier.inner_idx]
    scale = np.max(np.abs(full_eval["grad"]))
    assert np.all(np.abs(inner_grad) <= 1e-5 * max(scale, 1.0))
    np.testing.assert_allclose(r1["grad"], np.delete(full_eval["grad"], tab.hier.inner_idx), rtol=1e-12)
