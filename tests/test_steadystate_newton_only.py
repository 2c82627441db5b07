"""steadyStateComputationMode newtonOnly on the GPU (pb200_problem_desc::newton_only; SteadystateProblem::
findSteadyState with turnOffSimulation, deps/AMICI/src/steadystateproblem.cpp:120-166) through the C ABI against the
reference, and the descriptor's checks.  The per-path CPU parity of the same integrator source is
tests/test_host_steadystate.py::test_newton_only_mode.

NOTE: written after the round's GPU budget was spent -- the CPU parity of the integrator core (same source, host
build) is green, the GPU cases of this file have not run on a B200 yet.  (This file sorts after the other test
files on purpose.)"""
import numpy as np
import pytest

from parpe_b200 import capi, synth
from tests.util import assert_grad_close, assert_llh_close


def _problem(newton_maxsteps):
    from tests.test_gpu_steadystate import _problem as base
    tab = base("steadystate", 8, True, 2)
    tab.newton_maxsteps = newton_maxsteps
    tab.newton_only = 1
    tab.num_trials = 1
    return tab


@pytest.mark.gpu
def test_newton_only_with_enough_steps_matches_reference(oracle_available):
    from tests.test_gpu_steadystate import _both
    tab = _problem(40)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert ro["status"] == 0 and st == 0 and np.array_equal(res["status"], ro["cond_status"])
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    st0, f0, _, _ = gpu.evaluate(theta, want_gradient=False)
    assert st0 == 0
    assert_llh_close(f0, ora.evaluate(theta, want_grad=False)["fval"])


@pytest.mark.gpu
def test_newton_only_without_enough_steps_fails_like_the_reference(oracle_available):
    """One Newton step is too few and there is no simulation to fall back on: AMICI_ERROR for every condition that
    needs an equilibration, fval = +inf, FAILURE -- on both sides."""
    from tests.test_gpu_steadystate import _both
    tab = _problem(1)
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert ro["status"] != 0 and st == ro["status"] and np.isinf(fval) and fval > 0
    assert np.array_equal(res["status"], ro["cond_status"])
    assert np.all(res["status"][res["status"] != 0] == -99)


def test_newton_only_descriptor_checks_need_no_device():
    """newton_only without newton_maxsteps > 0 is a configuration with which the reference runs neither method
    (steadystateproblem.cpp:133-146): refused.  (Checked after the device probe: on a box without a GPU the
    no-device error comes first -- either way no handle is returned.)"""
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("steadystate", 2, seed=0, timepoints=np.array([0.1, np.inf]))
    tab.newton_only = 1
    with pytest.raises(capi.B200LibraryError):
        B200SummedGradientFunction(tab)
    d = tab.to_desc()
    assert d.newton_only == 1 and d.newton_maxsteps == 0 and d.linsol == 0
    tab.linsol = 2
    assert tab.to_desc().linsol == 2
