"""GPU cases of what was added after the round's GPU budget was spent (the CPU parity of the same integrator /
model source is green; NONE of the `gpu`-marked cases of this file has run on a B200 yet -- the file sorts last so
that the driver's `-x` run reaches them after everything that was verified).

The steady-state modes `newtonOnly` on the GPU through the C ABI against the reference, and the descriptor's
checks: steadyStateComputationMode newtonOnly (pb200_problem_desc::newton_only; SteadystateProblem::findSteadyState
with turnOffSimulation, deps/AMICI/src/steadystateproblem.cpp:120-166) and steadyStateSensitivityMode newtonOnly
(pb200_problem_desc::steady_sens_newton; the simulation to steady state runs without the sensitivity columns,
:728-736, which are solved for at the steady state, :91-103).  The per-path CPU parity of the same integrator source
is tests/test_host_steadystate.py::test_newton_only_mode / test_sensitivity_mode_newton_only.

Further down: the simultaneous corrector (pb200_problem_desc::ism = 1) and the imported model with a spline input
(deps/AMICI/models/model_jakstat_adjoint through csrc/amici_functions.cuh) against its committed reference fixture."""
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


@pytest.mark.gpu
@pytest.mark.parametrize("newton_maxsteps", [0, 2], ids=["simulation", "newton-fails-then-simulation"])
def test_sensitivities_of_steady_states_by_the_linear_solve(oracle_available, newton_maxsteps):
    """steady_sens_newton: pre- and post-equilibration simulated without the sensitivity columns (fewer steps than
    with them in the error test -- the reference's counts), sx = -J^-1 dxdotdp at each steady state."""
    from tests.test_gpu_steadystate import _both, _problem as base
    tab = base("steadystate_scaled", 6, True, 2, scale="mixed", sigma_mode="mixed")
    tab.steady_sens_newton = 1
    tab.newton_maxsteps = newton_maxsteps
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert ro["status"] == 0 and st == 0 and np.all(res["status"] == 0)
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    # the plain mode integrates the sensitivities along and needs more steps for the same problem
    tab.steady_sens_newton = 0
    tab.newton_maxsteps = 0
    plain, _ = _both(tab)
    st2, f2, g2, _ = plain.evaluate(theta)
    assert st2 == 0
    assert_llh_close(f2, fval)
    nst, nst_plain = gpu.condition_results()["stats"][:, 0], plain.condition_results()["stats"][:, 0]
    assert np.sum(nst) < np.sum(nst_plain)


@pytest.mark.gpu
def test_large_model_kernel_refuses_the_linear_solve_for_steady_state_sensitivities():
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("net32", 2, seed=0, timepoints=np.array([0.1, np.inf]))
    tab.steady_sens_newton = 1
    with pytest.raises(capi.B200LibraryError, match="steady_sens_newton"):
        B200SummedGradientFunction(tab)
    # without steady states in the problem the flag has no role
    tab = synth.make_problem("net32", 2, seed=0)
    tab.steady_sens_newton = 1
    assert B200SummedGradientFunction(tab).evaluate(synth.random_theta(tab, seed=1, width=0.3))[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("model,C,kw", [("boehm", 40, dict(scale="log10")),
                                        ("steadystate_scaled", 12, dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.2, ragged=True))],
                         ids=["boehm", "steadystate_scaled"])
def test_simultaneous_corrector_matches_reference(oracle_available, model, C, kw):
    """pb200_problem_desc::ism = 1: the simultaneous corrector (CV_SIMULTANEOUS) of the full-feature kernel
    instantiation against the reference run with InternalSensitivityMethod::simultaneous; CPU parity of the same
    source: tests/test_host_core.py::test_simultaneous_corrector_matches_cvodes."""
    from tests.test_gpu_steadystate import _both
    tab = synth.make_problem(model, C, seed=3, **kw)
    tab.ism = 1
    gpu, ora = _both(tab)  # (the oracle honours the descriptor's ism)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert ro["status"] == 0 and st == 0 and np.all(res["status"] == 0)
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    # CVODES' counters of the simultaneous corrector: no separate sensitivity iterations / error-test failures
    assert np.all(res["stats"][:, 5] == 0) and np.all(res["stats"][:, 8] == 0)
    nst, nst_ref = res["stats"][:, 0], ro["stats"][:, 0]
    assert abs(np.sum(nst) - np.sum(nst_ref)) <= 0.05 * np.sum(nst_ref)  # (per condition rounding decides: a few per cent either way)
    st0, f0, _, _ = gpu.evaluate(theta, want_gradient=False)
    assert st0 == 0
    assert_llh_close(f0, ora.evaluate(theta, want_grad=False)["fval"])


@pytest.mark.gpu
@pytest.mark.parametrize("linsol", [1, 2], ids=["dense", "sparse"])
def test_simultaneous_corrector_of_the_large_model_kernel(oracle_available, linsol):
    """ism = 1 on a table-driven network model (one thread block per condition): BdfTeam::nls_simultaneous of the
    steady-state instantiation, both linear solvers, against the reference run with the simultaneous corrector; CPU
    parity of the same source with a one-thread team: tests/test_host_team.py::test_simultaneous_corrector_team."""
    from tests.test_gpu_steadystate import _both
    tab = synth.make_problem("net32", 6, seed=3, scale="log10")
    tab.ism = 1
    tab.linsol = linsol
    gpu, ora = _both(tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    st, fval, grad, _ = gpu.evaluate(theta)
    res = gpu.condition_results()
    assert ro["status"] == 0 and st == 0 and np.all(res["status"] == 0)
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], ro["sllh"][c], "per-condition sllh[%d]" % c)
    assert np.all(res["stats"][:, 5] == 0) and np.all(res["stats"][:, 8] == 0)  # no nniS / netfS
    assert abs(np.sum(res["stats"][:, 0]) - np.sum(ro["stats"][:, 0])) <= 0.05 * np.sum(ro["stats"][:, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("corrector", ["staggered", "simultaneous"])
def test_imported_model_with_a_spline_input_against_the_reference_fixture(corrector):
    """deps/AMICI/models/model_jakstat_adjoint (w[0] = amici::spline_pos(t, 5, ...): a time-dependent input through the
    classic layout's spline functions, csrc/amici_functions.cuh) as device code, on the reference's own test inputs,
    against the outputs of the reference running its own vendored model (tests/golden/imported/
    jakstat_sensiforward.json, tools/make_golden_jakstat.py); CPU parity of the same generated model source:
    tests/test_amici_import.py."""
    import json
    import os
    from parpe_b200.problem import B200SummedGradientFunction
    from tests.util import jakstat_tables
    if "amici_jakstat_adjoint" not in capi.models():
        pytest.skip("the library was built without /root/reference: no imported model")
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imported", "jakstat_sensiforward.json")))
    tab = jakstat_tables(doc)
    tab.ism = 1 if corrector == "simultaneous" else 0
    ref = doc["reference"][corrector]
    gpu = B200SummedGradientFunction(tab)
    st, fval, grad, _ = gpu.evaluate(np.array(doc["theta"]))
    res = gpu.condition_results()
    assert st == 0 and np.all(res["status"] == 0)
    assert_llh_close(fval, ref["nllh"])
    assert_grad_close(grad, np.array(ref["grad"]))
    assert_llh_close(res["llh"], np.array(ref["llh"]))
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], np.array(ref["sllh"][c]), "per-condition sllh[%d]" % c)
    assert abs(np.sum(res["stats"][:, 0]) - np.sum(ref["nst"])) <= 0.1 * np.sum(ref["nst"])
