"""Back-end for AMICI-generated models (parpe_b200/codegen/amici_import.py): the reference's own
pre-generated model deps/AMICI/models/model_steadystate is re-homed as device code at build time and compared with
the SAME model run by the reference (oracle: the vendored Model_model_steadystate compiled in place) -- a chain that
does not touch the repo's own sympy generator."""
import os

import numpy as np
import pytest

from parpe_b200 import capi, synth
from tests.util import assert_grad_close, assert_llh_close, condition_inputs

REF_MODEL = "/root/reference/deps/AMICI/models/model_steadystate"
IMPORTED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "parpe_b200", "csrc", "models",
                        "imported", "amici_steadystate.cuh")


def _problem(**kw):
    tab = synth.make_problem("steadystate", 4, seed=2, sigma_mode="data", **kw)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    tab.model_name = "amici_steadystate"  # same dimensions and parameter order as the repo's restated model
    return tab, theta


def test_importer_reads_the_vendored_model(tmp_path):
    if not os.path.isdir(REF_MODEL):
        pytest.skip("/root/reference not present")
    from parpe_b200.codegen import amici_import
    out = amici_import.import_model(REF_MODEL, "amici_steadystate", outdir=str(tmp_path))
    text = open(out).read()
    assert "NX = 3, NP = 5, NK = 4, NY = 3, NNZ = 9" in text
    # (no x0_fixedParameters function in the classic layout: the reference re-initialises nothing for this model)
    assert "DYN_MASK = 0x1fu" in text and "X0K_MASK = 0x0u" in text
    # repeating elements are evaluated ahead of their users; the CSC Jacobian lands in the dense matrix
    assert "w[1] = x[0]*x[0];" in text and "J[8] = -k[3]-dwdx[1];" in text
    assert amici_import.parse_dimensions(open(os.path.join(REF_MODEL, "model_steadystate.h")).read())["nw"] == 2
    # models outside the supported scope are refused, not mistranslated
    for bad, why in (("model_dirac", "events"), ("model_jakstat_adjoint", "spline|events|not supported|no device counterpart"),
                     ("model_robertson", "not supported|not found|no device counterpart")):
        d = os.path.join(os.path.dirname(REF_MODEL), bad)
        if os.path.isdir(d):
            with pytest.raises(amici_import.ImportError_, match=why):
                amici_import.import_model(d, outdir=str(tmp_path))


def test_imported_model_is_registered_when_built():
    if not os.path.exists(IMPORTED):
        pytest.skip("the library was built without /root/reference: no imported model")
    assert capi.models()["amici_steadystate"] == dict(nx=3, np=5, nk=4, ny=3)


@pytest.mark.parametrize("steady", [False, True], ids=["plain", "preeq+posteq"])
def test_host_core_with_the_imported_model_matches_the_reference_running_its_own_model(oracle_available, steady):
    if not os.path.exists(IMPORTED):
        pytest.skip("no imported model in this build")
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_ss
    tab, theta = _problem(preeq=steady, posteq=2 if steady else 0)
    try:
        ora = OracleProblem(tab, model_name="steadystate_vendored")
    except RuntimeError:
        pytest.skip("oracle built without the vendored model")
    ro = ora.evaluate(theta)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p, plist, fac = condition_inputs(tab, theta, c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        pre = steady and tab.has_preeq[c]
        r = simulate_ss("amici_steadystate", p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c),
                        k_preeq=tab.k_preeq[c] if pre else None, reinit_idx=tab.reinit_idx[c] if pre else ())
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])


@pytest.mark.gpu
def test_cuda_path_with_the_imported_model(oracle_available):
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    if "amici_steadystate" not in capi.models():
        pytest.skip("no imported model in this build")
    tab, theta = _problem(preeq=True, posteq=1)
    try:
        ora = OracleProblem(tab, model_name="steadystate_vendored")
    except RuntimeError:
        pytest.skip("oracle built without the vendored model")
    gpu = B200SummedGradientFunction(tab)
    ro = ora.evaluate(theta)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])
    # the repo's own generator emits the same chemistry: the same results up to the expressions' rounding (without
    # pre-equilibration: the repo's model re-initialises states the classic-layout model has no function for)
    tab1, _ = _problem(posteq=1)
    tab2 = synth.make_problem("steadystate", 4, seed=2, sigma_mode="data", posteq=1)
    st1, f1, g1, _ = B200SummedGradientFunction(tab1).evaluate(theta)
    st2, f2, g2, _ = B200SummedGradientFunction(tab2).evaluate(theta)
    assert st1 == st2 == 0 and abs(f2 - f1) <= 1e-7 * abs(f1)
