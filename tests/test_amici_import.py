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
    for bad, why in (("model_dirac", "events"), ("model_events", "events"),
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


# ---------------------------------------------------------------- the python-generated layout
PYGEN = [("pygen_steadystate", "steadystate"), ("pygen_boehm_w", "boehm")]


def _pygen_available(name):
    return os.path.exists(os.path.join(os.path.dirname(IMPORTED), name + ".cuh"))


def _pygen_problem(name, spec_model, **kw):
    tab = synth.make_problem(spec_model, 4, seed=2, **kw)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    tab.model_name = name  # same dimensions and parameter order: the generator was fed the repo's description
    return tab, theta


def test_amicis_python_generator_runs_here_and_its_output_is_imported(tmp_path):
    """AMICI's own python code generator (amici.de_export.DEExporter, run in place from the reference tree with
    libsbml / toposort / the SWIG extension stubbed) writes a model directory in the python-generated layout; the
    back-end re-homes it: macros, sparse blocks, the chain rule through the repeating elements."""
    from parpe_b200.codegen import amici_import, amici_pygen
    if not amici_pygen.available():
        pytest.skip("/root/reference not present")
    src = amici_pygen.generate("steadystate", True, root=str(tmp_path / "src"))
    assert os.path.exists(os.path.join(src, "dxdotdw.cpp")) and os.path.exists(os.path.join(src, "x0_fixedParameters.cpp"))
    dims = amici_import.parse_dimensions(open(os.path.join(src, "pygen_steadystate_w.h")).read())
    assert dims["nx_solver"] == 3 and dims["np"] == 5 and dims["nw"] > 0 and dims["ndxdotdw"] == dims["nw"]
    text = open(amici_import.import_model(src, outdir=str(tmp_path / "out"))).read()
    assert "struct Model_pygen_steadystate_w" in text and "NX = 3, NP = 5, NK = 4, NY = 3, NNZ = 9" in text
    assert "X0K_MASK = 0x7u" in text and "DYN_MASK = 0x1fu" in text
    assert "dxdotdw[" in text and "*dwdx[" in text and "*dwdp[" in text and "#define" in text and "#undef" in text


@pytest.mark.parametrize("name,spec_model", PYGEN)
@pytest.mark.parametrize("steady", [False, True], ids=["plain", "preeq+posteq"])
def test_host_core_with_a_python_generated_model_matches_the_reference_running_it(oracle_available, name, spec_model,
                                                                                  steady):
    """The generated model through the device core (host build) against the SAME generated model compiled as an
    amici::Model_ODE and run by the reference (pythonGenerated = true: Model_ODE assembles J and dxdotdp from the
    sparse blocks; x0_fixedParameters re-initialises after pre-equilibration)."""
    if not _pygen_available(name):
        pytest.skip("no generated model in this build")
    from oracle import binding as ob
    from tests.host_harness.binding import simulate_ss
    if not ob.lib().ora_has_model(name.encode()):
        pytest.skip("oracle built without the generated models")
    tab, theta = _pygen_problem(name, spec_model, preeq=steady, posteq=1 if steady else 0,
                                scale="mixed", sigma_mode="mixed")
    ro = ob.OracleProblem(tab).evaluate(theta)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p, plist, fac = condition_inputs(tab, theta, c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        pre = steady and tab.has_preeq[c]
        r = simulate_ss(name, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c),
                        k_preeq=tab.k_preeq[c] if pre else None, reinit_idx=tab.reinit_idx[c] if pre else ())
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])


@pytest.mark.gpu
@pytest.mark.parametrize("name,spec_model", PYGEN)
def test_cuda_path_with_a_python_generated_model(oracle_available, name, spec_model):
    from oracle import binding as ob
    from parpe_b200.problem import B200SummedGradientFunction
    if name not in capi.models() or not ob.lib().ora_has_model(name.encode()):
        pytest.skip("no generated model in this build")
    tab, theta = _pygen_problem(name, spec_model, preeq=True, posteq=1, scale="mixed", sigma_mode="mixed")
    ro = ob.OracleProblem(tab).evaluate(theta)
    st, fval, grad, _ = B200SummedGradientFunction(tab).evaluate(theta)
    assert st == ro["status"] == 0
    assert_llh_close(fval, ro["fval"])
    assert_grad_close(grad, ro["grad"])


# ---------------------------------------------------------------------------------------------------------------
# a model with a time-dependent input through the classic layout's spline functions: deps/AMICI/models/
# model_jakstat_adjoint (w[0] = amici::spline_pos(t, 5, 0.0, p[5], 5.0, p[6], ..., 0.0, 0.0); SURVEY H8 and the
# survey's offline stand-in for the Boehm configuration)
JAK_MODEL = "/root/reference/deps/AMICI/models/model_jakstat_adjoint"
JAK_IMPORTED = os.path.join(os.path.dirname(IMPORTED), "amici_jakstat_adjoint.cuh")
JAK_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imported", "jakstat_sensiforward.json")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_importer_maps_the_spline_functions(tmp_path):
    if not os.path.isdir(JAK_MODEL):
        pytest.skip("/root/reference not present")
    from parpe_b200.codegen import amici_import
    text = open(amici_import.import_model(JAK_MODEL, outdir=str(tmp_path))).read()
    assert "NX = 9, NP = 17, NK = 2, NY = 3, NNZ = 18" in text and '#include "../../amici_functions.cuh"' in text
    assert "w[0] = pb200::amifn::spline_pos(t,5,0.0,p[5],5.0,p[6],1.0E1,p[7],2.0E1,p[8],6.0E1,p[9],0.0,0.0);" in text
    assert "pb200::amifn::Dspline_pos(4,t,5," in text and "amici::" not in text
    # parameters the states depend on: rate constants, the initial value and the five spline nodes
    assert "DYN_MASK = 0x3ffu" in text


def test_spline_functions_match_the_reference(tmp_path):
    """csrc/amici_functions.cuh (variadic templates, the Forsythe-Malcolm-Moler interpolant written from the book)
    against the reference's own amici::spline / spline_pos / Dspline / Dspline_pos (C varargs,
    deps/AMICI/src/symbolic_functions.cpp:109-269, spline.cpp:16-257) compiled in place: 2 to 5 nodes, both end
    conditions, every node derivative, inside and outside the node range."""
    import subprocess
    lib = os.path.join(ROOT, "oracle", "_ref", "libamici_ref.a")
    if not os.path.isdir("/root/reference") or not os.path.exists(lib):
        pytest.skip("needs /root/reference and the reference build (oracle/_ref)")
    src = tmp_path / "check.cpp"
    src.write_text(r"""
#include <cmath>
#include <cstdio>
#include "amici/symbolic_functions.h"
#include "amici_functions.cuh"
#define BOTH(expr) { using namespace amici; const double a = expr; { using namespace pb200::amifn; const double b = expr; \
                     worst = fmax(worst, fabs(a - b) / fmax(fabs(a), 1e-12)); ++n; } }
int main() {
    double worst = 0.0; long n = 0;
    const double p[5] = {0.31, 1.7, 0.9, 0.05, 0.4};
    for (int ss = 0; ss <= 1; ++ss)
        for (double t = -3.0; t <= 90.0; t += 0.37) {
            const double e = ss, s = 0.3;
            BOTH(spline_pos(t, 5, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], 2.0E1, p[3], 6.0E1, p[4], e, s))
            BOTH(spline(t, 5, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], 2.0E1, p[3], 6.0E1, p[4], e, s))
            BOTH(spline(t, 4, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], 2.0E1, p[3], e, s))
            BOTH(spline(t, 3, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], e, s))
            BOTH(spline(t, 2, 0.0, p[0], 5.0, p[1], e, s))
            for (int id = 4; id <= 12; id += 2) {
                BOTH(Dspline_pos(id, t, 5, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], 2.0E1, p[3], 6.0E1, p[4], e, s))
                BOTH(Dspline(id, t, 5, 0.0, p[0], 5.0, p[1], 1.0E1, p[2], 2.0E1, p[3], 6.0E1, p[4], e, s))
            }
        }
    printf("%ld comparisons, worst relative difference %.3e\n", n, worst);
    return worst < 1e-12 ? 0 : 1;
}
""")
    A = "/root/reference/deps/AMICI"
    exe = str(tmp_path / "check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + A + "/include", "-I" + os.path.join(ROOT, "oracle", "_ref", "include"),
                           "-I" + A + "/ThirdParty/gsl", "-I" + A + "/ThirdParty/sundials/include",
                           "-I" + A + "/ThirdParty/sundials/include/sundials/amici_matlab",
                           "-I" + os.path.join(ROOT, "parpe_b200", "csrc"), str(src), lib, "-o", exe, "-fopenmp", "-ldl", "-lm"])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout


def _jakstat(corrector):
    import json
    from tests.util import jakstat_tables
    doc = json.load(open(JAK_FIXTURE))
    tab = jakstat_tables(doc)
    tab.ism = 1 if corrector == "simultaneous" else 0
    return doc, tab, np.array(doc["theta"])


@pytest.mark.parametrize("corrector", ["staggered", "simultaneous"])
def test_host_core_with_the_imported_spline_model_against_the_reference_fixture(corrector):
    """No oracle needed: the imported model_jakstat_adjoint in the host build of the device core against the committed
    outputs of the reference running its own vendored model on its own test inputs (tools/make_golden_jakstat.py)."""
    if not os.path.exists(JAK_IMPORTED):
        pytest.skip("no imported model in this build")
    from tests.host_harness.binding import simulate_ss
    doc, tab, theta = _jakstat(corrector)
    ref = doc["reference"][corrector]
    for c in range(tab.n_conditions):
        p, plist, fac = condition_inputs(tab, theta, c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate_ss("amici_jakstat_adjoint", p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c),
                        errconS=5 if corrector == "simultaneous" else 1)
        assert r["status"] == 0
        assert_llh_close(r["llh"], ref["llh"][c])
        assert_grad_close(r["sllh"] * fac, np.array(ref["sllh"][c])[:len(plist)])
        assert abs(r["numsteps"][-1] - ref["nst"][c]) <= 0.1 * ref["nst"][c] + 2


def test_reference_reproduces_the_spline_model_fixture(oracle_available):
    from oracle.binding import OracleProblem
    doc, tab, theta = _jakstat("staggered")
    try:
        ora = OracleProblem(tab, 2, 9, model_name="jakstat_vendored")
    except RuntimeError:
        pytest.skip("oracle built without the vendored model")
    r = ora.evaluate(theta)
    ref = doc["reference"]["staggered"]
    assert r["fval"] == pytest.approx(ref["nllh"], rel=1e-10)
    np.testing.assert_allclose(r["grad"], ref["grad"], rtol=1e-7, atol=1e-9 * np.max(np.abs(ref["grad"])))
