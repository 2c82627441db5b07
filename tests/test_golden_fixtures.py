"""Committed reference outputs (tests/golden/*.json, made by tools/make_golden.py with the reference's
own AMICI/CVODES build) against: the oracle as built now, the host build of the device core, and --
on a GPU -- the CUDA path through the C ABI.  The GPU box has no /root/reference; these fixtures are
how reference outputs reach it."""
import glob
import json
import os

import numpy as np
import pytest

from parpe_b200 import synth
from parpe_b200.tables import SCALE_LN, SCALE_LOG10
from tests.util import assert_grad_close, assert_llh_close, assert_outputs_close

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.json")))


def _load(path):
    doc = json.load(open(path))
    tab = synth.make_problem(doc["model"], doc["n_conditions"], **doc["kwargs"])
    return doc, tab, np.array(doc["theta"])


def test_fixtures_exist():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_oracle_reproduces_fixture(oracle_available, path):
    from oracle.binding import OracleProblem
    doc, tab, theta = _load(path)
    if doc["case"] == "froehlich_full" and not os.environ.get("PB200_SLOW_TESTS"):
        pytest.skip("one nx = 1200 full-horizon solve takes > 4 min on a host core (set PB200_SLOW_TESTS=1); the fixture "
                    "comes from this same oracle build (tools/make_golden.py) and chen_full exercises the same path")
    if doc["case"].endswith("_full"):
        # (full-horizon large-model cases take minutes per condition on a host core: the first condition only)
        r = OracleProblem(tab).evaluate(theta, idx=[0])
        assert r["llh"][0] == pytest.approx(doc["llh"][0], rel=1e-10)
        np.testing.assert_allclose(r["sllh"][0], doc["sllh"][0], rtol=1e-7, atol=1e-9 * np.max(np.abs(doc["sllh"][0])))
        return
    r = OracleProblem(tab).evaluate(theta)
    assert r["fval"] == pytest.approx(doc["nllh"], rel=1e-10)
    np.testing.assert_allclose(r["grad"], doc["grad"], rtol=1e-7, atol=1e-9 * np.max(np.abs(doc["grad"])))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_host_core_against_fixture(path):
    """No oracle library needed: device algorithm (host build) vs committed reference outputs."""
    from tests.host_harness.binding import simulate, simulate_team
    from parpe_b200.models.netmodels import NET_MODELS
    doc, tab, theta = _load(path)
    if doc["case"] == "froehlich_full":
        pytest.skip("nx = 1200 over the full horizon takes tens of minutes on one host thread; the CUDA path "
                    "is checked against this fixture on the GPU")
    if doc["model"] in NET_MODELS:  # large-model core (bdf_team.cuh with the one-thread team)
        import functools
        # the full-horizon Chen-shape case runs with the sparse linear solver (as the CUDA path does by default)
        simulate = functools.partial(simulate_team, sparse=True) if doc["case"] == "chen_full" else simulate_team
    steady = tab.has_preeq is not None or bool(np.any(np.isinf(tab.t)))
    for c in range(1 if doc["case"].endswith("_full") else tab.n_conditions):
        p = np.empty(tab.np_)
        for i in range(tab.np_):
            j = tab.mapping[c, i]
            if j < 0:  # unmapped: the override value, given in the simulation scale
                v, sc = tab.overrides[c, i], tab.pscale_sim[c, i]
                p[i] = 10.0 ** v if sc == SCALE_LOG10 else (np.exp(v) if sc == SCALE_LN else v)
                continue
            lin = 10.0 ** theta[j] if tab.pscale_opt[j] == SCALE_LOG10 else (
                np.exp(theta[j]) if tab.pscale_opt[j] == SCALE_LN else theta[j])
            p[i] = lin
        plist = [i for i in range(tab.np_) if tab.mapping[c, i] >= 0]
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        if steady:  # the steady-state instantiation of the core (pre-/post-equilibration)
            from tests.host_harness.binding import simulate_ss
            pre = tab.has_preeq is not None and tab.has_preeq[c]
            r = simulate_ss(doc["model"], p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c),
                            k_preeq=tab.k_preeq[c] if pre else None, reinit_idx=tab.reinit_idx[c] if pre else ())
        else:
            r = simulate(doc["model"], p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c))
        assert r["status"] == 0
        assert_llh_close(r["llh"], doc["llh"][c])
        fac = np.array([p[i] * np.log(10) if tab.pscale_sim[c, i] == SCALE_LOG10 else
                        (p[i] if tab.pscale_sim[c, i] == SCALE_LN else 1.0) for i in plist])
        assert_grad_close(r["sllh"] * fac, np.array(doc["sllh"][c])[:len(plist)])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_cuda_path_against_fixture(path):
    from parpe_b200.problem import B200SummedGradientFunction
    doc, tab, theta = _load(path)
    gpu = B200SummedGradientFunction(tab)
    st, fval, grad, _ = gpu.evaluate(theta)
    assert st == 0
    assert_llh_close(fval, doc["nllh"])
    assert_grad_close(grad, doc["grad"])
    res = gpu.condition_results()
    assert_llh_close(res["llh"], doc["llh"])
    for c in range(tab.n_conditions):
        assert_grad_close(res["sllh"][c], doc["sllh"][c], "per-condition sllh[%d]" % c)
    if doc.get("nst") is not None:  # step counts stay close to the reference's (a wrong factorisation converges, but slower)
        nst_ref = np.array(doc["nst"])
        tol = 0.2 if "preeq_steps" in doc and max(doc["preeq_steps"]) > 0 else 0.1  # (after a pre-equilibration the
        # main simulation starts near a steady state and its step sequence is more sensitive: DESIGN 6)
        assert np.all(np.abs(res["stats"][:, 0] - nst_ref) <= tol * nst_ref + 5), (res["stats"][:, 0], nst_ref)
        assert np.all(res["stats"][:, 16] == 0)  # no sparse-factorisation fallback
    if doc["case"].endswith("_full") and doc["model"] == "froehlich":
        return  # (the subset / output passes repeat the same integration: skipped for the nx = 1200 case)
    st, fs, gs, _ = gpu.evaluate(theta, doc["subset"])
    assert_llh_close(fs, doc["subset_nllh"])
    assert_grad_close(gs, doc["subset_grad"])
    rc, y, s = gpu.getModelOutputsAndSigmas(theta)
    assert rc == 0
    yref = np.array([np.nan if v is None else v for v in doc["y"]])
    assert_outputs_close(y, yref, tab)
    np.testing.assert_allclose(s, np.array([np.nan if v is None else v for v in doc["sigmay"]]), rtol=1e-12)
