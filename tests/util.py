"""Shared helpers for the parity tests."""
import numpy as np

LLH_RTOL = 1e-6    # north star: llh within 1e-6 relative
GRAD_RTOL = 1e-4   # north star: gradient within 1e-4 relative


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)


def assert_llh_close(a, b, what="llh"):
    e = float(np.max(rel_err(a, b)))
    assert e <= LLH_RTOL, "%s: relative error %.3e > %.1e (%r vs %r)" % (what, e, LLH_RTOL, a, b)


def assert_grad_close(a, b, what="gradient"):
    """Relative to the entry, with a floor of 1e-4 * max|gradient| so that entries that are
    zero up to integration error are compared on the gradient's own scale."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.max(np.abs(b)) if b.size else 0.0
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-4 * scale + 1e-300)
    assert np.all(e <= GRAD_RTOL), "%s: max relative error %.3e > %.1e\n got %r\n ref %r" % (
        what, float(np.max(e)), GRAD_RTOL, a, b)


def condition_inputs(tab, theta, c):
    """What the kernel prologue derives for condition c (A5): unscaled model parameters, the sensitivity
    parameter list, and the chain-rule factors d p / d (simulation-scale parameter) per plist entry."""
    from parpe_b200.tables import SCALE_LN, SCALE_LOG10
    p = np.empty(tab.np_)
    for i in range(tab.np_):
        j = tab.mapping[c, i]
        if j < 0:  # unmapped: the override value, given in the simulation scale
            v, sc = tab.overrides[c, i], tab.pscale_sim[c, i]
            p[i] = 10.0 ** v if sc == SCALE_LOG10 else (np.exp(v) if sc == SCALE_LN else v)
            continue
        p[i] = 10.0 ** theta[j] if tab.pscale_opt[j] == SCALE_LOG10 else (
            np.exp(theta[j]) if tab.pscale_opt[j] == SCALE_LN else theta[j])
    plist = [i for i in range(tab.np_) if tab.mapping[c, i] >= 0]
    fac = np.array([p[i] * np.log(10) if tab.pscale_sim[c, i] == SCALE_LOG10 else
                    (p[i] if tab.pscale_sim[c, i] == SCALE_LN else 1.0) for i in plist])
    return p, plist, fac


def assert_outputs_close(y, yref, tab, what="model outputs"):
    """Model outputs in the layout of the measurement table: rows of finite output times to 1e-6 relative; rows at
    t = inf (post-equilibration) to 2e-5 -- the reference's stopping rule (wrms(xdot) < 1 at the steady-state
    tolerances rtol 1e-6 / atol 1e-10) determines the steady state to about that tolerance only, and the step at
    which it fires depends on the step sequence (the reference's own KLU / dense variants differ likewise)."""
    y = np.asarray(y, dtype=float).reshape(-1, tab.ny)
    yref = np.asarray(yref, dtype=float).reshape(-1, tab.ny)
    inf_rows = np.isinf(tab.t)
    np.testing.assert_allclose(y[~inf_rows], yref[~inf_rows], rtol=1e-6, atol=1e-9, err_msg=what)
    if np.any(inf_rows):
        np.testing.assert_allclose(y[inf_rows], yref[inf_rows], rtol=2e-5, atol=1e-9, err_msg=what + " (t = inf)")


def jakstat_tables(doc):
    """ProblemTables of the fixture tests/golden/imported/jakstat_sensiforward.json (tools/make_golden_jakstat.py): the
    reference's own test inputs of deps/AMICI/models/model_jakstat_adjoint (np = 17 parameters, all log10; nk = 2
    compartment volumes; ny = 3 observables at 16 output times) as a multi-condition problem."""
    from parpe_b200.tables import ProblemTables
    nan = lambda rows: np.array([[np.nan if v is None else v for v in row] for row in rows], dtype=float)
    ts, Y, S = np.array(doc["ts"], dtype=float), nan(doc["Y"]), nan(doc["Sigma_Y"])
    C, nt = len(doc["conditions"]), len(ts)
    y, sd, k = [], [], []
    for cond in doc["conditions"]:
        yc, sc = Y.copy(), S.copy()
        if "data_sigma" in cond:
            sc[:] = cond["data_sigma"]
        if "drop_observable" in cond:
            yc[:, cond["drop_observable"]] = np.nan
        y.append(yc.ravel())
        sd.append(sc.ravel())
        k.append(np.array(doc["kappa"]) * np.array(cond["kappa_factor"]))
    return ProblemTables(
        model_name=doc["model"], np_=17, nk=2, ny=3, n_opt=17, t_offsets=np.arange(C + 1, dtype=np.int64) * nt,
        t=np.tile(ts, C), y=np.concatenate(y), sigma=np.concatenate(sd), k=np.concatenate(k),
        mapping=np.tile(np.arange(17, dtype=np.int32), C), pscale_sim=np.full(17 * C, 2, dtype=np.int32),
        pscale_opt=np.full(17, 2, dtype=np.int32), overrides=None, atol=1e-12, rtol=1e-8)
