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
