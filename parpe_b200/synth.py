"""Synthetic measurement tables of the shapes named in BASELINE.json / SURVEY.md 8(d).

None of the PEtab benchmark problems is vendored in the reference tree, so every
workload is synthetic: model structure from ``parpe_b200.models``, condition
tables and measurements generated here with fixed seeds.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .models import get_spec
from .tables import SCALE_LOG10, SCALE_NONE, HierTables, ProblemTables

DEFAULT_TIMEPOINTS = {
    # examples/parpeamici/steadystate/create_steadystate_amici_model.py:107
    "steadystate": np.logspace(-5, 1, 20),
    "steadystate_scaled": np.logspace(-5, 1, 20),
    "boehm": np.array([0, 2.5, 5, 10, 15, 20, 30, 40, 50, 60, 80, 100, 120, 160, 200, 240], dtype=float),
    "fujita": np.array([0, 60, 120, 180, 300, 600, 900, 1200, 1800, 2400, 3000, 3600], dtype=float),
    # mass-action networks (nt = 8)
    "net32": np.array([0.1, 0.3, 1, 3, 10, 30, 100, 300], dtype=float),
    "net31": np.array([0.1, 0.3, 1, 3, 10, 30, 100, 300], dtype=float),
    "chen": np.array([0.1, 0.3, 1, 3, 10, 30, 100, 300], dtype=float),
    "froehlich": np.array([0.1, 0.3, 1, 3, 10, 30, 100, 300], dtype=float),
}


def condition_constants(model: str, C: int, rng: np.random.Generator, heterogeneous: bool = False) -> np.ndarray:
    spec = get_spec(model)
    k = np.tile(np.asarray(spec.k_nominal, dtype=float), (C, 1))
    if model == "steadystate":
        k[:, 3] = 1.0 + 0.1 * (np.arange(C) % 4)
    elif model == "steadystate_scaled":
        # condition table of create_steadystate_amici_model.py:436-441
        k[:, 0] = 1.0 + 0.1 * (np.arange(C) % 4)
    elif model == "fujita":
        k[:, 0] = np.array([0.1, 0.3, 1.0, 3.0, 10.0, 30.0])[np.arange(C) % 6]
    elif model in ("net32", "net31", "chen", "froehlich"):
        k[:, 0] = np.array([0.5, 1.0, 1.5, 2.0])[np.arange(C) % 4]
        k[:, 1] = 1.0 + 0.05 * rng.standard_normal(C)
    elif model == "boehm":
        if heterogeneous:
            # conditions that really differ (as the 6 doses of the Fujita problem span 300x): the constant is
            # log-uniform over 2.5 decades below its nominal value (larger values make the REFERENCE integrator
            # fail on ~10 % of the conditions: not a throughput workload), so step counts and Newton behaviour
            # vary across the warp
            k[:, 0] = 0.693 * 10.0 ** rng.uniform(-2.4, 0.08, C)
        else:
            k[:, 0] = 0.693 * (1.0 + 0.02 * rng.standard_normal(C))
    return k


def make_problem(model: str, n_conditions: int, *, scale: str = "log10", seed: int = 0,
                 timepoints: Optional[np.ndarray] = None, sigma_mode: str = "model",
                 simulate: Optional[Callable] = None, noise: float = 0.1,
                 missing_fraction: float = 0.0, ragged: bool = False,
                 condition_specific: int = 0, rtol: float = 1e-8, atol: float = 1e-12,
                 holes: int = 0, heterogeneous: bool = False, posteq: int = 0, preeq: bool = False,
                 **kw) -> ProblemTables:
    """Build tables for ``model`` with ``n_conditions`` conditions.

    scale: 'log10' | 'lin' | 'mixed' (optimiser scale of every parameter; the
           simulation scale equals it except in 'mixed', where odd parameters are
           simulated in linear scale from a log10 optimisation parameter).
    sigma_mode: 'model' (sigma table all NaN -> model sigma), 'data' (fixed 0.1*|y|+0.01),
                'mixed'.
    condition_specific: number of leading model parameters that get their own
           optimisation parameter per group of conditions (c % 2), as the
           offset-by-batch construction of the reference example does.
    holes: this many model parameters per condition (rotating with the condition index) are NOT mapped
           to an optimisation parameter (mapping = -1): they take their value from
           ``/parameters/parameterOverrides`` (a perturbed nominal value in the simulation scale) and drop
           out of the sensitivity list (multiConditionDataProvider.cpp:84-102,126-169,323-339).
    posteq: the last `posteq` output times of every condition become t = inf (post-equilibration: the data
           point is evaluated at the steady state, deps/AMICI/src/steadystateproblem.cpp).
    preeq: three of four conditions (c % 4 != 1) get a pre-equilibration condition -- the simulation
           condition's constants perturbed by up to 50 % -- and the re-initialisation list [0, 2][: c % 3]
           (multiConditionDataProvider.cpp:224-250).
    simulate(tables, theta) -> y_flat: optional callable producing noise-free
           outputs to derive measurements from; without it measurements are
           positive pseudo-data (throughput work is independent of them).
    """
    spec = get_spec(model)
    rng = np.random.default_rng(seed)
    C = n_conditions
    tp = DEFAULT_TIMEPOINTS[model] if timepoints is None else np.asarray(timepoints, dtype=float)
    nts = np.full(C, len(tp))
    if ragged:
        nts = np.maximum(1, len(tp) - (np.arange(C) % 3))
    t_offsets = np.concatenate([[0], np.cumsum(nts)]).astype(np.int64)

    def _times(n):
        tt = np.array(tp[:n], dtype=float)
        if posteq:
            tt[max(0, n - posteq):] = np.inf
        return tt
    t = np.concatenate([_times(n) for n in nts])
    np_, nk, ny = spec.np, spec.nk, spec.ny

    n_groups = 2 if condition_specific else 1
    n_opt = np_ + condition_specific * (n_groups - 1)
    mapping = np.tile(np.arange(np_, dtype=np.int32), (C, 1))
    if condition_specific:
        grp = np.arange(C) % n_groups
        for j in range(condition_specific):
            mapping[grp == 1, j] = np_ + j
    if scale == "lin":
        pscale_opt = np.full(n_opt, SCALE_NONE, dtype=np.int32)
    else:
        pscale_opt = np.full(n_opt, SCALE_LOG10, dtype=np.int32)
    pscale_sim = pscale_opt[np.maximum(mapping, 0)].astype(np.int32)
    if scale == "mixed":
        pscale_sim[:, 1::2] = SCALE_NONE
    k = condition_constants(model, C, rng, heterogeneous)
    overrides = None
    if holes:
        pn = np.asarray(spec.p_nominal, dtype=float)
        overrides = np.full((C, np_), np.nan)
        hrng = np.random.default_rng(seed + 1000)
        for c in range(C):
            for h_ in range(holes):
                i = (3 * c + 5 * h_ + 1) % np_
                lin = pn[i] * (1.0 + 0.2 * hrng.uniform(-1, 1))
                sc = pscale_sim[c, i]
                overrides[c, i] = np.log10(lin) if sc == SCALE_LOG10 else (np.log(lin) if sc == 1 else lin)
                mapping[c, i] = -1

    y = np.full(t.size * ny, 1.0)
    sigma = np.full(t.size * ny, np.nan)
    tab = ProblemTables(model_name=model, np_=np_, nk=nk, ny=ny, n_opt=n_opt, t_offsets=t_offsets, t=t, y=y,
                        sigma=sigma, k=k, mapping=mapping, pscale_sim=pscale_sim, pscale_opt=pscale_opt,
                        overrides=overrides, rtol=rtol, atol=atol, **kw)
    theta = nominal_theta(tab)
    if simulate is not None:
        ysim = np.asarray(simulate(tab, theta), dtype=float).reshape(-1)
        y = np.abs(ysim * (1.0 + noise * rng.standard_normal(ysim.size)))
    else:
        y = np.abs(1.0 + noise * rng.standard_normal(t.size * ny))
    if sigma_mode == "data":
        sigma = 0.1 * np.abs(y) + 0.01
    elif sigma_mode == "mixed":
        sigma = np.where(rng.random(y.size) < 0.5, 0.1 * np.abs(y) + 0.01, np.nan)
    if missing_fraction > 0:
        y = np.where(rng.random(y.size) < missing_fraction, np.nan, y)
    tab.y = np.ascontiguousarray(y)
    tab.sigma = np.ascontiguousarray(sigma)
    if preeq:
        prng = np.random.default_rng(seed + 2000)
        tab.has_preeq = np.array([1 if c % 4 != 1 else 0 for c in range(C)], dtype=np.int32)
        tab.k_preeq = tab.k * (1.0 + 0.5 * prng.uniform(-1, 1, size=tab.k.shape))
        tab.reinit_idx = [[0, 2][: (c % 3)] for c in range(C)]
    return tab


def nominal_theta(tab: ProblemTables) -> np.ndarray:
    """Optimisation-parameter vector at the model's nominal values."""
    spec = get_spec(tab.model_name)
    pn = np.asarray(spec.p_nominal, dtype=float)
    theta = np.zeros(tab.n_opt)
    for c in range(tab.n_conditions):
        for i in range(tab.np_):
            j = tab.mapping[c, i]
            if j >= 0:
                theta[j] = np.log10(pn[i]) if tab.pscale_opt[j] == SCALE_LOG10 else (
                    np.log(pn[i]) if tab.pscale_opt[j] == 1 else pn[i])
        if c >= 1 and tab.n_opt == tab.np_:
            break
    return theta


def random_theta(tab: ProblemTables, seed: int = 1, width: float = 1.0) -> np.ndarray:
    """'Random parameters in bounds': U(nominal - width, nominal + width) in log10 units
    (SURVEY.md 8(d)); linear-scale parameters are perturbed multiplicatively."""
    rng = np.random.default_rng(seed)
    th = nominal_theta(tab)
    u = rng.uniform(-width, width, size=th.size)
    out = th.copy()
    for j in range(th.size):
        if tab.pscale_opt[j] == SCALE_LOG10:
            out[j] = th[j] + u[j]
        elif tab.pscale_opt[j] == 1:
            out[j] = th[j] + u[j] * np.log(10)
        else:
            out[j] = th[j] * 10 ** u[j]
    return out


def fujita_hierarchical(n_conditions: int, seed: int = 0, **kw) -> ProblemTables:
    """Fujita-shape hierarchical problem (BASELINE config 2): 6 EGF doses tiled to
    ``n_conditions``; the three scaling factors and the three noise parameters are
    analytic inner parameters, one per observable, shared by all conditions."""
    tab = make_problem("fujita", n_conditions, scale="log10", seed=seed, **kw)
    spec = get_spec("fujita")
    names = [str(s) for s in spec.params]
    sc = [names.index(n) for n in ("scaleFactor_pEGFR", "scaleFactor_pAkt", "scaleFactor_pS6")]
    sg = [names.index(n) for n in ("sd_pEGFR_tot", "sd_pAkt_tot", "sd_pS6_tot")]
    C = n_conditions
    rows = lambda: np.array([[i, c, i] for i in range(3) for c in range(C)], dtype=np.int32)
    tab.hier = HierTables(scaling_idx=sc, offset_idx=[], sigma_idx=sg, scaling_map=rows(),
                          offset_map=np.zeros((0, 3), np.int32), sigma_map=rows())
    return tab


def steadystate_hierarchical(n_conditions: int, seed: int = 0, **kw) -> ProblemTables:
    """parPE's steadystate example with hierarchical optimisation (BASELINE config 0,
    examples/parpeamici/steadystate/create_steadystate_amici_model.py): one scaling parameter
    (obs_x1_scaled, log10), two batch-specific offsets (obs_x2_offsetted, linear; odd and even
    conditions), one sigma (obs_x1withsigma, log10)."""
    tab = make_problem("steadystate_scaled", n_conditions, scale="log10", seed=seed, condition_specific=0, **kw)
    spec = get_spec("steadystate_scaled")
    names = [str(s) for s in spec.params]
    i_scal = names.index("observableParameter1_obs_x1_scaled")
    i_off = names.index("observableParameter1_obs_x2_offsetted")
    i_sig = names.index("noiseParameter1_obs_x1withsigma")
    C = n_conditions
    n_opt = tab.np_ + 1
    mapping = tab.mapping.copy()
    mapping[1::2, i_off] = tab.np_  # second batch -> its own offset parameter
    pscale_opt = np.full(n_opt, SCALE_LOG10, dtype=np.int32)
    pscale_opt[[i_off, tab.np_]] = SCALE_NONE
    pscale_sim = np.full((C, tab.np_), SCALE_LOG10, dtype=np.int32)
    pscale_sim[:, i_off] = SCALE_NONE
    sigma = tab.sigma.reshape(-1, tab.ny).copy()
    sigma[:, :5] = 0.1      # sigmas from the data table, except the observable with a sigma parameter
    sigma[:, 5] = np.nan
    new = ProblemTables(model_name=tab.model_name, np_=tab.np_, nk=tab.nk, ny=tab.ny, n_opt=n_opt,
                        t_offsets=tab.t_offsets, t=tab.t, y=tab.y, sigma=sigma, k=tab.k, mapping=mapping,
                        pscale_sim=pscale_sim, pscale_opt=pscale_opt, rtol=tab.rtol, atol=tab.atol)
    even = [c for c in range(C) if c % 2 == 0]
    odd = [c for c in range(C) if c % 2 == 1]
    new.hier = HierTables(
        scaling_idx=[i_scal], offset_idx=[i_off, tab.np_], sigma_idx=[i_sig],
        scaling_map=[[0, c, 3] for c in range(C)],
        offset_map=[[0, c, 4] for c in even] + [[1, c, 4] for c in odd],
        sigma_map=[[0, c, 5] for c in range(C)])
    return new
