"""Fujita_SciSignal2010-*shape* EGF/Akt/S6 signalling model (BASELINE config 2).

Stand-in with the published model's dimensions (9 states; 16 kinetic + 3
scaling parameters, plus one noise parameter per observable for the
hierarchical sigma reduction; 3 observables; EGF dose as the condition-specific
constant).  The real PEtab problem is not vendored; measurement tables are
synthetic.
"""
import sympy as sp

from ..codegen.modelgen import ModelSpec


def spec():
    t = sp.Symbol("t", real=True)
    EGFR, pEGFR, pEGFR_Akt, Akt, pAkt, S6, pAkt_S6, pS6, EGF_EGFR = x = sp.symbols(
        "EGFR pEGFR pEGFR_Akt Akt pAkt S6 pAkt_S6 pS6 EGF_EGFR", real=True)
    names = ("EGFR_turnover init_AKT init_EGFR init_S6 reaction_1_k1 reaction_1_k2 reaction_2_k1 "
             "reaction_2_k2 reaction_3_k1 reaction_4_k1 reaction_5_k1 reaction_5_k2 reaction_6_k1 "
             "reaction_7_k1 reaction_8_k1 reaction_9_k1 scaleFactor_pEGFR scaleFactor_pAkt "
             "scaleFactor_pS6 sd_pEGFR_tot sd_pAkt_tot sd_pS6_tot")
    p = sp.symbols(names, real=True)
    (turn, iAkt, iEGFR, iS6, r1k1, r1k2, r2k1, r2k2, r3k1, r4k1, r5k1, r5k2, r6k1, r7k1, r8k1, r9k1,
     sEGFR, sAkt, sS6, sdE, sdA, sdS) = p
    EGF = sp.Symbol("EGF", real=True)
    v1 = r1k1 * EGF * EGFR - r1k2 * EGF_EGFR
    v2 = r2k1 * pEGFR * Akt - r2k2 * pEGFR_Akt
    v3 = r3k1 * pEGFR_Akt
    v4 = r4k1 * pEGFR
    v5 = r5k1 * pAkt * S6 - r5k2 * pAkt_S6
    v6 = r6k1 * pAkt_S6
    v7 = r7k1 * pAkt
    v8 = r8k1 * pS6
    v9 = r9k1 * EGF_EGFR
    xdot = [
        turn * (iEGFR - EGFR) - v1,
        v9 - v2 + v3 - v4,
        v2 - v3,
        -v2 + v7,
        v3 - v5 + v6 - v7,
        -v5 + v8,
        v5 - v6,
        v6 - v8,
        v1 - v9,
    ]
    x0 = [iEGFR, 0, 0, iAkt, 0, iS6, 0, 0, 0]
    obs = [sEGFR * (pEGFR + pEGFR_Akt), sAkt * (pAkt + pAkt_S6), sS6 * pS6]
    return ModelSpec(
        name="fujita", states=list(x), params=list(p), consts=[EGF], t=t,
        x0=[sp.sympify(e) for e in x0], xdot=xdot, observables=obs, sigmay=[sdE, sdA, sdS],
        p_nominal=[1.06e-3, 0.0433, 3.5e4, 3.55, 3.3e-3, 4.0e-2, 1.55e-5, 0.2, 0.45, 0.031,
                   2.1e-2, 1.2e-3, 5.1e-3, 0.016, 6.9e-4, 0.027, 1.8e-5, 60.0, 8.0e4, 0.05, 0.05, 0.05],
        k_nominal=[1.0], doc=__doc__)
