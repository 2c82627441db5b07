"""Boehm_JProteomeRes2014-*shape* STAT5 dimerisation model (BASELINE config 1).

The real PEtab problem is not vendored in the reference tree and cannot be
imported offline (no libsbml/petab); this is a stand-in with the published
model's dimensions and structure (8 states, 9 estimated parameters, 3
observables, explicit time dependence through the decaying Epo input), written
from the published reaction scheme.  Measurement tables used with it are
synthetic.  The two fixed quantities of the PEtab problem (``ratio``,
``specC17``) are constants k[0], k[1].
"""
import sympy as sp

from ..codegen.modelgen import ModelSpec


def spec():
    t = sp.Symbol("t", real=True)
    STAT5A, STAT5B, pApB, pApA, pBpB, nucpApA, nucpApB, nucpBpB = x = sp.symbols(
        "STAT5A STAT5B pApB pApA pBpB nucpApA nucpApB nucpBpB", real=True)
    (Epo_deg, k_exp_hetero, k_exp_homo, k_imp_hetero, k_imp_homo, k_phos,
     sd_pSTAT5A, sd_pSTAT5B, sd_rSTAT5A) = p = sp.symbols(
        "Epo_degradation_BaF3 k_exp_hetero k_exp_homo k_imp_hetero k_imp_homo k_phos "
        "sd_pSTAT5A_rel sd_pSTAT5B_rel sd_rSTAT5A_rel", real=True)
    ratio, specC17 = k = sp.symbols("ratio specC17", real=True)
    cyt, nuc = sp.Float(1.4), sp.Float(0.45)
    epo = sp.Float(1.25e-7) * sp.exp(-Epo_deg * t)
    v1 = cyt * epo * STAT5A**2 * k_phos
    v2 = cyt * epo * STAT5A * STAT5B * k_phos
    v3 = cyt * epo * STAT5B**2 * k_phos
    v4 = cyt * k_imp_homo * pApA
    v5 = cyt * k_imp_hetero * pApB
    v6 = cyt * k_imp_homo * pBpB
    v7 = nuc * k_exp_homo * nucpApA
    v8 = nuc * k_exp_hetero * nucpApB
    v9 = nuc * k_exp_homo * nucpBpB
    xdot = [
        (-2 * v1 - v2 + 2 * v7 + v8) / cyt,
        (-v2 - 2 * v3 + v8 + 2 * v9) / cyt,
        (v2 - v5) / cyt,
        (v1 - v4) / cyt,
        (v3 - v6) / cyt,
        (v4 - v7) / nuc,
        (v5 - v8) / nuc,
        (v6 - v9) / nuc,
    ]
    x0 = [sp.Float(207.6) * ratio, sp.Float(207.6) - sp.Float(207.6) * ratio, 0, 0, 0, 0, 0, 0]
    obs = [
        (100 * pApB + 200 * pApA * specC17) / (pApB + STAT5A * specC17 + 2 * pApA * specC17),
        -(100 * pApB - 200 * pBpB * (specC17 - 1)) / ((STAT5B * (specC17 - 1) - pApB) + 2 * pBpB * (specC17 - 1)),
        (100 * pApB + 100 * STAT5A * specC17 + 200 * pApA * specC17)
        / (2 * pApB + STAT5A * specC17 + 2 * pApA * specC17 - STAT5B * (specC17 - 1) - 2 * pBpB * (specC17 - 1)),
    ]
    return ModelSpec(
        name="boehm", states=list(x), params=list(p), consts=list(k), t=t,
        x0=[sp.sympify(e) for e in x0], xdot=xdot, observables=obs,
        sigmay=[sd_pSTAT5A, sd_pSTAT5B, sd_rSTAT5A],
        p_nominal=[0.0269738, 1.00094e-5, 0.00617, 0.0163708, 96945.5, 15766.8, 3.85, 6.59, 3.15],
        k_nominal=[0.693, 0.107], doc=__doc__)
