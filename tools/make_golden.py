#!/usr/bin/env python
"""Generate tests/golden/*.json with the reference itself: AMICI 0.31.0 / CVODES 7.1.1 compiled in
place from /root/reference (oracle/_ref) driven by the restated parPE glue.  Run in the build
container (needs /root/reference); the fixtures are committed so that the GPU box, which has no
/root/reference, can check the CUDA path against reference outputs directly.

    python tools/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle.binding import ISM_STAGGERED, LINSOL_KLU, OracleProblem  # noqa: E402
from parpe_b200 import synth  # noqa: E402

CASES = {
    "steadystate_lin": ("steadystate", 4, dict(scale="lin", seed=11, sigma_mode="data")),
    "steadystate_scaled_mixed": ("steadystate_scaled", 6, dict(scale="mixed", seed=12, sigma_mode="mixed",
                                                               missing_fraction=0.2, ragged=True,
                                                               condition_specific=2)),
    "boehm_log10": ("boehm", 4, dict(scale="log10", seed=13)),
    "fujita_log10": ("fujita", 6, dict(scale="log10", seed=14)),
    # large-model path (table-driven mass-action network, one CTA per condition)
    "net32_mixed": ("net32", 4, dict(scale="mixed", seed=15, sigma_mode="mixed", missing_fraction=0.1)),
    # nx = 500, nplist = 50 (Chen shape), first two output times
    "chen_log10": ("chen", 1, dict(scale="log10", seed=16, timepoints=[0.1, 0.3])),
    # A5 holes: unmapped model parameters (mapping = -1) take /parameters/parameterOverrides and leave the
    # sensitivity list (nplist < np); small-model and large-model kernels
    "boehm_holes_mixed": ("boehm", 5, dict(scale="mixed", seed=17, holes=2, sigma_mode="mixed")),
    "net32_holes": ("net32", 4, dict(scale="log10", seed=18, holes=2)),
    # the FULL horizon of the large-model workloads (all 8 output times to t = 300: the stiff part, hundreds of
    # linear-solver setups per solve), two conditions each -- what bench.py integrates
    "chen_full": ("chen", 2, dict(scale="log10", seed=19)),
    "froehlich_full": ("froehlich", 2, dict(scale="log10", seed=20)),
    # steady states: output times t = inf (post-equilibration) and pre-equilibration with
    # re-initialisation of states, deps/AMICI/src/steadystateproblem.cpp
    "steadystate_posteq": ("steadystate", 4, dict(scale="lin", seed=21, sigma_mode="data", posteq=2,
                                                  timepoints=[0.001, 0.01, 0.1, 1.0, 10.0, 100.0, 1000.0])),
    "steadystate_scaled_preeq": ("steadystate_scaled", 6, dict(scale="mixed", seed=22, sigma_mode="mixed",
                                                               missing_fraction=0.2, ragged=True, preeq=True,
                                                               condition_specific=2)),
    "boehm_preeq_posteq": ("boehm", 5, dict(scale="log10", seed=23, preeq=True, posteq=1,
                                            timepoints=[0.0, 2.5, 5.0, 10.0, 20.0, 40.0, 80.0])),
}


def kw_has_inf(kw):
    return bool(kw.get("posteq"))


def build_case(name):
    model, C, kw = CASES[name]
    tab = synth.make_problem(model, C, **kw)
    theta = synth.random_theta(tab, seed=5, width=0.3)
    return tab, theta


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    for name in (sys.argv[1:] or CASES):
        tab, theta = build_case(name)
        o = OracleProblem(tab, ISM_STAGGERED, LINSOL_KLU)  # AMICI's default linear solver
        r = o.evaluate(theta, collect_stats=True)
        sub = [tab.n_conditions - 1, 0, 0]
        rs = o.evaluate(theta, idx=sub)
        rc, y, s = o.get_model_outputs(theta)
        assert r["status"] == 0 and rc == 0
        doc = dict(
            generator="tools/make_golden.py: reference AMICI 0.31.0 + CVODES 7.1.1 (KLU, staggered) built from "
                      "/root/reference, rtol 1e-8, atol 1e-12",
            case=name, model=CASES[name][0], n_conditions=CASES[name][1], kwargs=CASES[name][2],
            theta=theta.tolist(), nllh=r["fval"], grad=r["grad"].tolist(), llh=r["llh"].tolist(),
            sllh=r["sllh"].tolist(), subset=sub, subset_nllh=rs["fval"], subset_grad=rs["grad"].tolist(),
            # (with t = inf points the reference reports 0 steps at the last output time: no comparable count)
            nst=None if kw_has_inf(CASES[name][2]) else r["stats"][:, 0].tolist(),
            preeq_steps=r["stats"][:, 14].tolist(), posteq_steps=r["stats"][:, 15].tolist(),
            y=np.where(np.isnan(y), None, y).tolist(), sigmay=np.where(np.isnan(s), None, s).tolist())
        with open(os.path.join(outdir, name + ".json"), "w") as fh:
            json.dump(doc, fh, indent=1)
        print(name, "nllh", r["fval"])


if __name__ == "__main__":
    main()
