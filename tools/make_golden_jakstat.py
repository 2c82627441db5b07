#!/usr/bin/env python
"""Fixture for the imported model with a spline input: the reference's own pre-generated model
deps/AMICI/models/model_jakstat_adjoint (w[0] = amici::spline_pos(t, 5, ...)) on the reference's own test inputs
(deps/AMICI/tests/cpp/testOptions.h5, group /model_jakstat_adjoint/sensiforward: theta, kappa, ts, Y, Sigma_Y -- read
in place with the repo's HDF5 reader), run by the reference build (oracle/_ref: the vendored Model_model_jakstat_adjoint
compiled in place).  The GPU box has neither /root/reference nor a live copy of that model: the committed file
tests/golden/imported/jakstat_sensiforward.json carries inputs and reference outputs there.

    python tools/make_golden_jakstat.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle.binding import OracleProblem  # noqa: E402
from parpe_b200 import capi  # noqa: E402
from tests.util import jakstat_tables  # noqa: E402

SRC = "/root/reference/deps/AMICI/tests/cpp/testOptions.h5"
GROUP = "/model_jakstat_adjoint/sensiforward"


def main():
    f = capi.H5File(SRC)
    rd = lambda p: np.asarray(f.read(GROUP + p), dtype=float)
    doc = dict(case="jakstat_sensiforward", model="amici_jakstat_adjoint", source=SRC + ":" + GROUP,
               theta=rd("/options/theta").tolist(), kappa=rd("/options/kappa").tolist(), ts=rd("/options/ts").tolist(),
               Y=[[None if np.isnan(v) else v for v in row] for row in rd("/data/Y").reshape(16, 3)],
               Sigma_Y=[[None if np.isnan(v) else v for v in row] for row in rd("/data/Sigma_Y").reshape(16, 3)],
               # three conditions: the reference's test case, other compartment volumes, and data sigmas in place of
               # the model's (p[14..16]) with one observable unmeasured
               conditions=[dict(kappa_factor=[1.0, 1.0]), dict(kappa_factor=[1.3, 0.8]),
                           dict(kappa_factor=[1.0, 1.0], data_sigma=0.15, drop_observable=1)])
    out = {}
    for ism in (2, 1):
        tab = jakstat_tables(doc)
        tab.ism = 1 if ism == 1 else 0
        theta = np.array(doc["theta"])
        r = OracleProblem(tab, 2, 9, model_name="jakstat_vendored").evaluate(theta, collect_stats=True)
        assert r["status"] == 0
        out["staggered" if ism == 2 else "simultaneous"] = dict(
            nllh=r["fval"], grad=list(map(float, r["grad"])), llh=list(map(float, r["llh"])),
            sllh=[list(map(float, s)) for s in r["sllh"]], nst=[int(v) for v in r["stats"][:, 0]])
    doc["reference"] = out
    path = os.path.join(ROOT, "tests", "golden", "imported", "jakstat_sensiforward.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1)
    print("wrote", path, "nllh", out["staggered"]["nllh"], out["simultaneous"]["nllh"])


if __name__ == "__main__":
    main()
