"""3-state conversion-reaction network of AMICI's vendored test model.

Restated from ``deps/AMICI/models/model_steadystate/{xdot,w,x0,y,sigmay}.cpp``
(nx=3, np=5, nk=4, ny=3): the same chemistry as parPE's steadystate example
(``examples/parpeamici/steadystate/createSteadystateExampleSBML.py:120-137``),
with the initial state supplied through the constants k[0..2] and the decay
rate of x3 through k[3].
"""
import sympy as sp

from ..codegen.modelgen import ModelSpec


def spec():
    t = sp.Symbol("t", real=True)
    x = sp.symbols("x1 x2 x3", real=True)
    p = sp.symbols("p1 p2 p3 p4 p5", real=True)
    k = sp.symbols("k1 k2 k3 k4", real=True)
    x1, x2, x3 = x
    p1, p2, p3, p4, p5 = p
    xdot = [
        p5 + p4 * x3 - 2 * p1 * x1**2 + 2 * p3 * x2 - p2 * x1 * x2,
        p4 * x3 + p1 * x1**2 - p3 * x2 - p2 * x1 * x2,
        -p4 * x3 - k[3] * x3 + p2 * x1 * x2,
    ]
    return ModelSpec(
        name="steadystate", states=list(x), params=list(p), consts=list(k), t=t,
        x0=[k[0], k[1], k[2]], xdot=xdot, observables=[x1, x2, x3],
        sigmay=[sp.Integer(1)] * 3,
        p_nominal=[1.0, 0.5, 0.4, 2.0, 0.1], k_nominal=[0.1, 0.4, 0.7, 1.0],
        doc=__doc__)
