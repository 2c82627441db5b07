"""parPE's own example model ``model_steadystate_scaled`` (BASELINE config 0).

Restated from ``examples/parpeamici/steadystate/createSteadystateExampleSBML.py:
112-137`` (species x1..x3 with initial amounts 0.1/0.4/0.7, parameters p1..p5,
condition-specific constant k0, six reactions) and the observable table of
``examples/parpeamici/steadystate/create_steadystate_amici_model.py:444-462``:
six observables, one with a scaling and one with an offset parameter, and one
noise parameter per observable.  nx=3, np=13, nk=1, ny=6.
"""
import sympy as sp

from ..codegen.modelgen import ModelSpec


def spec():
    t = sp.Symbol("t", real=True)
    x = sp.symbols("x1 x2 x3", real=True)
    p = sp.symbols("p1 p2 p3 p4 p5", real=True)
    scal, offs = sp.symbols("observableParameter1_obs_x1_scaled observableParameter1_obs_x2_offsetted", real=True)
    noise = sp.symbols("noiseParameter1_obs_x1 noiseParameter1_obs_x2 noiseParameter1_obs_x3 "
                       "noiseParameter1_obs_x1_scaled noiseParameter1_obs_x2_offsetted "
                       "noiseParameter1_obs_x1withsigma", real=True)
    k0 = sp.Symbol("k0", real=True)
    x1, x2, x3 = x
    p1, p2, p3, p4, p5 = p
    xdot = [
        -2 * p1 * x1**2 - p2 * x1 * x2 + 2 * p3 * x2 + p4 * x3 + p5,
        p1 * x1**2 - p2 * x1 * x2 - p3 * x2 + p4 * x3,
        p2 * x1 * x2 - p4 * x3 - k0 * x3,
    ]
    return ModelSpec(
        name="steadystate_scaled", states=list(x), params=list(p) + [scal, offs] + list(noise),
        consts=[k0], t=t,
        x0=[sp.Float(0.1), sp.Float(0.4), sp.Float(0.7)], xdot=xdot,
        observables=[x1, x2, x3, scal * x1, offs + x2, x1],
        sigmay=list(noise),
        p_nominal=[1.0, 0.5, 0.4, 2.0, 0.1, 2.0, 3.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2],
        k_nominal=[1.0], doc=__doc__)
