"""Runs AMICI's OWN python code generator (``amici.de_export.DEExporter``, ``deps/AMICI/python/sdist/amici``) on one
of the repo's symbolic model descriptions and leaves the model directory it writes -- the python-generated layout
(``xdot / w / dwdx / dwdp / dxdotdw / dxdotdx_explicit / dxdotdp_explicit / x0 / x0_fixedParameters / y / dydx /
sigmay / Jy ...``, macro headers ``x.h p.h k.h w.h ...``) -- for ``amici_import.py`` to re-home as device code and for
the oracle to compile as an ``amici::Model_ODE``.

The generator is the reference's code, used in place from ``/root/reference`` (build container only).  What this
image lacks is stubbed, none of it on the path that is exercised: ``libsbml`` (SBML import: the model is built
programmatically as an ``amici.de_model.DEModel``, the way ``amici.pysb_import`` builds one,
``pysb_import.py:363-420,534-745``), ``toposort`` (a 10-line topological sort), and the package's SWIG extension
(``amici/__init__.py`` skips it when imported by its own build).

``with_expressions``: every additive term of a right-hand side becomes a repeating element ("flux") ``w_k`` -- the
shape SBML reactions take -- so that ``w / dwdx / dwdp / dxdotdw`` and the chain rule through them are exercised.

The output is produced at BUILD time under ``parpe_b200/csrc/models/imported/pygen_src/`` (git-ignored).

    python -m parpe_b200.codegen.amici_pygen boehm [w]
"""
from __future__ import annotations

import os
import sys
import types

AMICI_PY = "/root/reference/deps/AMICI/python/sdist"
HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.path.join(os.path.dirname(HERE), "csrc", "models", "imported", "pygen_src")


def available() -> bool:
    return os.path.isdir(os.path.join(AMICI_PY, "amici"))


def _toposort(data):
    data = {k: set(v) - {k} for k, v in data.items()}
    extra = set().union(*data.values()) - set(data) if data else set()
    data.update({e: set() for e in extra})
    while True:
        ready = {k for k, v in data.items() if not v}
        if not ready:
            break
        yield ready
        data = {k: v - ready for k, v in data.items() if k not in ready}
    if data:
        raise ValueError("circular dependency")


class _AnyModule(types.ModuleType):
    """stand-in for libsbml: only names are looked up at import time (type annotations, isinstance checks)"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def _import_amici():
    if AMICI_PY not in sys.path:
        sys.path.insert(0, AMICI_PY)
    sys.modules.setdefault("libsbml", _AnyModule("libsbml"))
    ts = types.ModuleType("toposort")
    ts.toposort = _toposort
    ts.toposort_flatten = lambda d, sort=True: [x for s in _toposort(d) for x in (sorted(s) if sort else s)]
    ts.CircularDependencyError = ValueError
    sys.modules.setdefault("toposort", ts)
    # amici/__init__.py imports its SWIG extension unless it believes it is imported by its own build
    # (_imported_from_setup, __init__.py:53-75): only the pure-python code generator is needed here
    exec(compile("import amici", os.sep + os.path.join("x", "setuptools", "build_meta.py"), "exec"), {})
    import amici  # noqa: F401
    return amici


def generate(model: str, with_expressions: bool = False, root: str = SRC_ROOT) -> str:
    """-> the model directory AMICI's generator wrote (model name ``pygen_<model>[_w]``)"""
    import sympy as sp
    _import_amici()
    from amici.de_export import DEExporter
    from amici.de_model import DEModel
    from amici.de_model_components import (Constant, DifferentialState, Expression, LogLikelihoodY, Observable,
                                           Parameter, SigmaY)
    from amici.import_utils import (_get_str_symbol_identifiers, generate_measurement_symbol,
                                    noise_distribution_to_cost_function,
                                    noise_distribution_to_observable_transformation)
    from ..models import get_spec
    spec = get_spec(model)
    name = "pygen_" + model + ("_w" if with_expressions else "")
    ode = DEModel(verbose=False, simplify=None)
    for i, (s, x0, xd) in enumerate(zip(spec.states, spec.x0, spec.xdot)):
        xd = sp.sympify(xd)
        if with_expressions:
            syms = []
            for j, term in enumerate(sp.Add.make_args(sp.expand(xd))):
                fs = sp.Symbol("flux_%d_%d" % (i, j))
                ode.add_component(Expression(fs, str(fs), term))
                syms.append(fs)
            xd = sp.Add(*syms)
        ode.add_component(DifferentialState(s, str(s), sp.sympify(x0), xd))
    for p, v in zip(spec.params, spec.p_nominal):
        ode.add_component(Parameter(p, str(p), sp.Float(v)))
    for k, v in zip(spec.consts, spec.k_nominal):
        ode.add_component(Constant(k, str(k), sp.Float(v)))
    for iy, (obs, sig) in enumerate(zip(spec.observables, spec.sigmay)):
        nm = "obs%d" % iy
        y = sp.Symbol(nm)
        o = Observable(y, nm, sp.sympify(obs), transformation=noise_distribution_to_observable_transformation("normal"))
        ode.add_component(o)
        sigma = sp.Symbol("sigma_" + nm)
        ode.add_component(SigmaY(sigma, "sigma_" + nm, sp.sympify(sig)))
        cost = noise_distribution_to_cost_function("normal")(nm)
        my = generate_measurement_symbol(o.get_id())
        expr = sp.sympify(cost, locals=dict(zip(_get_str_symbol_identifiers(nm), (y, my, sigma), strict=True)))
        ode.add_component(LogLikelihoodY(sp.Symbol("llh_" + nm), "llh_" + nm, expr))
    ode._has_quadratic_nllh = True
    ode.generate_basic_variables()
    outdir = os.path.join(root, name)
    os.makedirs(root, exist_ok=True)
    DEExporter(ode, outdir=outdir, model_name=name, verbose=False,
               allow_reinit_fixpar_initcond=True).generate_model_code()
    return outdir


def write_registry(root: str = SRC_ROOT) -> str:
    """``pygen_registry.cpp``: name -> ``amici::Model`` factory over the model directories present, for the oracle
    (oracle/Makefile compiles it together with the generated sources)"""
    names = sorted(d for d in os.listdir(root) if os.path.exists(os.path.join(root, d, d + ".h")))
    L = ["// GENERATED by parpe_b200/codegen/amici_pygen.py -- TEST INFRASTRUCTURE (oracle): the models AMICI's python",
         "// code generator wrote, by name.", "#include <memory>", "#include <string>", '#include "amici/model.h"']
    for n in names:  # (the generator's header template leaves its include guard unsubstituted: the same for every model)
        L += ['#include "%s/%s.h"' % (n, n), "#undef _amici_TPL_MODELNAME_h"]
    L.append("std::unique_ptr<amici::Model> pb200_pygen_model(const std::string& name) {")
    for n in names:
        L.append('    if (name == "%s") return std::make_unique<amici::model_%s::Model_%s>();' % (n, n, n))
    L += ["    return nullptr;", "}"]
    path = os.path.join(root, "pygen_registry.cpp")
    text = "\n".join(L) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as fh:
            fh.write(text)
    return path


if __name__ == "__main__":
    print(generate(sys.argv[1], len(sys.argv) > 2))
    print(write_registry())
