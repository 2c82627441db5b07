"""Symbolic model descriptions consumed by ``parpe_b200.codegen.modelgen``.

Each module exposes ``spec() -> ModelSpec``.  ``REGISTRY`` lists the models for
which device code and oracle code are generated and compiled into the
libraries; the integer id is the ``model_id`` of the C ABI
(``include/parpe_b200.h``).
"""
import importlib

REGISTRY = {
    "steadystate": 0,
    "steadystate_scaled": 1,
    "boehm": 2,
    "fujita": 3,
    # table-driven mass-action networks (large-model path, one CTA per condition)
    "net32": 4,
    "chen": 5,
    "froehlich": 6,
    "net31": 7,
}


def get_spec(name):
    from .netmodels import NET_MODELS, spec as net_spec
    if name in NET_MODELS:
        return net_spec(name)
    return importlib.import_module("parpe_b200.models." + name).spec()
