"""Large mass-action network models (table-driven, ``parpe_b200.codegen.netgen``).

``chen`` and ``froehlich`` are *shape* stand-ins for BASELINE configs 3 and 4
(Chen_MSB2009: ~500 states, stiff; Froehlich_CellSystems2018: ~1200 states): the
real PEtab problems are not vendored in the reference tree and cannot be imported
offline, so the networks are generated (fixed seeds) with the dimensions and sparsity
SURVEY.md 8(d) names: nnz(J) ~ 10 nx, rate constants log-uniform over six decades,
nplist = 50 / 100.  ``net32`` and ``net31`` are small members of the same family for fast tests.
"""
import functools

from ..codegen.netgen import make_network

NET_MODELS = {
    # name: (nx, nr, np, ny, seed)
    "net32": (32, 112, 6, 4, 11),
    "net31": (31, 100, 5, 3, 14),   # odd nx, one short block: the fallback paths of the rows kernel
    "chen": (500, 2000, 50, 8, 12),
    "froehlich": (1200, 4800, 100, 8, 13),
}


@functools.lru_cache(maxsize=None)
def spec(name):
    nx, nr, np_, ny, seed = NET_MODELS[name]
    return make_network(name, nx, nr, np_, ny, seed, doc=__doc__)
