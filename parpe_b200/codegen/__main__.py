"""``python -m parpe_b200.codegen`` regenerates device headers and oracle models."""
import os
import sys

from ..models import REGISTRY, get_spec
from ..models.netmodels import NET_MODELS
from . import netgen
from .modelgen import generate, count_flops

root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
names = sys.argv[1:] or list(REGISTRY)
for n in names:
    if n in NET_MODELS:
        s = get_spec(n)
        netgen.generate(s, root)
        print(n, "nx=%d nr=%d np=%d nk=%d ny=%d nnz=%d" % (s.nx, s.nr, s.np, s.nk, s.ny, s.nnz), s.flops())
        continue
    d = generate(get_spec(n), root)
    print(n, "nx=%d np=%d nk=%d ny=%d" % (d.spec.nx, d.spec.np, d.spec.nk, d.spec.ny), count_flops(d))
