"""``python -m parpe_b200.codegen`` regenerates device headers and oracle models."""
import os
import sys

from ..models import REGISTRY, get_spec
from .modelgen import generate, count_flops

root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
names = sys.argv[1:] or list(REGISTRY)
for n in names:
    d = generate(get_spec(n), root)
    print(n, "nx=%d np=%d nk=%d ny=%d" % (d.spec.nx, d.spec.np, d.spec.nk, d.spec.ny), count_flops(d))
