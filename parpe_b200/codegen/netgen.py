"""Generator for large mass-action reaction networks (the "Chen-shape" / "Froehlich-shape"
workloads of SURVEY.md 8(d): nx = 500 / 1200 states, nnz(J) ~ 10 nx, rate constants
log-uniform over six decades).

What it replaces in the reference: AMICI's model export for one model
(``deps/AMICI/python/sdist/amici/de_export.py``, function catalogue
``_codegen/cxx_functions.py:50-340``).  For a model with hundreds of states the model
functions are not emitted as straight-line code -- a CTA evaluates them in parallel over
reactions / Jacobian entries -- but as *tables* (reaction list, stoichiometry in CSR form,
the symbolic Jacobian as one term list per structural non-zero, observables) plus one generic
evaluator (``parpe_b200/csrc/net_model.cuh``).  The same emitted header is consumed by the
CUDA library (uploaded to HBM at ``pb200_create``), by the host harness and by the oracle's
``amici::Model_ODE`` subclass (``oracle/net_model.hpp``), so all three evaluate the same model.

Reactions (rate v = kappa * p[pidx] * product of reactants):
    conversion A -> B, binding A + B -> C, dissociation C -> A + B, degradation A -> 0,
    synthesis 0 -> A.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List

import numpy as np


@dataclasses.dataclass
class NetworkSpec:
    name: str
    nx: int
    np: int
    nk: int
    ny: int
    r_a: np.ndarray       # first reactant or -1
    r_b: np.ndarray       # second reactant or -1
    kappa: np.ndarray     # fixed multiplier of the rate constant
    pidx: np.ndarray      # model parameter scaling the rate constant
    s_ptr: np.ndarray     # stoichiometry, CSR by species
    s_reac: np.ndarray
    s_coef: np.ndarray
    j_ptr: np.ndarray     # Jacobian pattern, CSR
    j_col: np.ndarray
    jt_ptr: np.ndarray    # terms of every Jacobian entry
    jt_reac: np.ndarray
    jt_other: np.ndarray  # the other reactant (-1: none)
    jt_coef: np.ndarray
    csc_ptr: np.ndarray   # the same pattern in CSC (AMICI's JSparse) ...
    csc_row: np.ndarray
    csc_from_csr: np.ndarray  # ... and where each CSC entry sits in the CSR value array
    jell_w: int               # the same pattern in ELL form (device: coalesced row-parallel J*s):
    jell_col: np.ndarray      # column of the t-th entry of row i at [t*nx + i], -1 = padding
    jell_pos: np.ndarray      # position [t*nx + i] of CSR entry e
    x0_base: np.ndarray
    x0_kidx: np.ndarray   # constant multiplying the initial value, or -1
    o_ptr: np.ndarray     # observables: y_j = sum w * x_i
    o_idx: np.ndarray
    o_w: np.ndarray
    sigma0: np.ndarray    # model sigma per observable (constant)
    p_nominal: List[float] = dataclasses.field(default_factory=list)
    k_nominal: List[float] = dataclasses.field(default_factory=list)
    doc: str = ""

    @property
    def nr(self):
        return len(self.r_a)

    @property
    def nnz(self):
        return len(self.j_col)

    # operation counts (F_f, F_J, F_dfdp)
    def flops(self):
        nreact = int(np.sum(self.r_a >= 0) + np.sum(self.r_b >= 0))
        return {
            "xdot": 2 * self.nr + nreact + 2 * len(self.s_reac),
            "J": 4 * len(self.jt_reac),
            "dxdotdp": 2 * len(self.s_reac),
            "nnzJ": self.nnz,
        }


def make_network(name: str, nx: int, nr: int, np_: int, ny: int, seed: int, doc: str = "") -> NetworkSpec:
    rng = np.random.default_rng(seed)
    ra, rb, prod = [], [], []  # prod: list of product species per reaction
    # Species carry a mass: the first n1 are monomers (1), the rest complexes (2).  Every reaction
    # conserves mass except synthesis (constant inflow of monomers) and degradation, so the total
    # stays bounded and the solution positive.
    n1 = (3 * nx) // 5
    mono = np.arange(n1)
    cplx = np.arange(n1, nx)
    # two rings of conversions keep the network connected
    for cls in (mono, cplx):
        for q in range(len(cls)):
            ra.append(int(cls[q])); rb.append(-1); prod.append([int(cls[(q + 1) % len(cls)])])
    n_syn = max(2, nx // 25)
    for i in rng.choice(mono, n_syn, replace=False):
        ra.append(-1); rb.append(-1); prod.append([int(i)])          # synthesis
    for i in range(nx):
        if i % 2 == 0:
            ra.append(i); rb.append(-1); prod.append([])             # degradation
    while len(ra) < nr:
        kind = rng.integers(0, 3)
        if kind == 0:
            cls = mono if rng.random() < 0.6 else cplx
            a, b = (int(v) for v in rng.choice(cls, 2, replace=False))
            ra.append(a); rb.append(-1); prod.append([b])            # conversion within a mass class
        else:
            a, b = (int(v) for v in rng.choice(mono, 2, replace=False))
            c = int(rng.choice(cplx))
            ra.append(a); rb.append(b); prod.append([c])             # binding ...
            ra.append(c); rb.append(-1); prod.append([a, b])         # ... and its dissociation
    nr = len(ra)
    # Species are numbered by decreasing number of Jacobian entries in their row, so that the 32 rows
    # a warp works on have (nearly) the same length: the row-parallel J*s loops run to the longest
    # row of the warp.
    dep = [set([i]) for i in range(nx)]
    for r in range(nr):
        touched = [q for q in (ra[r], rb[r]) if q >= 0] + list(prod[r])
        for sp_ in touched:
            for q in (ra[r], rb[r]):
                if q >= 0:
                    dep[sp_].add(q)
    order = sorted(range(nx), key=lambda i: (-len(dep[i]), i))
    newidx = np.empty(nx, np.int64)
    newidx[order] = np.arange(nx)
    ra = [int(newidx[q]) if q >= 0 else -1 for q in ra]
    rb = [int(newidx[q]) if q >= 0 else -1 for q in rb]
    prod = [[int(newidx[q]) for q in pr_] for pr_ in prod]
    ra = np.asarray(ra, np.int32)
    rb = np.asarray(rb, np.int32)
    kappa = 10.0 ** rng.uniform(-4.0, 2.0, nr)
    kappa[(ra < 0)] = 10.0 ** rng.uniform(-2.0, 0.0, int(np.sum(ra < 0)))   # synthesis: moderate
    pidx = rng.permutation(nr).astype(np.int32) % np_

    # stoichiometry rows
    rows = [dict() for _ in range(nx)]
    for r in range(nr):
        for sp_, co in [(ra[r], -1.0), (rb[r], -1.0)] + [(q, 1.0) for q in prod[r]]:
            if sp_ >= 0:
                rows[sp_][r] = rows[sp_].get(r, 0.0) + co
    s_ptr, s_reac, s_coef = [0], [], []
    for i in range(nx):
        for r in sorted(rows[i]):
            if rows[i][r] != 0.0:
                s_reac.append(r); s_coef.append(rows[i][r])
        s_ptr.append(len(s_reac))

    # Jacobian: d xdot_i / d x_j = sum_r S_ir * kappa_r p_r * d(prod reactants)/d x_j
    jrows = [dict() for _ in range(nx)]
    for i in range(nx):
        for e in range(s_ptr[i], s_ptr[i + 1]):
            r, co = s_reac[e], s_coef[e]
            a, b = int(ra[r]), int(rb[r])
            if a >= 0:
                jrows[i].setdefault(a, []).append((r, b, co))
            if b >= 0:
                jrows[i].setdefault(b, []).append((r, a, co))
    j_ptr, j_col, jt_ptr, jt_reac, jt_other, jt_coef = [0], [], [0], [], [], []
    for i in range(nx):
        jrows[i].setdefault(i, [])  # keep the diagonal structurally present
        for j in sorted(jrows[i]):
            j_col.append(j)
            for (r, other, co) in jrows[i][j]:
                jt_reac.append(r); jt_other.append(other); jt_coef.append(co)
            jt_ptr.append(len(jt_reac))
        j_ptr.append(len(j_col))
    # CSC view
    nnz = len(j_col)
    row_of = np.repeat(np.arange(nx), np.diff(j_ptr))
    order = np.lexsort((row_of, np.asarray(j_col)))  # by column, then row
    csc_row = row_of[order]
    csc_from_csr = order
    csc_ptr = np.concatenate([[0], np.cumsum(np.bincount(np.asarray(j_col), minlength=nx))])

    jell_w = int(np.max(np.diff(j_ptr)))
    jell_col = np.full(jell_w * nx, -1, np.int32)
    jell_pos = np.zeros(nnz, np.int32)
    for i in range(nx):
        for t, e in enumerate(range(j_ptr[i], j_ptr[i + 1])):
            jell_col[t * nx + i] = j_col[e]
            jell_pos[e] = t * nx + i

    x0_base = rng.uniform(0.1, 2.0, nx)
    x0_kidx = np.full(nx, -1, np.int32)
    x0_kidx[0::10] = 0
    x0_kidx[1::10] = 1
    o_ptr, o_idx, o_w = [0], [], []
    for j in range(ny):
        for i in sorted(rng.choice(nx, 5, replace=False)):
            o_idx.append(int(i)); o_w.append(float(rng.integers(1, 3)))
        o_ptr.append(len(o_idx))
    I32 = lambda a: np.asarray(a, np.int32)
    F64 = lambda a: np.asarray(a, np.float64)
    return NetworkSpec(
        name=name, nx=nx, np=np_, nk=2, ny=ny, r_a=ra, r_b=rb, kappa=kappa, pidx=pidx,
        s_ptr=I32(s_ptr), s_reac=I32(s_reac), s_coef=F64(s_coef),
        j_ptr=I32(j_ptr), j_col=I32(j_col), jt_ptr=I32(jt_ptr), jt_reac=I32(jt_reac), jt_other=I32(jt_other),
        jt_coef=F64(jt_coef), csc_ptr=I32(csc_ptr), csc_row=I32(csc_row), csc_from_csr=I32(csc_from_csr),
        jell_w=jell_w, jell_col=jell_col, jell_pos=jell_pos,
        x0_base=x0_base, x0_kidx=x0_kidx, o_ptr=I32(o_ptr), o_idx=I32(o_idx), o_w=F64(o_w),
        sigma0=np.full(ny, 0.1), p_nominal=[1.0] * np_, k_nominal=[1.0, 1.0], doc=doc)


def _arr(name, ctype, a):
    if ctype == "double":
        body = ", ".join(repr(float(v)) for v in a)
    else:
        body = ", ".join(str(int(v)) for v in a)
    if len(a) == 0:
        body = "0"
    return "static const %s %s[] = {%s};" % (ctype, name, body)


def emit_tables_header(s: NetworkSpec, path: str) -> None:
    fl = s.flops()
    L = []
    A = L.append
    A("// GENERATED by parpe_b200/codegen/netgen.py (model '%s') -- do not edit." % s.name)
    A("// Mass-action network as tables for the generic evaluator net_model.cuh.")
    A("#ifndef PB200_NET_%s_H" % s.name.upper())
    A("#define PB200_NET_%s_H" % s.name.upper())
    A('#include "../net_model.cuh"')
    A("namespace pb200 { namespace nets { namespace %s {" % s.name)
    for nm, ct in [("r_a", "int"), ("r_b", "int"), ("kappa", "double"), ("pidx", "int"), ("s_ptr", "int"),
                   ("s_reac", "int"), ("s_coef", "double"), ("j_ptr", "int"), ("j_col", "int"), ("jt_ptr", "int"),
                   ("jt_reac", "int"), ("jt_other", "int"), ("jt_coef", "double"), ("csc_ptr", "int"),
                   ("csc_row", "int"), ("csc_from_csr", "int"), ("jell_col", "int"), ("jell_pos", "int"), ("x0_base", "double"), ("x0_kidx", "int"),
                   ("o_ptr", "int"), ("o_idx", "int"), ("o_w", "double"), ("sigma0", "double")]:
        A(_arr(nm, ct, getattr(s, nm)))
    A("inline NetTables tables() {")
    A("    NetTables t;")
    A('    t.name = "%s";' % s.name)
    A("    t.nx = %d; t.nr = %d; t.np = %d; t.nk = %d; t.ny = %d; t.nnz = %d;" % (s.nx, s.nr, s.np, s.nk, s.ny, s.nnz))
    A("    t.n_s = %d; t.n_jt = %d; t.n_o = %d; t.jell_w = %d;" % (len(s.s_reac), len(s.jt_reac), len(s.o_idx), s.jell_w))
    A("    t.flops_xdot = %d; t.flops_J = %d; t.flops_dxdotdp = %d;" % (fl["xdot"], fl["J"], fl["dxdotdp"]))
    for nm in ["r_a", "r_b", "kappa", "pidx", "s_ptr", "s_reac", "s_coef", "j_ptr", "j_col", "jt_ptr", "jt_reac",
               "jt_other", "jt_coef", "csc_ptr", "csc_row", "csc_from_csr", "jell_col", "jell_pos", "x0_base", "x0_kidx", "o_ptr", "o_idx",
               "o_w", "sigma0"]:
        A("    t.%s = %s;" % (nm, nm))
    A("    return t;")
    A("}")
    A("}}} // namespace pb200::nets::%s" % s.name)
    A("#endif")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("\n".join(L) + "\n")


def generate(spec: NetworkSpec, repo_root: str) -> None:
    mdir = os.path.join(repo_root, "parpe_b200", "csrc", "models")
    emit_tables_header(spec, os.path.join(mdir, "net_%s.h" % spec.name))
    with open(os.path.join(mdir, "inst_net_%s.cu" % spec.name), "w") as fh:
        fh.write("// GENERATED by parpe_b200/codegen/netgen.py -- team-kernel registration for network model %s.\n"
                 "#include \"../team_kernels.cuh\"\n#include \"net_%s.h\"\nPB200_REGISTER_NET(%s)\n"
                 % (spec.name, spec.name, spec.name))
