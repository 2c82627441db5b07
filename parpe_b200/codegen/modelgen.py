"""Model code generator: one symbolic ODE model -> (i) a C++ ``amici::Model_ODE``
subclass for the CPU oracle and (ii) the same expressions as ``__host__
__device__`` functions for the CUDA integrator.

What it replaces in the reference: AMICI's model export (function catalogue
``deps/AMICI/python/sdist/amici/_codegen/cxx_functions.py:50-340``; a generated
example is ``deps/AMICI/models/model_steadystate/*.cpp``, class layout
``deps/AMICI/models/model_steadystate/model_steadystate.h``).  The C++ emitted
for the oracle uses the MATLAB-generated layout (``pythonGenerated=false``,
simple path of ``deps/AMICI/src/model_ode.cpp:59-66,175-187``): ``xdot``,
``JSparse`` (CSC), ``dxdotdp(ip)``, ``x0``, ``sx0(ip)``, ``y``, ``dydx``,
``dydp(ip)``, ``sigmay``, ``dsigmaydp(ip)``, Gaussian ``Jy/dJydy/dJydsigma``
(same formulas as ``deps/AMICI/models/model_steadystate/Jy.cpp:16-26``).
``w`` is not used (``nw = 0``); common sub-expressions are local temporaries.

Both back-ends are printed from the *same* sympy expressions after the same
CSE pass, so the CPU oracle and the GPU kernel evaluate identical formulas.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Sequence

import sympy as sp
from sympy.printing.c import C99CodePrinter


# --------------------------------------------------------------------------
# model description
# --------------------------------------------------------------------------
@dataclasses.dataclass
class ModelSpec:
    name: str
    states: List[sp.Symbol]
    params: List[sp.Symbol]
    consts: List[sp.Symbol]
    t: sp.Symbol
    x0: List[sp.Expr]
    xdot: List[sp.Expr]
    observables: List[sp.Expr]
    sigmay: List[sp.Expr]
    # nominal values, used by tests / benches to draw "random parameters in bounds"
    p_nominal: List[float] = dataclasses.field(default_factory=list)
    k_nominal: List[float] = dataclasses.field(default_factory=list)
    doc: str = ""

    @property
    def nx(self):
        return len(self.states)

    @property
    def np(self):
        return len(self.params)

    @property
    def nk(self):
        return len(self.consts)

    @property
    def ny(self):
        return len(self.observables)


class _Printer(C99CodePrinter):
    """C printer: x/p/k symbols become array accesses; small integer powers
    become products (no pow() calls in the hot path)."""

    def __init__(self, symmap: Dict[sp.Symbol, str]):
        super().__init__({"contract": False})
        self._symmap = symmap

    def _print_Symbol(self, expr):
        return self._symmap.get(expr, super()._print_Symbol(expr))

    def _print_Pow(self, expr):
        b, e = expr.as_base_exp()
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self.parenthesize(b, 1000)
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self.parenthesize(b, 1000)
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        return super()._print_Pow(expr)

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (expr.p, expr.q)

    def _print_Integer(self, expr):
        return "%d.0" % int(expr) if int(expr) >= 0 else "(%d.0)" % int(expr)


def _symmap(spec: ModelSpec) -> Dict[sp.Symbol, str]:
    m = {spec.t: "t"}
    for i, s in enumerate(spec.states):
        m[s] = "x[%d]" % i
    for i, s in enumerate(spec.params):
        m[s] = "p[%d]" % i
    for i, s in enumerate(spec.consts):
        m[s] = "k[%d]" % i
    return m


# --------------------------------------------------------------------------
# derivation
# --------------------------------------------------------------------------
@dataclasses.dataclass
class Derived:
    spec: ModelSpec
    J: sp.Matrix  # nx x nx  d xdot / d x
    dxdotdp: sp.Matrix  # nx x np
    sx0: sp.Matrix  # nx x np
    dydx: sp.Matrix  # ny x nx
    dydp: sp.Matrix  # ny x np
    dsigmaydp: sp.Matrix  # ny x np
    # CSC structure of J
    colptrs: List[int]
    rowvals: List[int]


def derive(spec: ModelSpec) -> Derived:
    X = sp.Matrix(spec.states)
    P = sp.Matrix(spec.params)
    f = sp.Matrix(spec.xdot)
    J = f.jacobian(X) if spec.nx else sp.zeros(0, 0)
    dfdp = f.jacobian(P) if spec.np else sp.zeros(spec.nx, 0)
    x0 = sp.Matrix(spec.x0)
    sx0 = x0.jacobian(P) if spec.np else sp.zeros(spec.nx, 0)
    y = sp.Matrix(spec.observables)
    dydx = y.jacobian(X)
    dydp = y.jacobian(P) if spec.np else sp.zeros(spec.ny, 0)
    sig = sp.Matrix(spec.sigmay)
    dsig = sig.jacobian(P) if spec.np else sp.zeros(spec.ny, 0)
    colptrs, rowvals = [0], []
    for c in range(spec.nx):
        for r in range(spec.nx):
            if J[r, c] != 0:
                rowvals.append(r)
        colptrs.append(len(rowvals))
    return Derived(spec, J, dfdp, sx0, dydx, dydp, dsig, colptrs, rowvals)


def _cse_block(pr: _Printer, targets: Sequence[str], exprs: Sequence[sp.Expr],
               indent: str = "    ", tmp_prefix: str = "c") -> List[str]:
    """Emit ``const double cK = ...;`` temporaries and ``target = expr;`` lines."""
    pairs = [(t, e) for t, e in zip(targets, exprs)]
    if not pairs:
        return []
    syms = sp.numbered_symbols(tmp_prefix)
    repl, red = sp.cse([e for _, e in pairs], symbols=syms, optimizations="basic")
    lines = []
    for s, e in repl:
        lines.append("%sconst double %s = %s;" % (indent, s, pr.doprint(e)))
    for (t, _), e in zip(pairs, red):
        lines.append("%s%s = %s;" % (indent, t, pr.doprint(e)))
    return lines


def _nonzero_assign(pr, fmt, exprs, indent="    ", tmp_prefix="c"):
    """Only structurally non-zero entries are assigned (callers zero the buffer),
    as the reference's generated code does."""
    tg, ex = [], []
    for idx, e in exprs:
        if e != 0:
            tg.append(fmt % idx)
            ex.append(e)
    return _cse_block(pr, tg, ex, indent, tmp_prefix)


def _switch_ip(pr, n_rows, mat, out, indent="    "):
    """switch (ip) { case c: out[r] = mat[r,c]; ... }"""
    lines = ["%sswitch (ip) {" % indent]
    for c in range(mat.shape[1]):
        body = _nonzero_assign(pr, out + "[%d]", [(r, mat[r, c]) for r in range(n_rows)],
                               indent + "    ")
        if body:
            lines.append("%scase %d: {" % (indent, c))
            lines += body
            lines.append("%s} break;" % indent)
    lines.append("%sdefault: break;" % indent)
    lines.append("%s}" % indent)
    return lines


def count_flops(d: Derived) -> Dict[str, int]:
    """Operation counts of the generated expressions after CSE (the F_f, F_J,
    F_dfdp figures of SURVEY.md section 8(d))."""
    def ops(exprs):
        exprs = [e for e in exprs if e != 0]
        if not exprs:
            return 0
        repl, red = sp.cse(exprs, optimizations="basic")
        return int(sum(sp.count_ops(e) for _, e in repl) + sum(sp.count_ops(e) for e in red))
    s = d.spec
    return {
        "xdot": ops(list(s.xdot)),
        "J": ops(list(d.J)),
        "dxdotdp": ops(list(d.dxdotdp)),
        "y": ops(list(s.observables)),
        "nnzJ": len(d.rowvals),
        "nnz_dxdotdp": int(sum(1 for e in d.dxdotdp if e != 0)),
    }


# --------------------------------------------------------------------------
# back-end 1: CUDA / host-device header
# --------------------------------------------------------------------------
def emit_device_header(d: Derived, path: str) -> None:
    s = d.spec
    pr = _Printer(_symmap(s))
    fl = count_flops(d)
    L: List[str] = []
    A = L.append
    guard = "PB200_MODEL_%s_CUH" % s.name.upper()
    A("// GENERATED by parpe_b200/codegen/modelgen.py from parpe_b200/models/%s.py -- do not edit." % s.name)
    A("// Model functions as __host__ __device__ code (what AMICI emits as C++ per model,")
    A("// reference catalogue deps/AMICI/python/sdist/amici/_codegen/cxx_functions.py:50-340).")
    A("#ifndef %s" % guard)
    A("#define %s" % guard)
    A('#include "../pb_hd.h"')
    A("namespace pb200 { namespace models {")
    A("struct Model_%s {" % s.name)
    A('    static constexpr const char* NAME = "%s";' % s.name)
    A("    static constexpr int NX = %d, NP = %d, NK = %d, NY = %d, NNZ = %d;" %
      (s.nx, s.np, max(s.nk, 0), s.ny, len(d.rowvals)))
    dyn = [any(d.dxdotdp[r, c] != 0 for r in range(s.nx)) or any(d.sx0[r, c] != 0 for r in range(s.nx))
           for c in range(s.np)]
    A("    // parameters the state depends on (a non-zero column of dxdotdp or sx0): only their sensitivity")
    A("    // columns are integrated; for the others sx == 0 identically (bit DYN_MASK >> ip)")
    A("    static constexpr int NPD = %d;" % max(sum(dyn), 0))
    A("    static constexpr unsigned DYN_MASK = 0x%xu;" % sum(1 << i for i, v in enumerate(dyn) if v))
    x0k = [any(sym in set(s.consts) for sym in sp.sympify(e).free_symbols) for e in s.x0]
    A("    // states whose initial value depends on the constants k (AMICI's _x0_fixedParameters_idx,")
    A("    // de_export.py:688-749): the ones a re-initialisation after pre-equilibration may touch (bit i)")
    A("    static constexpr unsigned X0K_MASK = 0x%xu;" % sum(1 << i for i, v in enumerate(x0k) if v))
    A("    // operation counts of the generated expressions (after CSE)")
    A("    static constexpr int FLOPS_XDOT = %d, FLOPS_J = %d, FLOPS_DXDOTDP = %d, NNZ_DXDOTDP = %d;" %
      (fl["xdot"], fl["J"], fl["dxdotdp"], fl["nnz_dxdotdp"]))
    sig = "const double t, const double* __restrict__ x, const double* __restrict__ p, const double* __restrict__ k"

    A("    PB_HD static void x0(double* __restrict__ x0, const double t, const double* __restrict__ p, const double* __restrict__ k) {")
    A("        (void)t; (void)p; (void)k;")
    L.extend(_cse_block(pr, ["x0[%d]" % i for i in range(s.nx)], list(s.x0), "        "))
    A("    }")
    A("    // d x0 / d p[ip]; caller zero-fills sx0")
    A("    PB_HD static void sx0(double* __restrict__ sx0, const double t, const double* __restrict__ p, const double* __restrict__ k, const int ip) {")
    A("        (void)t; (void)p; (void)k; (void)sx0;")
    L.extend(_switch_ip(pr, s.nx, d.sx0, "sx0", "        "))
    A("    }")
    A("    PB_HD static void xdot(double* __restrict__ xdot, %s) {" % sig)
    A("        (void)t; (void)p; (void)k;")
    L.extend(_cse_block(pr, ["xdot[%d]" % i for i in range(s.nx)], list(s.xdot), "        "))
    A("    }")
    A("    // dense column-major Jacobian J[r + c*NX]; caller zero-fills J")
    A("    PB_HD static void jac(double* __restrict__ J, %s) {" % sig)
    A("        (void)t; (void)p; (void)k;")
    L.extend(_nonzero_assign(pr, "J[%d]", [(r + c * s.nx, d.J[r, c]) for c in range(s.nx) for r in range(s.nx)], "        "))
    A("    }")
    A("    // column ip of d xdot / d p; caller zero-fills out")
    A("    PB_HD static void dxdotdp(double* __restrict__ out, %s, const int ip) {" % sig)
    A("        (void)t; (void)p; (void)k; (void)x; (void)out;")
    L.extend(_switch_ip(pr, s.nx, d.dxdotdp, "out", "        "))
    A("    }")
    A("    // all non-zero columns of d xdot / d p at once: out[r + ipd*NX] with ipd the rank of ip among the")
    A("    // dynamic parameters (popcount of DYN_MASK below bit ip); structural non-zeros only (caller zero-fills once)")
    A("    PB_HD static void dxdotdp_dense(double* __restrict__ out, %s) {" % sig)
    A("        (void)t; (void)p; (void)k; (void)x; (void)out;")
    rank = {}
    for c in range(s.np):
        if dyn[c]:
            rank[c] = len(rank)
    L.extend(_nonzero_assign(pr, "out[%d]", [(r + rank[c] * s.nx, d.dxdotdp[r, c]) for c in range(s.np) if dyn[c]
                                             for r in range(s.nx)], "        "))
    A("    }")
    A("    PB_HD static void y(double* __restrict__ y, %s) {" % sig)
    A("        (void)t; (void)p; (void)k; (void)x;")
    L.extend(_cse_block(pr, ["y[%d]" % i for i in range(s.ny)], list(s.observables), "        "))
    A("    }")
    A("    // dydx[iy + ix*NY] column-major; caller zero-fills")
    A("    PB_HD static void dydx(double* __restrict__ dydx, %s) {" % sig)
    A("        (void)t; (void)p; (void)k; (void)x;")
    L.extend(_nonzero_assign(pr, "dydx[%d]", [(r + c * s.ny, d.dydx[r, c]) for c in range(s.nx) for r in range(s.ny)], "        "))
    A("    }")
    A("    PB_HD static void dydp(double* __restrict__ out, %s, const int ip) {" % sig)
    A("        (void)t; (void)p; (void)k; (void)x; (void)out;")
    L.extend(_switch_ip(pr, s.ny, d.dydp, "out", "        "))
    A("    }")
    A("    PB_HD static void sigmay(double* __restrict__ sigmay, const double t, const double* __restrict__ p, const double* __restrict__ k) {")
    A("        (void)t; (void)p; (void)k;")
    L.extend(_cse_block(pr, ["sigmay[%d]" % i for i in range(s.ny)], list(s.sigmay), "        "))
    A("    }")
    A("    PB_HD static void dsigmaydp(double* __restrict__ out, const double t, const double* __restrict__ p, const double* __restrict__ k, const int ip) {")
    A("        (void)t; (void)p; (void)k; (void)out;")
    L.extend(_switch_ip(pr, s.ny, d.dsigmaydp, "out", "        "))
    A("    }")
    A("};")
    A("}} // namespace pb200::models")
    A("#endif")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("\n".join(L) + "\n")


# --------------------------------------------------------------------------
# back-end 2: amici::Model_ODE subclass (oracle side)
# --------------------------------------------------------------------------
def emit_amici_model(d: Derived, outdir: str) -> None:
    s = d.spec
    pr = _Printer(_symmap(s))
    n = s.name
    os.makedirs(outdir, exist_ok=True)
    ubw = max(s.nx - 1, 0)  # bandwidths only matter for the band solver (unused)
    L: List[str] = []
    A = L.append
    A("// GENERATED by parpe_b200/codegen/modelgen.py from parpe_b200/models/%s.py -- do not edit." % n)
    A("// TEST INFRASTRUCTURE (oracle): amici::Model_ODE subclass in the MATLAB-generated layout")
    A("// (pythonGenerated=false; cf. deps/AMICI/models/model_steadystate/model_steadystate.h).")
    A("#include <algorithm>")
    A("#include <memory>")
    A("#include <cmath>")
    A('#include "amici/model_ode.h"')
    A('#include "amici/solver_cvodes.h"')
    A("#include <sunmatrix/sunmatrix_sparse.h>")
    A("namespace pb200_oracle_models {")
    A("using amici::realtype;")
    A("class Model_%s : public amici::Model_ODE {" % n)
    A("  public:")
    A("    Model_%s()" % n)
    A("        : amici::Model_ODE(")
    A("              amici::ModelDimensions(%d, %d, %d, %d, 0, %d, %d, %d, %d, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, {}, 0, 0, 0, %d, %d, %d)," %
      (s.nx, s.nx, s.nx, s.nx, s.np, s.nk, s.ny, s.ny, len(d.rowvals), ubw, ubw))
    A("              amici::SimulationParameters(std::vector<realtype>(%d, 1.0), std::vector<realtype>(%d, 1.0))," % (s.nk, s.np))
    A("              amici::SecondOrderMode::none, std::vector<realtype>(%d, 0.0), std::vector<int>{}) {}" % s.nx)
    A("    amici::Model* clone() const override { return new Model_%s(*this); }" % n)
    A('    std::string getAmiciCommit() const override { return "pb200-modelgen"; }')
    A('    std::string getName() const override { return "%s"; }' % n)

    A("    void fxdot(realtype* xdot, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const realtype*) override {")
    L.extend(_cse_block(pr, ["xdot[%d]" % i for i in range(s.nx)], list(s.xdot), "        "))
    A("    }")
    A("    void fJSparse(SUNMatrixContent_Sparse JSparse, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const realtype*, const realtype*) override {")
    for i, r in enumerate(d.rowvals):
        A("        JSparse->indexvals[%d] = %d;" % (i, r))
    for i, c in enumerate(d.colptrs):
        A("        JSparse->indexptrs[%d] = %d;" % (i, c))
    ex, tg = [], []
    for c in range(s.nx):
        for kk in range(d.colptrs[c], d.colptrs[c + 1]):
            tg.append("JSparse->data[%d]" % kk)
            ex.append(d.J[d.rowvals[kk], c])
    L.extend(_cse_block(pr, tg, ex, "        "))
    A("    }")
    A("    void fdxdotdp(realtype* dxdotdp, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const int ip, const realtype*, const realtype*) override {")
    L.extend(_switch_ip(pr, s.nx, d.dxdotdp, "dxdotdp", "        "))
    A("    }")
    A("    void fx0(realtype* x0, const realtype t, const realtype* p, const realtype* k) override {")
    L.extend(_cse_block(pr, ["x0[%d]" % i for i in range(s.nx)], list(s.x0), "        "))
    A("    }")
    A("    void fsx0(realtype* sx0, const realtype t, const realtype* x0, const realtype* p, const realtype* k, const int ip) override {")
    L.extend(_switch_ip(pr, s.nx, d.sx0, "sx0", "        "))
    A("    }")
    # re-initialisation of states whose initial value depends on the constants, in the layout of the
    # python-generated models (deps/AMICI/python/sdist/amici/de_export.py:688-749)
    x0k = [i for i, e in enumerate(s.x0) if any(sym in set(s.consts) for sym in sp.sympify(e).free_symbols)]
    A("    static bool reinit_has(gsl::span<const int> idxs, int i) { return std::find(idxs.begin(), idxs.end(), i) != idxs.end(); }")
    A("    void fx0_fixedParameters(realtype* x0_fixedParameters, const realtype t, const realtype* p, const realtype* k, gsl::span<const int> reinitialization_state_idxs) override {")
    for i in x0k:
        A("        if (reinit_has(reinitialization_state_idxs, %d)) x0_fixedParameters[%d] = %s;" % (i, i, pr.doprint(s.x0[i])))
    A("    }")
    A("    void fsx0_fixedParameters(realtype* sx0_fixedParameters, const realtype t, const realtype* x0, const realtype* p, const realtype* k, const int ip, gsl::span<const int> reinitialization_state_idxs) override {")
    for i in x0k:
        A("        if (reinit_has(reinitialization_state_idxs, %d)) sx0_fixedParameters[%d] = 0.0;" % (i, i))
    A("        switch (ip) {")
    for c in range(s.np):
        body = ["            if (reinit_has(reinitialization_state_idxs, %d)) sx0_fixedParameters[%d] = %s;" % (i, i, pr.doprint(d.sx0[i, c]))
                for i in x0k if d.sx0[i, c] != 0]
        if body:
            A("        case %d:" % c)
            L.extend(body)
            A("            break;")
    A("        default: break;")
    A("        }")
    A("    }")
    A("    void fy(double* y, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const realtype*) override {")
    L.extend(_cse_block(pr, ["y[%d]" % i for i in range(s.ny)], list(s.observables), "        "))
    A("    }")
    A("    void fdydx(double* dydx, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const realtype*, const realtype*) override {")
    L.extend(_nonzero_assign(pr, "dydx[%d]", [(r + c * s.ny, d.dydx[r, c]) for c in range(s.nx) for r in range(s.ny)], "        "))
    A("    }")
    A("    void fdydp(double* dydp, const realtype t, const realtype* x, const realtype* p, const realtype* k, const realtype*, const int ip, const realtype*, const realtype*) override {")
    L.extend(_switch_ip(pr, s.ny, d.dydp, "dydp", "        "))
    A("    }")
    A("    void fsigmay(double* sigmay, const realtype t, const realtype* p, const realtype* k, const realtype*) override {")
    L.extend(_cse_block(pr, ["sigmay[%d]" % i for i in range(s.ny)], list(s.sigmay), "        "))
    A("    }")
    A("    void fdsigmaydp(double* dsigmaydp, const realtype t, const realtype* p, const realtype* k, const realtype*, const int ip) override {")
    L.extend(_switch_ip(pr, s.ny, d.dsigmaydp, "dsigmaydp", "        "))
    A("    }")
    # Gaussian objective, same formulas as the reference's generated Jy/dJydy/dJydsigma
    A("    void fJy(double* nllh, const int iy, const realtype*, const realtype*, const double* y, const double* sigmay, const double* my) override {")
    A("        nllh[0] = std::log((sigmay[iy]*sigmay[iy])*3.141592653589793*2.0)*5.0E-1 + 1.0/(sigmay[iy]*sigmay[iy])*std::pow(my[iy]-y[iy],2.0)*5.0E-1;")
    A("    }")
    A("    void fdJydy(double* dJydy, const int iy, const realtype*, const realtype*, const double* y, const double* sigmay, const double* my) override {")
    A("        dJydy[iy] = 1.0/(sigmay[iy]*sigmay[iy])*(my[iy]*2.0-y[iy]*2.0)*-5.0E-1;")
    A("    }")
    A("    void fdJydsigma(double* dJydsigma, const int iy, const realtype*, const realtype*, const double* y, const double* sigmay, const double* my) override {")
    A("        dJydsigma[iy] = 1.0/(sigmay[iy]*sigmay[iy]*sigmay[iy])*std::pow(my[iy]-y[iy],2.0)*-1.0 + 1.0/sigmay[iy];")
    A("    }")
    # nw == 0: helper-variable hooks are no-ops; event hooks are no-ops (ne == 0)
    A("    void fw(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, bool) override {}")
    A("    void fdwdx(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, bool) override {}")
    A("    void fdwdp(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, bool) override {}")
    A("    void froot(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*, const realtype*) override {}")
    A("};")
    A("inline std::unique_ptr<amici::Model> create_%s() { return std::make_unique<Model_%s>(); }" % (n, n))
    A("} // namespace pb200_oracle_models")
    with open(os.path.join(outdir, "model_%s.h" % n), "w") as fh:
        fh.write("\n".join(L) + "\n")


def generate(spec: ModelSpec, repo_root: str) -> Derived:
    d = derive(spec)
    emit_device_header(d, os.path.join(repo_root, "parpe_b200", "csrc", "models", spec.name + ".cuh"))
    emit_amici_model(d, os.path.join(repo_root, "oracle", "models"))
    inst = os.path.join(repo_root, "parpe_b200", "csrc", "models", "inst_%s.cu" % spec.name)
    with open(inst, "w") as fh:
        fh.write("// GENERATED by parpe_b200/codegen/modelgen.py -- kernel instantiations for model %s.\n"
                 "#include \"../kernels.cuh\"\n#include \"%s.cuh\"\nPB200_REGISTER_MODEL(Model_%s)\n"
                 % (spec.name, spec.name, spec.name))
    return d
