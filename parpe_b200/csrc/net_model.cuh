// net_model.cuh -- generic evaluator of table-driven mass-action network models (the model
// functions of the large-model path).  What it replaces in the reference: the generated
// per-model C++ that AMICI calls through Model_ODE::fxdot / fJSparse / fdxdotdp / fx0 / fy /
// fdydx / fsigmay (deps/AMICI/src/model_ode.cpp:14-67,109-118,145-188; generated example
// deps/AMICI/models/model_steadystate/{xdot,JSparse,dxdotdp,x0,y,dydx,sigmay}.cpp).  The tables
// are emitted by parpe_b200/codegen/netgen.py; every function here is evaluated by a whole team
// (thread block) in parallel over reactions / species / Jacobian entries.
#ifndef PB200_NET_MODEL_CUH
#define PB200_NET_MODEL_CUH

#include "pb_hd.h"

namespace pb200 {

struct NetTables {
    const char* name;
    int nx, nr, np, nk, ny, nnz;
    int n_s, n_jt, n_o;
    int jell_w;  // ELL width (max structural non-zeros per Jacobian row)
    int flops_xdot, flops_J, flops_dxdotdp;
    // reactions: v_r = kappa_r * p[pidx_r] * x[r_a] * x[r_b]   (index -1: factor 1)
    const int* r_a;
    const int* r_b;
    const double* kappa;
    const int* pidx;
    // stoichiometry, CSR by species: xdot_i = sum_e s_coef[e] * v[s_reac[e]]
    const int* s_ptr;
    const int* s_reac;
    const double* s_coef;
    // Jacobian pattern (CSR) and the terms of every entry:
    //   J_e = sum_t jt_coef[t] * kappa_r p_r * (jt_other[t] >= 0 ? x[jt_other[t]] : 1),  r = jt_reac[t]
    const int* j_ptr;
    const int* j_col;
    const int* jt_ptr;
    const int* jt_reac;
    const int* jt_other;
    const double* jt_coef;
    // ELL view (device: row-parallel J*s with coalesced loads): column of the t-th entry of row i at
    // [t*nx + i] (-1 = padding); jell_pos[e] = that position for CSR entry e
    const int* jell_col;
    const int* jell_pos;
    // CSC view of the same pattern (oracle: AMICI's JSparse)
    const int* csc_ptr;
    const int* csc_row;
    const int* csc_from_csr;
    // x0_i = x0_base[i] * (x0_kidx[i] >= 0 ? k[x0_kidx[i]] : 1);  no parameter dependence (sx0 = 0)
    const double* x0_base;
    const int* x0_kidx;
    // observables y_j = sum_e o_w[e] * x[o_idx[e]]; model sigma constant
    const int* o_ptr;
    const int* o_idx;
    const double* o_w;
    const double* sigma0;
};

namespace net {

/// vb_r = kappa_r * prod(reactants), v_r = p[pidx_r] * vb_r
template <class Team>
PB_HD void rates(const NetTables& T, const Team& tm, const double* x, const double* p, double* vb, double* v) {
    for (int r = tm.tid(); r < T.nr; r += tm.size()) {
        double w = T.kappa[r];
        const int a = T.r_a[r], b = T.r_b[r];
        if (a >= 0) w *= x[a];
        if (b >= 0) w *= x[b];
        vb[r] = w;
        if (v) v[r] = p[T.pidx[r]] * w;
    }
}

/// xdot from the reaction rates (call rates() and synchronise first); returns false (for this
/// thread's rows) if a component is not finite
template <class Team>
PB_HD bool xdot_from_rates(const NetTables& T, const Team& tm, const double* v, double* xdot) {
    bool ok = true;
    for (int i = tm.tid(); i < T.nx; i += tm.size()) {
        double s = 0.0;
        for (int e = T.s_ptr[i]; e < T.s_ptr[i + 1]; ++e) s += T.s_coef[e] * v[T.s_reac[e]];
        xdot[i] = s;
        ok = ok && PB_FINITE(s);
    }
    return ok;
}

/// Jacobian values in CSR order, or (ell) at the entries' ELL positions
template <class Team>
PB_HD bool jac_values(const NetTables& T, const Team& tm, const double* x, const double* p, double* Jv, bool ell = false) {
    bool ok = true;
    for (int e = tm.tid(); e < T.nnz; e += tm.size()) {
        double s = 0.0;
        for (int t = T.jt_ptr[e]; t < T.jt_ptr[e + 1]; ++t) {
            const int r = T.jt_reac[t];
            double w = T.jt_coef[t] * T.kappa[r] * p[T.pidx[r]];
            const int o = T.jt_other[t];
            if (o >= 0) w *= x[o];
            s += w;
        }
        Jv[ell ? T.jell_pos[e] : e] = s;
        ok = ok && PB_FINITE(s);
    }
    return ok;
}

/// dense d xdot / d p (column ip at [ip*nx]) from vb (call rates() and synchronise first); only
/// the structural non-zeros are touched, the rest stays at the zero written once per condition
template <class Team>
PB_HD void dxdotdp_dense(const NetTables& T, const Team& tm, const double* vb, double* dfdp) {
    for (int i = tm.tid(); i < T.nx; i += tm.size()) {
        for (int e = T.s_ptr[i]; e < T.s_ptr[i + 1]; ++e) dfdp[i + (size_t)T.pidx[T.s_reac[e]] * T.nx] = 0.0;
        for (int e = T.s_ptr[i]; e < T.s_ptr[i + 1]; ++e) {
            const int r = T.s_reac[e];
            dfdp[i + (size_t)T.pidx[r] * T.nx] += T.s_coef[e] * vb[r];
        }
    }
}

template <class Team>
PB_HD void x0(const NetTables& T, const Team& tm, const double* k, double* x) {
    for (int i = tm.tid(); i < T.nx; i += tm.size()) {
        const int ki = T.x0_kidx[i];
        x[i] = T.x0_base[i] * (ki >= 0 ? k[ki] : 1.0);
    }
}

/// observable j (serial)
PB_HD double observable(const NetTables& T, int j, const double* x) {
    double s = 0.0;
    for (int e = T.o_ptr[j]; e < T.o_ptr[j + 1]; ++e) s += T.o_w[e] * x[T.o_idx[e]];
    return s;
}

}  // namespace net
}  // namespace pb200
#endif
