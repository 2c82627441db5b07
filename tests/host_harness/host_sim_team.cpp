// TEST INFRASTRUCTURE -- host instantiation of the large-model integrator core.
//
// Compiles parpe_b200/csrc/bdf_team.cuh and the generated network tables as plain C++ with
// HostTeam (one thread), so the algorithm the CTA-per-condition kernel runs -- blocked
// Gauss-Jordan inversion, solve GEMM through the pivot order, sparse right-hand sides -- can be
// stepped on the CPU and diffed against the CVODES oracle (tests/test_host_team.py).  It is
// never linked into, or called by, the product library (libparpe_b200.so).
#define PB200_HOST_TRACE 1
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../parpe_b200/csrc/bdf_team.cuh"
#include "../../parpe_b200/csrc/models/net_chen.h"
#include "../../parpe_b200/csrc/models/net_froehlich.h"
#include "../../parpe_b200/csrc/models/net_net31.h"
#include "../../parpe_b200/csrc/models/net_net32.h"

using namespace pb200;

static bool tables_for(const char* model, NetTables* T) {
    if (!std::strcmp(model, "net32")) { *T = nets::net32::tables(); return true; }
    if (!std::strcmp(model, "net31")) { *T = nets::net31::tables(); return true; }
    if (!std::strcmp(model, "chen")) { *T = nets::chen::tables(); return true; }
    if (!std::strcmp(model, "froehlich")) { *T = nets::froehlich::tables(); return true; }
    return false;
}

/// symbolic factorisation of a model's pattern, as team_create_state builds it (deg_max <= 0: its rule)
static SparseLUHost symbolic_for(const NetTables& T, int deg_max) {
    if (deg_max <= 0) return SparseLUHost::build_for_team(T.nx, T.j_ptr, T.j_col, T.jell_pos);
    return SparseLUHost::build(T.nx, T.j_ptr, T.j_col, T.jell_pos, deg_max, T.nx >= 128 ? 32 : 2);
}

/// sparse != 0: sparse LU + dense Schur block (deg_max: largest degree of a sparse pivot, <= 0 = default);
/// stage_doubles: capacity of the staging area of the triangular solves (small values force column batches)
/// SS: the steady-state instantiation (t = inf output times; with k_preeq a pre-equilibration phase followed by
/// the re-initialisation of the listed states) -- the sequence simulate_team_kernel<.., true> runs per condition
template <bool SS>
static int team_sim(const char* model, const double* p_unscaled, const double* k, const int* plist, int ns,
                    int nt, const double* t, const double* my, const double* sd, double t0, double rtol,
                    double atol, int maxsteps, int stldet, int errconS, int panel_width, double* llh,
                    double* sllh, int64_t* stats, double* y_out, double* sigmay_out, int* trace, int sparse,
                    int deg_max, int stage_doubles, int fuse_history, const double* k_preeq, const int* reinit,
                    int n_reinit) {
    NetTables T;
    if (!tables_for(model, &T)) return -12345;
    SparseLUHost SH;
    if (sparse) SH = symbolic_for(T, deg_max);
    std::vector<double> hbm(TeamWork::hbm_doubles(T, sparse ? SH.nV : 0), 0.0);
    const int NB = panel_width > 0 ? panel_width : TEAM_NB_MAX;
    std::vector<double> panel((size_t)T.nx * NB), rbuf(TEAM_RBUF), rowk(TEAM_NB_MAX);
    std::vector<int> piv(TEAM_NB_MAX);
    if (stage_doubles <= 0) stage_doubles = 24064;
    std::vector<double> stage(stage_doubles > T.nx ? stage_doubles : T.nx);
    HostTeam tm;
    BdfTeam<HostTeam, SS> bdf(tm);
    bdf.T = T;
    if (sparse) bdf.SP = SH.view();
    bdf.use_sparse = sparse != 0;
    bdf.fuse_history = (fuse_history & 1) != 0;
    bdf.fuse_newton = (fuse_history & 2) != 0;
    bdf.w.carve(hbm.data(), T, sparse ? SH.nV : 0);
    bdf.w.panel = panel.data();
    bdf.w.rbuf = rbuf.data();
    bdf.w.rowk = rowk.data();
    bdf.w.piv = piv.data();
    bdf.w.stage = stage.data();
    bdf.w.stage_doubles = stage_doubles;
    bdf.w.panel_doubles = (int)panel.size();
    bdf.w.NB = NB;
    for (int i = 0; i < T.np; ++i) bdf.w.p[i] = p_unscaled[i];
    for (int i = 0; i < T.nk; ++i) bdf.w.k[i] = k[i];
    for (int i = 0; i < ns; ++i) bdf.w.plist[i] = plist[i];
    bdf.opt.t0 = t0;
    bdf.opt.maxsteps = maxsteps;
    bdf.opt.stldet = stldet;
    bdf.opt.errconS = errconS;
    ConditionIO io{nt, t, my, sd, ns, y_out, sigmay_out, trace};
    int status;
    int64_t nst_preeq = 0;
    if (SS && k_preeq) {
        for (int i = 0; i < T.nk; ++i) bdf.w.k[i] = k_preeq[i];
        status = bdf.preequilibrate(io, rtol, atol);
        nst_preeq = bdf.nst;
        if (status == SIM_SUCCESS) {
            for (int i = 0; i < T.nk; ++i) bdf.w.k[i] = k[i];
            status = bdf.simulate(io, rtol, atol, true, reinit, n_reinit);
        }
    } else {
        status = bdf.simulate(io, rtol, atol);
    }
    *llh = bdf.llh();
    for (int j = 0; j < ns; ++j) sllh[j] = bdf.w.sllh[j + 1];
    bdf.export_stats(stats, 1);
    if (SS) stats[14] = nst_preeq;  // (the cycle counters are zero on the host)
    return status;
}

extern "C" int hs_team_simulate(const char* model, const double* p_unscaled, const double* k, const int* plist, int ns,
                                int nt, const double* t, const double* my, const double* sd, double t0, double rtol,
                                double atol, int maxsteps, int stldet, int errconS, int panel_width, double* llh,
                                double* sllh, int64_t* stats, double* y_out, double* sigmay_out, int* trace, int sparse,
                                int deg_max, int stage_doubles, int fuse_history) {
    return team_sim<false>(model, p_unscaled, k, plist, ns, nt, t, my, sd, t0, rtol, atol, maxsteps, stldet, errconS,
                           panel_width, llh, sllh, stats, y_out, sigmay_out, trace, sparse, deg_max, stage_doubles,
                           fuse_history, nullptr, nullptr, 0);
}

extern "C" int hs_team_simulate_ss(const char* model, const double* p_unscaled, const double* k, const int* plist,
                                   int ns, int nt, const double* t, const double* my, const double* sd, double t0,
                                   double rtol, double atol, int maxsteps, int stldet, int errconS, int panel_width,
                                   double* llh, double* sllh, int64_t* stats, double* y_out, double* sigmay_out,
                                   int* trace, int sparse, int deg_max, int stage_doubles, int fuse_history,
                                   const double* k_preeq, const int* reinit, int n_reinit) {
    return team_sim<true>(model, p_unscaled, k, plist, ns, nt, t, my, sd, t0, rtol, atol, maxsteps, stldet, errconS,
                          panel_width, llh, sllh, stats, y_out, sigmay_out, trace, sparse, deg_max, stage_doubles,
                          fuse_history, k_preeq, reinit, n_reinit);
}

/// the blocked inversion + solve GEMM alone: x = A^-1 b for a dense column-major A given through
/// the model's sparsity pattern is not general enough for a unit test, so this entry point takes
/// gamma and the CSR Jacobian values of `model` and returns (I - gamma J)^-1 B.
extern "C" int hs_team_solve(const char* model, const double* Jcsr, double gamma, const double* B, int ncols,
                             int panel_width, double* X, int sparse, int deg_max, int stage_doubles, int* info) {
    NetTables T;
    if (!tables_for(model, &T)) return -12345;
    SparseLUHost SH;
    if (sparse) SH = symbolic_for(T, deg_max);
    std::vector<double> hbm(TeamWork::hbm_doubles(T, sparse ? SH.nV : 0), 0.0);
    const int NB = panel_width > 0 ? panel_width : TEAM_NB_MAX;
    std::vector<double> panel((size_t)T.nx * NB), rbuf(TEAM_RBUF), rowk(TEAM_NB_MAX);
    std::vector<int> piv(TEAM_NB_MAX);
    if (stage_doubles <= 0) stage_doubles = 24064;
    std::vector<double> stage(stage_doubles > T.nx ? stage_doubles : T.nx);
    HostTeam tm;
    BdfTeam<HostTeam> bdf(tm);
    bdf.T = T;
    if (sparse) bdf.SP = SH.view();
    bdf.use_sparse = sparse != 0;
    bdf.fuse_newton = false;  // (this entry point reads the solution from w.scr through acor: either works; keep the plain path)
    bdf.w.carve(hbm.data(), T, sparse ? SH.nV : 0);
    bdf.w.panel = panel.data();
    bdf.w.rbuf = rbuf.data();
    bdf.w.rowk = rowk.data();
    bdf.w.piv = piv.data();
    bdf.w.stage = stage.data();
    bdf.w.stage_doubles = stage_doubles;
    bdf.w.panel_doubles = (int)panel.size();
    bdf.w.NB = NB;
    if (info && sparse) {
        info[0] = SH.n1; info[1] = SH.m; info[2] = SH.nnzL; info[3] = SH.nnzU;
        info[4] = (int)(SH.n_pairs_sparse + SH.n_pairs_schur);
        info[5] = (int)SH.tlevel_ptr.size() - 1; info[6] = (int)SH.flev_ptr.size() - 1; info[7] = (int)SH.blev_ptr.size() - 1;
    }
    bdf.nx = T.nx;
    bdf.ns = ncols - 1;
    bdf.ncol = ncols;
    bdf.gamma = gamma;
    bdf.gamrat = 1.0;
    if (ncols > T.np + 1) return -2;
    for (int e = 0; e < T.nnz; ++e) bdf.w.Jsaved[T.jell_pos[e]] = Jcsr[e];
    bdf.dn = T.nx;
    bdf.lin_sparse = false;
    bdf.nfallback = 0;
    bdf.cyc_spfactor = bdf.cyc_sptri = bdf.cyc_panel = bdf.cyc_update = bdf.cyc_solve = 0;
    if (sparse) {
        const int r = bdf.factor_sparse();
        if (r != 0) return 10 + r;
    } else if (bdf.invert_newton_matrix() != 0) {
        return 1;
    }
    for (size_t e = 0; e < (size_t)T.nx * ncols; ++e) {
        bdf.w.b[e] = B[e];
        bdf.w.acor[e] = 0.0;
        bdf.w.ewt[e] = 1.0;
    }
    bdf.newton_update(0, 0);
    if (ncols > 1) bdf.newton_update(1, ncols - 1);
    for (size_t e = 0; e < (size_t)T.nx * ncols; ++e) X[e] = -bdf.w.acor[e];  // newton_update applies the factor -1
    return 0;
}
