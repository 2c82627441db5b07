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

extern "C" int hs_team_simulate(const char* model, const double* p_unscaled, const double* k, const int* plist, int ns,
                                int nt, const double* t, const double* my, const double* sd, double t0, double rtol,
                                double atol, int maxsteps, int stldet, int errconS, int panel_width, double* llh,
                                double* sllh, int64_t* stats, double* y_out, double* sigmay_out, int* trace) {
    NetTables T;
    if (!tables_for(model, &T)) return -12345;
    std::vector<double> hbm(TeamWork::hbm_doubles(T), 0.0);
    const int NB = panel_width > 0 ? panel_width : TEAM_NB_MAX;
    std::vector<double> panel((size_t)T.nx * NB), rbuf(TEAM_RBUF), rowk(TEAM_NB_MAX);
    std::vector<int> piv(TEAM_NB_MAX);
    HostTeam tm;
    BdfTeam<HostTeam> bdf(tm);
    bdf.T = T;
    bdf.w.carve(hbm.data(), T);
    bdf.w.panel = panel.data();
    bdf.w.rbuf = rbuf.data();
    bdf.w.rowk = rowk.data();
    bdf.w.piv = piv.data();
    bdf.w.NB = NB;
    for (int i = 0; i < T.np; ++i) bdf.w.p[i] = p_unscaled[i];
    for (int i = 0; i < T.nk; ++i) bdf.w.k[i] = k[i];
    for (int i = 0; i < ns; ++i) bdf.w.plist[i] = plist[i];
    bdf.opt.t0 = t0;
    bdf.opt.maxsteps = maxsteps;
    bdf.opt.stldet = stldet;
    bdf.opt.errconS = errconS;
    ConditionIO io{nt, t, my, sd, ns, y_out, sigmay_out, trace};
    const int status = bdf.simulate(io, rtol, atol);
    *llh = bdf.llh();
    for (int j = 0; j < ns; ++j) sllh[j] = bdf.w.sllh[j + 1];
    bdf.export_stats(stats, 1);
    return status;
}

/// the blocked inversion + solve GEMM alone: x = A^-1 b for a dense column-major A given through
/// the model's sparsity pattern is not general enough for a unit test, so this entry point takes
/// gamma and the CSR Jacobian values of `model` and returns (I - gamma J)^-1 B.
extern "C" int hs_team_solve(const char* model, const double* Jcsr, double gamma, const double* B, int ncols,
                             int panel_width, double* X) {
    NetTables T;
    if (!tables_for(model, &T)) return -12345;
    std::vector<double> hbm(TeamWork::hbm_doubles(T), 0.0);
    const int NB = panel_width > 0 ? panel_width : TEAM_NB_MAX;
    std::vector<double> panel((size_t)T.nx * NB), rbuf(TEAM_RBUF), rowk(TEAM_NB_MAX);
    std::vector<int> piv(TEAM_NB_MAX);
    HostTeam tm;
    BdfTeam<HostTeam> bdf(tm);
    bdf.T = T;
    bdf.w.carve(hbm.data(), T);
    bdf.w.panel = panel.data();
    bdf.w.rbuf = rbuf.data();
    bdf.w.rowk = rowk.data();
    bdf.w.piv = piv.data();
    bdf.w.NB = NB;
    bdf.nx = T.nx;
    bdf.ns = ncols - 1;
    bdf.ncol = ncols;
    bdf.gamma = gamma;
    bdf.gamrat = 1.0;
    if (ncols > T.np + 1) return -2;
    for (int e = 0; e < T.nnz; ++e) bdf.w.Jsaved[T.jell_pos[e]] = Jcsr[e];
    if (bdf.invert_newton_matrix() != 0) return 1;
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
