// bdf_core.cuh -- batched stiff BDF integrator with Newton iteration and staggered
// forward sensitivities, one *lane group* per experimental condition.
//
// What it replaces in the reference: for one
// condition, amici::ForwardProblem::workForwardProblem
// (deps/AMICI/src/forwardproblem.cpp:67-203) driving CVODES
// (deps/AMICI/ThirdParty/sundials/src/cvodes/cvodes.c: CVode :2890-3636, cvHin
// :5436-5549, cvStep :5800-6094, cvSetBDF :6788-6879, cvNls :6894-7007,
// cvStgrNls :7167-7210, cvDoErrorTest :7461-7614, cvCompleteStep :7634-7736,
// cvPrepareNextStep..cvChooseEta :7747-8000, cvBDFStab/cvSLdet :8158-8556), the
// Newton solver (src/sunnonlinsol/newton/sunnonlinsol_newton.c:176-335) with its
// CVODES glue (cvodes_nls.c, cvodes_nls_stg.c), the linear-solver interface
// (cvodes_ls.c:1386-1470,1611-1708,1720-1927) on a dense LU with partial pivoting
// (src/sundials/sundials_dense.c:82-182), and the likelihood/gradient
// accumulation at the output times (deps/AMICI/src/model.cpp:1106-1152,
// 2012-2392; rdata.cpp:305-370,644-716).  CVODES' heuristics and constants are
// kept so that the step sequence stays close to the reference's; the data layout
// and the execution model are not CVODES':
//
//   * a condition is owned by a group of G lanes of one warp (G = 1..32);
//   * the state vector and every sensitivity vector is a *column*; column c is
//     held by lane (c mod G), entirely in registers, for all Nordsieck orders
//     (zn[0..5][column][NX]); CPL = ceil((1+NP)/G) columns per lane;
//   * vector operations are therefore lane-local and fully unrolled; norms are
//     warp-shuffle reductions (max over sensitivity columns, broadcast of the
//     state column);
//   * the Jacobian, its LU factors and the group-uniform BDF coefficient arrays
//     live once per group in shared memory (GroupShared); the state lane
//     factorises, every lane back-substitutes its own columns against the same
//     factors (the staggered corrector's "same matrix, nplist right-hand sides");
//   * model functions are the generated __host__ __device__ code of
//     csrc/models/<name>.cuh.
//
// The same source compiles for the host with the SingleLane group (G = 1); that
// instantiation exists only in tests/host_harness to diff the algorithm against
// the CVODES oracle without a GPU.  The product library has no host integrator.
#ifndef PB200_BDF_CORE_CUH
#define PB200_BDF_CORE_CUH

#include <math.h>
#include <stdint.h>

#include "pb_hd.h"

namespace pb200 {

// ---- status codes (values of CVODES / AMICI, include/parpe_b200.h) ------------
enum : int {
    SIM_SUCCESS = 0,
    SIM_TOO_MUCH_WORK = -1,
    SIM_TOO_MUCH_ACC = -2,
    SIM_ERR_FAILURE = -3,
    SIM_CONV_FAILURE = -4,
    SIM_LSETUP_FAIL = -6,
    SIM_RHSFUNC_FAIL = -8,
    SIM_FIRST_RHSFUNC_ERR = -9,
    SIM_REPTD_RHSFUNC_ERR = -10,
    SIM_UNREC_RHSFUNC_ERR = -11,
    SIM_ILL_INPUT = -22,
    SIM_TOO_CLOSE = -27,
    SIM_FIRST_SRHSFUNC_ERR = -42,
    SIM_REPTD_SRHSFUNC_ERR = -43,
    SIM_UNREC_SRHSFUNC_ERR = -44,
    SIM_ERROR = -99
};

enum : int {
    STAT_NST = 0, STAT_NFE, STAT_NSETUPS, STAT_NJE, STAT_NNI, STAT_NNIS, STAT_NFSE, STAT_NETF, STAT_NETFS,
    STAT_NCFN, STAT_NCFNS, STAT_NOR, STAT_TRIALS, STAT_QSUM, NSTATS = 16
};

// ---- CVODES constants (cvodes_impl.h:57-131,202-204; cvodes.c:298-305;
//      cvodes_ls_impl.h:40-42) --------------------------------------------------
namespace cv {
constexpr int Q_MAX = 5;          // BDF_Q_MAX
constexpr int L_MAX = Q_MAX + 1;
constexpr int NLS_MAXCOR = 3;
constexpr double CRDOWN = 0.3;
constexpr double RDIV = 2.0;
constexpr int MXNCF = 10;
constexpr int MXNEF = 7;
constexpr int MXNEF1 = 3;
constexpr int SMALL_NEF = 2;
constexpr int LONG_WAIT = 10;
constexpr double ETA_MIN_FX = 0.0;
constexpr double ETA_MAX_FX = 1.5;
constexpr double ETA_MAX_FS = 1.0e4;
constexpr double ETA_MAX_ES = 10.0;
constexpr double ETA_MAX_GS = 10.0;
constexpr double ETA_MIN = 0.1;
constexpr double ETA_MAX_EF = 0.2;
constexpr double ETA_MIN_EF = 0.1;
constexpr double ETA_CF = 0.25;
constexpr int SMALL_NST = 10;
constexpr double ONEPSM = 1.000001;
constexpr double ADDON = 1.0e-6;
constexpr double BIAS1 = 6.0;
constexpr double BIAS2 = 6.0;
constexpr double BIAS3 = 10.0;
constexpr int MSBP = 20;
constexpr double DGMAX_LSETUP = 0.3;
constexpr double CORTES = 0.1;
constexpr int MSBJ = 51;
constexpr double DGMAX_JBAD = 0.2;
constexpr double HLB_FACTOR = 100.0;
constexpr double HUB_FACTOR = 0.1;
constexpr double H_BIAS = 0.5;
constexpr int HIN_MAX_ITERS = 4;
constexpr double FUZZ_FACTOR = 100.0;
constexpr double UROUND = 2.220446049250313e-16;
constexpr double TINY = 1.0e-10;
// internal flags (cvodes.c:200-260)
constexpr int FIRST_CALL = 101, PREV_CONV_FAIL = 102, PREV_ERR_FAIL = 103;
constexpr int CONV_RECVR = 902;   // SUN_NLS_CONV_RECVR
constexpr int RHSFUNC_RECVR = 10, SRHSFUNC_RECVR = 12;
constexpr int NO_FAILURES = 0, FAIL_BAD_J = 1, FAIL_OTHER = 2;
constexpr int DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5;
constexpr int TSTOP_RETURN = 1;
}  // namespace cv

struct SolverOptions {
    double t0;
    double rtol, atol;
    int maxsteps;
    int stldet;
    int errconS;
};

// ---- lane groups ---------------------------------------------------------------
/// One lane owns every column (host harness; also a valid device mapping).
struct SingleLane {
    static constexpr int G = 1;
    PB_HD int lane() const { return 0; }
    PB_HD double max(double v) const { return v; }
    PB_HD double bcast0(double v) const { return v; }
    PB_HD int bcast0(int v) const { return v; }
    PB_HD void sync() const {}
};

#if defined(__CUDACC__)
/// G lanes of one warp (G a power of two <= 32) cooperate on one condition.
template <int G_>
struct LaneGroup {
    static constexpr int G = G_;
    unsigned mask;
    int ln, base;
    __device__ __forceinline__ LaneGroup() {
        const int wl = threadIdx.x & 31;
        ln = wl % G;
        base = wl - ln;
        mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << base);
    }
    __device__ __forceinline__ int lane() const { return ln; }
    __device__ __forceinline__ double max(double v) const {
        PB_UNROLL
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
        return v;
    }
    __device__ __forceinline__ double bcast0(double v) const { return __shfl_sync(mask, v, base); }
    __device__ __forceinline__ int bcast0(int v) const { return __shfl_sync(mask, v, base); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
#endif

/// Group-uniform state: one copy per condition (shared memory on the device).
template <class M>
struct GroupShared {
    static constexpr int NX = M::NX;
    double Jsaved[NX * NX];  // Jacobian at its last evaluation (cvLs savedJ)
    double LU[NX * NX];      // factors of I - gamma*J
    double Jcur[NX * NX];    // J(tn, y_n) for the sensitivity right-hand sides
    double ysh[NX];          // state handed from the state lane to the sensitivity lanes
    double p[M::NP > 0 ? M::NP : 1];  // unscaled model parameters
    double k[M::NK > 0 ? M::NK : 1];  // constants (fixed parameters)
    double l[cv::L_MAX + 1];
    double tau[cv::L_MAX + 1];
    double tq[6];
    double ssdat[6][4];      // STALD history (cvBDFStab)
    int piv[NX];
    int plist[M::NP > 0 ? M::NP : 1];
};

/// Per-condition inputs (device pointers into the resident tables).
struct ConditionIO {
    int nt;
    const double* t;      // [nt]
    const double* my;     // [nt*ny] measurements, NaN = none
    const double* sd;     // [nt*ny] sigma from data, NaN = model sigma
    int ns;               // number of sensitivity columns (nplist), 0 = cost only
    double* y_out;        // nullable [nt*ny]
    double* sigmay_out;   // nullable [nt*ny]
#if defined(PB200_HOST_TRACE)
    int* trace;           // host harness only: (nst, q) at every output time
#endif
};

template <class M, class Grp, int CPL>
struct Bdf {
    static constexpr int NX = M::NX, NP = M::NP, NK = M::NK, NY = M::NY;
    static constexpr int G = Grp::G;

    Grp g;
    GroupShared<M>* sh;
    SolverOptions opt;
    int ns;            // sensitivity columns
    bool sensi;        // ns > 0
    bool act[CPL];     // column exists
    int ipar[CPL];     // model-parameter index of a sensitivity column (-1: state column)

    // columns in registers
    double zn[cv::L_MAX][CPL][NX];
    double acor[CPL][NX];
    double ewt[CPL][NX];
    double tmp[CPL][NX];  // y / yS work vector
    double b[CPL][NX];    // residual -> Newton update

    // integrator scalars (uniform within the group)
    double tn, h, hprime, hscale, eta, etamax, etaq, etaqm1, etaqp1;
    double gamma, gammap, gamrat, rl1, crate, crateS, delp, acnrm, acnrmS, saved_tq5, tstop;
    int q, qprime, L, qwait, nst, nscon, nstlp, nstlj, convfail;
    bool jcur, jcurNLS, jcurNLSS, tstopset;
    // counters
    int nfe, nsetups, nje, nni, nniS, nfSe, netf, netfS, ncfn, ncfnS, nor, qsum;

    PB_HD bool is_state(int m) const { return m == 0 && g.lane() == 0; }

    // ---------------------------------------------------------------- helpers
    PB_HD static bool finite_vec(const double* v) {
        bool ok = true;
        PB_UNROLL
        for (int i = 0; i < NX; ++i) ok = ok && isfinite(v[i]);
        return ok;
    }

    /// WRMS norms of v: state column -> *ns_, max over sensitivity columns -> *nS
    PB_HD void norms(const double (&v)[CPL][NX], double& nstate, double& nsens) const {
        double ls = 0.0, lS = 0.0;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m]) continue;
            double s = 0.0;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double prod = v[m][i] * ewt[m][i];
                s += prod * prod;
            }
            s = sqrt(s / NX);
            if (is_state(m)) ls = s;
            else if (s > lS) lS = s;
        }
        nstate = g.bcast0(ls);
        nsens = sensi ? g.max(lS) : 0.0;
    }
    PB_HD double norm_state(const double (&v)[CPL][NX]) const {
        double ls = 0.0;
        if (g.lane() == 0) {
            double s = 0.0;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double prod = v[0][i] * ewt[0][i];
                s += prod * prod;
            }
            ls = sqrt(s / NX);
        }
        return g.bcast0(ls);
    }
    PB_HD double norm_sens(const double (&v)[CPL][NX]) const {
        double lS = 0.0;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            double s = 0.0;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double prod = v[m][i] * ewt[m][i];
                s += prod * prod;
            }
            s = sqrt(s / NX);
            if (s > lS) lS = s;
        }
        return g.max(lS);
    }
    /// norm over all variables in the error test: state, plus sensitivities if errconS
    PB_HD double norm_all(const double (&v)[CPL][NX]) const {
        if (sensi && opt.errconS) {
            double a, s;
            norms(v, a, s);
            return a > s ? a : s;  // cvSensUpdateNorm, cvodes.c:9449-9457
        }
        return norm_state(v);
    }

    /// cvEwtSetSS / cvSensEwtSetSS (cvodes.c:9118-9129, 9263-9279)
    PB_HD bool ewt_set() {
        bool ok = true;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m]) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double w = opt.rtol * fabs(zn[0][m][i]) + opt.atol;
                ok = ok && (w > 0.0);
                ewt[m][i] = 1.0 / w;
            }
        }
        return g.max(ok ? 0.0 : 1.0) == 0.0;
    }

    /// state right-hand side on the state lane; returns 0 or RHSFUNC_RECVR (checkFinite,
    /// deps/AMICI/src/solver_cvodes.cpp:1149-1173)
    PB_HD int f_state(double t, const double* y, double* out) {
        int bad = 0;
        if (g.lane() == 0) {
            M::xdot(out, t, y, sh->p, sh->k);
            bad = finite_vec(out) ? 0 : 1;
        }
        nfe++;
        return g.bcast0(bad) ? cv::RHSFUNC_RECVR : 0;
    }

    /// hand the state lane's vector to every lane through shared memory
    PB_HD void publish_state(const double* y) {
        g.sync();
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) sh->ysh[i] = y[i];
        }
        g.sync();
    }

    /// J(t, ysh) -> sh->Jcur by the state lane (Model_ODE::fsxdot at ip == 0,
    /// deps/AMICI/src/model_ode.cpp:492-498)
    PB_HD void refresh_Jcur(double t) {
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX * NX; ++i) sh->Jcur[i] = 0.0;
            M::jac(sh->Jcur, t, sh->ysh, sh->p, sh->k);
        }
        g.sync();
    }

    /// sensitivity right-hand sides for all columns: out = dxdotdp(ip) + Jcur * s
    /// (Model_ODE::fsxdot, model_ode.cpp:486-515); state is sh->ysh.
    PB_HD int f_sens(double t, const double (&s)[CPL][NX], double (&out)[CPL][NX]) {
        double bad = 0.0;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) out[m][i] = 0.0;
            M::dxdotdp(out[m], t, sh->ysh, sh->p, sh->k, ipar[m]);
            PB_UNROLL
            for (int j = 0; j < NX; ++j) {
                const double sj = s[m][j];
                PB_UNROLL
                for (int r = 0; r < NX; ++r) out[m][r] += sh->Jcur[r + j * NX] * sj;
            }
            if (!finite_vec(out[m])) bad = 1.0;
        }
        nfSe += ns;
        return g.max(bad) != 0.0 ? cv::SRHSFUNC_RECVR : 0;
    }

    // ---------------------------------------------------------------- linear algebra
    /// cvLsSetup + cvLsLinSys + dense GETRF (cvodes_ls.c:1611-1708,1386-1470;
    /// sundials_dense.c:82-148).  y = state the Jacobian is evaluated at.
    PB_HD int lsetup(const double* y) {
        const double dgamma = fabs((gamma / gammap) - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + cv::MSBJ) ||
                          ((convfail == cv::FAIL_BAD_J) && (dgamma < cv::DGMAX_JBAD)) ||
                          (convfail == cv::FAIL_OTHER);
        int ret = 0;
        if (g.lane() == 0) {
            if (jbad) {
                PB_UNROLL
                for (int i = 0; i < NX * NX; ++i) sh->Jsaved[i] = 0.0;
                M::jac(sh->Jsaved, tn, y, sh->p, sh->k);
                bool ok = true;
                for (int i = 0; i < NX * NX; ++i) ok = ok && isfinite(sh->Jsaved[i]);
                if (!ok) ret = 1;
            }
            if (ret == 0) {
                // A = I - gamma*J
                for (int c = 0; c < NX; ++c)
                    for (int r = 0; r < NX; ++r)
                        sh->LU[r + c * NX] = (r == c ? 1.0 : 0.0) - gamma * sh->Jsaved[r + c * NX];
                // LU with partial pivoting, column-oriented
                for (int k = 0; k < NX; ++k) {
                    double* col_k = sh->LU + k * NX;
                    int l = k;
                    for (int i = k + 1; i < NX; ++i)
                        if (fabs(col_k[i]) > fabs(col_k[l])) l = i;
                    sh->piv[k] = l;
                    if (col_k[l] == 0.0) { ret = 1; break; }
                    if (l != k) {
                        for (int i = 0; i < NX; ++i) {
                            const double t_ = sh->LU[l + i * NX];
                            sh->LU[l + i * NX] = sh->LU[k + i * NX];
                            sh->LU[k + i * NX] = t_;
                        }
                    }
                    const double mult = 1.0 / col_k[k];
                    for (int i = k + 1; i < NX; ++i) col_k[i] *= mult;
                    for (int j = k + 1; j < NX; ++j) {
                        double* col_j = sh->LU + j * NX;
                        const double a_kj = col_j[k];
                        if (a_kj != 0.0)
                            for (int i = k + 1; i < NX; ++i) col_j[i] -= a_kj * col_k[i];
                    }
                    // the back-substitution multiplies by the reciprocal pivot
                    col_k[k] = mult;
                }
            }
        }
        if (jbad) { nje++; nstlj = nst; }
        jcur = jbad;
        nsetups++;
        g.sync();
        return g.bcast0(ret);
    }

    /// dense GETRS on one register column + cvLsSolve's gamma-ratio scaling
    /// (sundials_dense.c:150-182; cvodes_ls.c:1849-1854)
    PB_HD void lsolve_col(double* x) const {
        PB_UNROLL
        for (int k = 0; k < NX; ++k) {
            const int pk = sh->piv[k];
            const double xk = x[k];
            PB_UNROLL
            for (int j = k + 1; j < NX; ++j) {
                if (pk == j) { x[k] = x[j]; x[j] = xk; }
            }
        }
        PB_UNROLL
        for (int k = 0; k < NX - 1; ++k) {
            PB_UNROLL
            for (int i = k + 1; i < NX; ++i) x[i] -= sh->LU[i + k * NX] * x[k];
        }
        PB_UNROLL
        for (int k = NX - 1; k > 0; --k) {
            x[k] *= sh->LU[k + k * NX];
            PB_UNROLL
            for (int i = 0; i < k; ++i) x[i] -= sh->LU[i + k * NX] * x[k];
        }
        x[0] *= sh->LU[0];
        if (gamrat != 1.0) {
            const double sc = 2.0 / (1.0 + gamrat);
            PB_UNROLL
            for (int i = 0; i < NX; ++i) x[i] *= sc;
        }
    }

    // ---------------------------------------------------------------- BDF coefficients
    /// cvSetBDF + cvSetTqBDF + cvSet (cvodes.c:6788-6879, 6617-6629); coefficients are
    /// computed redundantly by every lane (a few dozen flops) and kept in shared memory.
    PB_HD void cvset() {
        double l_[cv::L_MAX + 1];
        double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
        l_[0] = l_[1] = xi_inv = xistar_inv = 1.0;
        for (int i = 2; i <= cv::Q_MAX; ++i) l_[i] = 0.0;
        l_[cv::L_MAX] = 0.0;
        alpha0 = alpha0_hat = -1.0;
        hsum = h;
        if (q > 1) {
            for (int j = 2; j < q; ++j) {
                hsum += sh->tau[j - 1];
                xi_inv = h / hsum;
                alpha0 -= 1.0 / j;
                for (int i = j; i >= 1; --i) l_[i] += l_[i - 1] * xi_inv;
            }
            alpha0 -= 1.0 / q;
            xistar_inv = -l_[1] - alpha0;
            hsum += sh->tau[q - 1];
            xi_inv = h / hsum;
            alpha0_hat = -l_[1] - xi_inv;
            for (int i = q; i >= 1; --i) l_[i] += l_[i - 1] * xistar_inv;
        }
        // cvSetTqBDF
        double tq_[6];
        tq_[1] = sh->tq[1];
        tq_[3] = sh->tq[3];
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        tq_[2] = fabs(A1 / (alpha0 * A2));
        tq_[5] = fabs(A2 * xistar_inv / (l_[q] * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double C = xistar_inv / l_[q];
                const double A3 = alpha0 + 1.0 / q;
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = (1.0 - A4 + A3) / A3;
                tq_[1] = fabs(C * Cpinv);
            } else {
                tq_[1] = 1.0;
            }
            hsum += sh->tau[q];
            xi_inv = h / hsum;
            const double A5 = alpha0 - (1.0 / (q + 1));
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = (1.0 - A6 + A5) / A2;
            tq_[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
        }
        tq_[4] = cv::CORTES / tq_[2];
        g.sync();
        if (g.lane() == 0) {
            for (int i = 0; i <= cv::L_MAX; ++i) sh->l[i] = l_[i];
            for (int i = 1; i <= 5; ++i) sh->tq[i] = tq_[i];
        }
        g.sync();
        rl1 = 1.0 / l_[1];
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? gamma / gammap : 1.0;
    }

    // ---------------------------------------------------------------- history array
    /// cvPredict (cvodes.c:6507-6597)
    PB_HD void predict() {
        tn += h;
        if (tstopset) {
            if ((tn - tstop) * h > 0.0) tn = tstop;
        }
        PB_UNROLL
        for (int k = 1; k <= cv::Q_MAX; ++k) {
            PB_UNROLL
            for (int j = cv::Q_MAX; j >= 1; --j) {
                if (j >= k && j <= q && k <= q) {
                    PB_UNROLL
                    for (int m = 0; m < CPL; ++m) {
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) zn[j - 1][m][i] += zn[j][m][i];
                    }
                }
            }
        }
    }
    /// cvRestore (cvodes.c:7373-7424)
    PB_HD void restore(double saved_t) {
        tn = saved_t;
        PB_UNROLL
        for (int k = 1; k <= cv::Q_MAX; ++k) {
            PB_UNROLL
            for (int j = cv::Q_MAX; j >= 1; --j) {
                if (j >= k && j <= q && k <= q) {
                    PB_UNROLL
                    for (int m = 0; m < CPL; ++m) {
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) zn[j - 1][m][i] -= zn[j][m][i];
                    }
                }
            }
        }
    }
    /// cvRescale (cvodes.c:6425-6495)
    PB_HD void rescale() {
        double factor = eta;
        PB_UNROLL
        for (int j = 1; j <= cv::Q_MAX; ++j) {
            if (j <= q) {
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) zn[j][m][i] *= factor;
                }
                factor *= eta;
            }
        }
        h = hscale * eta;
        hscale = h;
        nscon = 0;
    }
    /// zn[j] += c_j * zn[src] for j = 2..jmax (N_VScaleAddMulti on columns 2..)
    PB_HD void add_multiple_of(int src, const double* coef, int jmax) {
        PB_UNROLL
        for (int s = 2; s <= cv::Q_MAX; ++s) {
            if (s != src) continue;
            PB_UNROLL
            for (int j = 2; j <= cv::Q_MAX; ++j) {
                if (j <= jmax && j != s) {
                    const double cj = coef[j];
                    PB_UNROLL
                    for (int m = 0; m < CPL; ++m) {
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) zn[j][m][i] += cj * zn[s][m][i];
                    }
                }
            }
        }
    }
    /// cvIncreaseBDF (cvodes.c:6256-6351): new column zn[L] from the saved correction zn[qmax]
    PB_HD void increase_order() {
        double lc[cv::L_MAX + 1];
        for (int i = 0; i <= cv::L_MAX; ++i) lc[i] = 0.0;
        double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, xi;
        lc[2] = 1.0;
        double hsum = hscale;
        if (q > 1) {
            for (int j = 1; j < q; ++j) {
                hsum += sh->tau[j + 1];
                xi = hsum / hscale;
                prod *= xi;
                alpha0 -= 1.0 / (j + 1);
                alpha1 += 1.0 / xi;
                for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xiold + lc[i - 1];
                xiold = xi;
            }
        }
        const double A1 = (-alpha0 - alpha1) / prod;
        // zn[L] = A1 * zn[indx_acor = qmax]   (L = q+1 <= qmax)
        PB_UNROLL
        for (int j = 2; j <= cv::Q_MAX; ++j) {
            if (j == L) {
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) zn[j][m][i] = A1 * zn[cv::Q_MAX][m][i];
                }
            }
        }
        if (q > 1) add_multiple_of(L, lc, q);
    }
    /// cvDecreaseBDF (cvodes.c:6363-6415)
    PB_HD void decrease_order() {
        double lc[cv::L_MAX + 1];
        for (int i = 0; i <= cv::L_MAX; ++i) lc[i] = 0.0;
        lc[2] = 1.0;
        double hsum = 0.0;
        for (int j = 1; j <= q - 2; ++j) {
            hsum += sh->tau[j];
            const double xi = hsum / hscale;
            for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xi + lc[i - 1];
        }
        if (q > 2) {
            for (int j = 2; j < q; ++j) lc[j] = -lc[j];
            add_multiple_of(q, lc, q - 1);
        }
    }
    /// cvAdjustOrder (cvodes.c:6134-6143)
    PB_HD void adjust_order(int deltaq) {
        if ((q == 2) && (deltaq != 1)) return;
        if (deltaq == 1) increase_order();
        else if (deltaq == -1) decrease_order();
    }
    /// cvAdjustParams (cvodes.c:6112-6122)
    PB_HD void adjust_params() {
        if (qprime != q) {
            adjust_order(qprime - q);
            q = qprime;
            L = q + 1;
            qwait = L;
        }
        rescale();
    }

    // ---------------------------------------------------------------- nonlinear solves
    /// cvNlsResidual (cvodes_nls.c:344-371): b(state) = rl1*zn1 + acor - gamma*f(tn, zn0+acor)
    PB_HD int residual_state() {
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) tmp[0][i] = zn[0][0][i] + acor[0][i];
        }
        double ft[NX];
        const int r = f_state(tn, tmp[0], ft);
        if (r) return r;
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double r1 = rl1 * zn[1][0][i] + acor[0][i];
                b[0][i] = -gamma * ft[i] + r1;
            }
        }
        return 0;
    }

    /// cvNls + SUNNonlinSolSolve_Newton for the state (cvodes.c:6894-7007;
    /// sunnonlinsol_newton.c:176-335; cvNlsConvTest cvodes_nls.c:298-342)
    PB_HD int nls_state(int nflag) {
        convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::NO_FAILURES : cv::FAIL_OTHER;
        bool callSetup = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                         (nst >= nstlp + cv::MSBP) || (fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) acor[0][i] = 0.0;
        }
        bool jbad = false;
        int ret;
        for (;;) {
            ret = residual_state();
            if (ret == 0 && callSetup) {
                if (jbad) convfail = cv::FAIL_BAD_J;
                const int lr = lsetup(tmp[0]);
                jcurNLS = jcur;
                gamrat = 1.0;
                gammap = gamma;
                crate = 1.0;
                crateS = 1.0;
                nstlp = nst;
                if (lr != 0) ret = cv::CONV_RECVR;
            }
            if (ret == 0) {
                int m = 0;
                for (;;) {
                    nni++;
                    double del = 0.0;
                    if (g.lane() == 0) {
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) b[0][i] = -b[0][i];
                        lsolve_col(b[0]);
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) acor[0][i] += b[0][i];
                    }
                    del = norm_state(b);
                    if (m > 0) crate = fmax(cv::CRDOWN * crate, del / delp);
                    const double dcon = del * fmin(1.0, crate) / sh->tq[4];
                    if (dcon <= 1.0) {
                        acnrm = (m == 0) ? del : norm_state(acor);
                        jcurNLS = false;
                        // y = zn0 + acor; Jacobian data no longer current
                        if (g.lane() == 0) {
                            PB_UNROLL
                            for (int i = 0; i < NX; ++i) tmp[0][i] = zn[0][0][i] + acor[0][i];
                        }
                        jcur = false;
                        return 0;
                    }
                    if ((m >= 1) && (del > cv::RDIV * delp)) { ret = cv::CONV_RECVR; break; }
                    delp = del;
                    m++;
                    if (m >= cv::NLS_MAXCOR) { ret = cv::CONV_RECVR; break; }
                    ret = residual_state();
                    if (ret != 0) break;
                }
            }
            if ((ret > 0) && !jcurNLS) {
                callSetup = true;
                jbad = true;
                if (g.lane() == 0) {
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) acor[0][i] = 0.0;
                }
                continue;
            }
            break;
        }
        return ret;
    }

    /// cvNlsResidualSensStg (cvodes_nls_stg.c:377-421) for every sensitivity column
    PB_HD int residual_sens() {
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) tmp[m][i] = zn[0][m][i] + acor[m][i];
        }
        double fs[CPL][NX];
        const int r = f_sens(tn, tmp, fs);
        if (r) return r;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) b[m][i] = rl1 * zn[1][m][i] + acor[m][i] - gamma * fs[m][i];
        }
        return 0;
    }

    /// cvStgrNls + Newton for all sensitivities at once with the state's factors
    /// (cvodes.c:7167-7210; cvodes_nls_stg.c:238-375).  y_n is in tmp[state] / sh->ysh.
    PB_HD int nls_sens() {
        bool callSetup = false;
        bool jbad = false;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) acor[m][i] = 0.0;
        }
        int ret;
        for (;;) {
            ret = residual_sens();
            if (ret == 0 && callSetup) {
                if (jbad) convfail = cv::FAIL_BAD_J;
                double ycur[NX];
                PB_UNROLL
                for (int i = 0; i < NX; ++i) ycur[i] = sh->ysh[i];
                const int lr = lsetup(ycur);
                jcurNLSS = jcur;
                gamrat = 1.0;
                gammap = gamma;
                crate = 1.0;
                crateS = 1.0;
                nstlp = nst;
                if (lr != 0) ret = cv::CONV_RECVR;
            }
            if (ret == 0) {
                int m_it = 0;
                for (;;) {
                    nniS++;
                    PB_UNROLL
                    for (int m = 0; m < CPL; ++m) {
                        if (!act[m] || is_state(m)) continue;
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) b[m][i] = -b[m][i];
                        lsolve_col(b[m]);
                        PB_UNROLL
                        for (int i = 0; i < NX; ++i) acor[m][i] += b[m][i];
                    }
                    const double Del = norm_sens(b);
                    if (m_it > 0) crateS = fmax(cv::CRDOWN * crateS, Del / delp);
                    const double dcon = Del * fmin(1.0, crateS) / sh->tq[4];
                    if (dcon <= 1.0) {
                        if (opt.errconS) acnrmS = (m_it == 0) ? Del : norm_sens(acor);
                        jcurNLSS = false;
                        jcur = false;
                        return 0;
                    }
                    if ((m_it >= 1) && (Del > cv::RDIV * delp)) { ret = cv::CONV_RECVR; break; }
                    delp = Del;
                    m_it++;
                    if (m_it >= cv::NLS_MAXCOR) { ret = cv::CONV_RECVR; break; }
                    ret = residual_sens();
                    if (ret != 0) break;
                }
            }
            if ((ret > 0) && !jcurNLSS) {
                callSetup = true;
                jbad = true;
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    if (!act[m] || is_state(m)) continue;
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) acor[m][i] = 0.0;
                }
                continue;
            }
            break;
        }
        return ret;
    }

    // ---------------------------------------------------------------- step control
    /// cvHandleNFlag (cvodes.c:7309-7363)
    PB_HD int handle_nflag(int& nflag, double saved_t, int& ncf, int& ncfn_total) {
        if (nflag == 0) return cv::DO_ERROR_TEST;
        ncfn_total++;
        restore(saved_t);
        if (nflag < 0) return nflag;
        ncf++;
        etamax = 1.0;
        if ((fabs(h) <= 0.0 * cv::ONEPSM) || (ncf == cv::MXNCF)) {
            if (nflag == cv::CONV_RECVR) return SIM_CONV_FAILURE;
            if (nflag == cv::RHSFUNC_RECVR) return SIM_REPTD_RHSFUNC_ERR;
            if (nflag == cv::SRHSFUNC_RECVR) return SIM_REPTD_SRHSFUNC_ERR;
        }
        eta = fmax(cv::ETA_CF, 0.0 / fabs(h));
        nflag = cv::PREV_CONV_FAIL;
        rescale();
        return cv::PREDICT_AGAIN;
    }

    /// full sensitivity right-hand side at (t, y) incl. a fresh Jacobian (cvSensRhsWrapper with
    /// AMICI's fsxdot): used at initialisation, in cvHin and on the order-1 restart
    PB_HD int sens_rhs_at(double t, const double* y_state_lane, const double (&s)[CPL][NX], double (&out)[CPL][NX]) {
        publish_state(y_state_lane);
        refresh_Jcur(t);
        return f_sens(t, s, out);
    }

    /// cvDoErrorTest (cvodes.c:7461-7614)
    PB_HD int do_error_test(int& nflag, double saved_t, double acor_nrm, int& nef, int& netf_total, double& dsm) {
        dsm = acor_nrm * sh->tq[2];
        if (dsm <= 1.0) return 0;
        nef++;
        netf_total++;
        nflag = cv::PREV_ERR_FAIL;
        restore(saved_t);
        if ((fabs(h) <= 0.0 * cv::ONEPSM) || (nef == cv::MXNEF)) return SIM_ERR_FAILURE;
        etamax = 1.0;
        if (nef <= cv::MXNEF1) {
            eta = 1.0 / (pow(cv::BIAS2 * dsm, 1.0 / L) + cv::ADDON);
            eta = fmax(cv::ETA_MIN_EF, fmax(eta, 0.0 / fabs(h)));
            if (nef >= cv::SMALL_NEF) eta = fmin(eta, cv::ETA_MAX_EF);
            rescale();
            return cv::TRY_AGAIN;
        }
        if (q > 1) {
            eta = fmax(cv::ETA_MIN_EF, 0.0 / fabs(h));
            adjust_order(-1);
            L = q;
            q--;
            qwait = L;
            rescale();
            return cv::TRY_AGAIN;
        }
        // already at order 1: restart, reload zn[1] from scratch
        eta = fmax(cv::ETA_MIN_EF, 0.0 / fabs(h));
        h *= eta;
        hscale = h;
        qwait = cv::LONG_WAIT;
        nscon = 0;
        double ft[NX];
        int r = f_state(tn, zn[0][0], ft);
        if (r) return SIM_UNREC_RHSFUNC_ERR;
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) zn[1][0][i] = h * ft[i];
        }
        if (sensi) {
            double fs[CPL][NX];
            r = sens_rhs_at(tn, zn[0][0], zn[0], fs);
            if (r) return SIM_UNREC_SRHSFUNC_ERR;
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                if (!act[m] || is_state(m)) continue;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[1][m][i] = h * fs[m][i];
            }
        }
        return cv::TRY_AGAIN;
    }

    /// cvCompleteStep (cvodes.c:7634-7736)
    PB_HD void complete_step() {
        nst++;
        nscon++;
        qsum += q;
        g.sync();
        if (g.lane() == 0) {
            for (int i = q; i >= 2; --i) sh->tau[i] = sh->tau[i - 1];
            if ((q == 1) && (nst > 1)) sh->tau[2] = sh->tau[1];
            sh->tau[1] = h;
        }
        g.sync();
        PB_UNROLL
        for (int j = 0; j <= cv::Q_MAX; ++j) {
            if (j <= q) {
                const double lj = sh->l[j];
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) zn[j][m][i] += lj * acor[m][i];
                }
            }
        }
        qwait--;
        if ((qwait == 1) && (q != cv::Q_MAX)) {
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[cv::Q_MAX][m][i] = acor[m][i];
            }
            saved_tq5 = sh->tq[5];
        }
    }

    /// cvSetEta (cvodes.c:7798-7827)
    PB_HD void set_eta() {
        if ((eta > cv::ETA_MIN_FX) && (eta < cv::ETA_MAX_FX)) {
            eta = 1.0;
            hprime = h;
        } else {
            if (eta >= cv::ETA_MAX_FX) {
                eta = fmin(eta, etamax);
                // hmax_inv = 0: no upper bound on the step
            } else {
                eta = fmax(eta, cv::ETA_MIN);
            }
            hprime = h * eta;
            if (qprime < q) nscon = 0;
        }
    }

    /// cvPrepareNextStep with cvComputeEtaqm1 / cvComputeEtaqp1 / cvChooseEta
    /// (cvodes.c:7747-8000)
    PB_HD void prepare_next_step(double dsm) {
        if (etamax == 1.0) {
            qwait = qwait > 2 ? qwait : 2;
            qprime = q;
            hprime = h;
            eta = 1.0;
            return;
        }
        etaq = 1.0 / (pow(cv::BIAS2 * dsm, 1.0 / L) + cv::ADDON);
        if (qwait != 0) {
            eta = etaq;
            qprime = q;
            set_eta();
            return;
        }
        qwait = 2;
        // etaqm1
        etaqm1 = 0.0;
        if (q > 1) {
            double ddn = 0.0;
            PB_UNROLL
            for (int j = 2; j <= cv::Q_MAX; ++j)
                if (j == q) ddn = norm_all(zn[j]);
            ddn = ddn * sh->tq[1];
            etaqm1 = 1.0 / (pow(cv::BIAS1 * ddn, 1.0 / q) + cv::ADDON);
        }
        // etaqp1
        etaqp1 = 0.0;
        if (q != cv::Q_MAX) {
            if (saved_tq5 != 0.0) {
                double hr = h / sh->tau[2], hp = 1.0;
                for (int i = 0; i < L; ++i) hp *= hr;  // SUNRpowerI
                const double cquot = (sh->tq[5] / saved_tq5) * hp;
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) tmp[m][i] = -cquot * zn[cv::Q_MAX][m][i] + acor[m][i];
                }
                double dup = norm_all(tmp);
                dup = dup * sh->tq[3];
                etaqp1 = 1.0 / (pow(cv::BIAS3 * dup, 1.0 / (L + 1)) + cv::ADDON);
            }
        }
        // cvChooseEta
        const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
        if ((etam > cv::ETA_MIN_FX) && (etam < cv::ETA_MAX_FX)) {
            eta = 1.0;
            qprime = q;
        } else if (etam == etaq) {
            eta = etaq;
            qprime = q;
        } else if (etam == etaqm1) {
            eta = etaqm1;
            qprime = q - 1;
        } else {
            eta = etaqp1;
            qprime = q + 1;
            // store Delta_n in zn[qmax] for the order increase (state always; sensitivities
            // only if they take part in the error test, cvodes.c:7973-7987)
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                if (!is_state(m) && !opt.errconS) continue;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[cv::Q_MAX][m][i] = acor[m][i];
            }
        }
        set_eta();
    }

    /// cvSLdet (cvodes.c:8256-8556): dominant characteristic root from the STALD history
    PB_HD int sldet() const {
        int kmin = 0, kflag = 0;
        double rat[5][4], rav[4], qkr[4], sigsq[4], smax[4], ssmax[4];
        double drr[4], rrc[4], sqmx[4], qjk[4][4], vrat[5], qc[6][4], qco[6][4];
        const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
        double rr = 0.0, sqmin = 0.0;
        double ss[6][4];
        for (int i = 1; i <= 5; ++i)
            for (int k = 1; k <= 3; ++k) ss[i][k] = sh->ssdat[i][k];
        for (int k = 1; k <= 3; ++k) {
            double smink = ss[1][k], smaxk = 0.0;
            for (int i = 1; i <= 5; ++i) {
                smink = fmin(smink, ss[i][k]);
                smaxk = fmax(smaxk, ss[i][k]);
            }
            if (smink < cv::TINY * smaxk) return -1;
            smax[k] = smaxk;
            ssmax[k] = smaxk * smaxk;
            double sumrat = 0.0, sumrsq = 0.0;
            for (int i = 1; i <= 4; ++i) {
                rat[i][k] = ss[i][k] / ss[i + 1][k];
                sumrat += rat[i][k];
                sumrsq += rat[i][k] * rat[i][k];
            }
            rav[k] = 0.25 * sumrat;
            vrat[k] = fabs(0.25 * sumrsq - rav[k] * rav[k]);
            qc[5][k] = ss[1][k] * ss[3][k] - ss[2][k] * ss[2][k];
            qc[4][k] = ss[2][k] * ss[3][k] - ss[1][k] * ss[4][k];
            qc[3][k] = 0.0;
            qc[2][k] = ss[2][k] * ss[5][k] - ss[3][k] * ss[4][k];
            qc[1][k] = ss[4][k] * ss[4][k] - ss[3][k] * ss[5][k];
            for (int i = 1; i <= 5; ++i) qco[i][k] = qc[i][k];
        }
        const double vmin = fmin(vrat[1], fmin(vrat[2], vrat[3]));
        const double vmax = fmax(vrat[1], fmax(vrat[2], vrat[3]));
        if (vmin < vrrtol * vrrtol) {
            if (vmax > vrrt2 * vrrt2) return -2;
            rr = (rav[1] + rav[2] + rav[3]) / 3.0;
            double drrmax = 0.0;
            for (int k = 1; k <= 3; ++k) drrmax = fmax(drrmax, fabs(rav[k] - rr));
            if (drrmax > vrrt2) return -3;
            kflag = 1;
        } else {
            if (fabs(qco[1][1]) < cv::TINY * ssmax[1]) return -4;
            double tem = qco[1][2] / qco[1][1];
            for (int i = 2; i <= 5; ++i) qco[i][2] -= tem * qco[i][1];
            qco[1][2] = 0.0;
            tem = qco[1][3] / qco[1][1];
            for (int i = 2; i <= 5; ++i) qco[i][3] -= tem * qco[i][1];
            qco[1][3] = 0.0;
            if (fabs(qco[2][2]) < cv::TINY * ssmax[2]) return -4;
            tem = qco[2][3] / qco[2][2];
            for (int i = 3; i <= 5; ++i) qco[i][3] -= tem * qco[i][2];
            if (fabs(qco[4][3]) < cv::TINY * ssmax[3]) return -4;
            rr = -qco[5][3] / qco[4][3];
            if (rr < cv::TINY || rr > 100.0) return -5;
            for (int k = 1; k <= 3; ++k) qkr[k] = qc[5][k] + rr * (qc[4][k] + rr * rr * (qc[2][k] + rr * qc[1][k]));
            double sqmax = 0.0;
            for (int k = 1; k <= 3; ++k) {
                const double saqk = fabs(qkr[k]) / ssmax[k];
                if (saqk > sqmax) sqmax = saqk;
            }
            if (sqmax < sqtol) {
                kflag = 2;
            } else {
                for (int it = 1; it <= 3; ++it) {
                    for (int k = 1; k <= 3; ++k) {
                        const double qp = qc[4][k] + rr * rr * (3.0 * qc[2][k] + rr * 4.0 * qc[1][k]);
                        drr[k] = 0.0;
                        if (fabs(qp) > cv::TINY * ssmax[k]) drr[k] = -qkr[k] / qp;
                        rrc[k] = rr + drr[k];
                    }
                    for (int k = 1; k <= 3; ++k) {
                        const double s = rrc[k];
                        double sqmaxk = 0.0;
                        for (int j = 1; j <= 3; ++j) {
                            qjk[j][k] = qc[5][j] + s * (qc[4][j] + s * s * (qc[2][j] + s * qc[1][j]));
                            const double saqj = fabs(qjk[j][k]) / ssmax[j];
                            if (saqj > sqmaxk) sqmaxk = saqj;
                        }
                        sqmx[k] = sqmaxk;
                    }
                    sqmin = sqmx[1] + 1.0;
                    for (int k = 1; k <= 3; ++k) {
                        if (sqmx[k] < sqmin) {
                            kmin = k;
                            sqmin = sqmx[k];
                        }
                    }
                    rr = rrc[kmin];
                    if (sqmin < sqtol) {
                        kflag = 3;
                        break;
                    } else {
                        for (int j = 1; j <= 3; ++j) qkr[j] = qjk[j][kmin];
                    }
                }
                if (sqmin > sqtol) return -6;
            }
        }
        for (int k = 1; k <= 3; ++k) {
            const double rsa = ss[1][k];
            const double rsb = ss[2][k] * rr;
            const double rsc = ss[3][k] * rr * rr;
            const double rsd = ss[4][k] * rr * rr * rr;
            const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
            const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c;
            const double rd3a = rd2a - rd2b;
            if (fabs(rd1b) < cv::TINY * smax[k]) return -7;
            const double cest1 = -rd3a / rd1b;
            if (cest1 < cv::TINY || cest1 > 4.0) return -7;
            const double corr1 = (rd2b / cest1) / (rr * rr);
            sigsq[k] = ss[3][k] + corr1;
        }
        if (sigsq[2] < cv::TINY) return -8;
        const double ratp = sigsq[3] / sigsq[2];
        const double ratm = sigsq[1] / sigsq[2];
        const double qfac1 = 0.25 * (q * q - 1.0);
        const double qfac2 = 2.0 / (q - 1.0);
        const double bb = ratp * ratm - 1.0 - qfac1 * ratp;
        const double tem = 1.0 - qfac2 * bb;
        if (fabs(tem) < cv::TINY) return -8;
        const double rrb = 1.0 / tem;
        if (fabs(rrb - rr) > rrtol) return -9;
        if (rr > rrcut) {
            if (kflag == 1) kflag = 4;
            if (kflag == 2) kflag = 5;
            if (kflag == 3) kflag = 6;
        }
        return kflag;
    }

    /// cvBDFStab (cvodes.c:8158-8218)
    PB_HD void bdf_stab() {
        if (q >= 3) {
            double sqm1 = 0.0, sqm2 = 0.0;
            PB_UNROLL
            for (int j = 2; j <= cv::Q_MAX; ++j) {
                if (j == q) sqm1 = norm_state(zn[j]);
                if (j == q - 1) sqm2 = norm_state(zn[j]);
            }
            int factorial = 1;
            for (int i = 1; i <= q - 1; ++i) factorial *= i;
            const double sq = factorial * q * (q + 1) * acnrm / fmax(sh->tq[5], cv::TINY);
            sqm1 *= factorial * q;
            sqm2 *= factorial;
            g.sync();
            if (g.lane() == 0) {
                for (int k = 1; k <= 3; ++k)
                    for (int i = 5; i >= 2; --i) sh->ssdat[i][k] = sh->ssdat[i - 1][k];
                sh->ssdat[1][1] = sqm2 * sqm2;
                sh->ssdat[1][2] = sqm1 * sqm1;
                sh->ssdat[1][3] = sq * sq;
            }
            g.sync();
        }
        if (qprime >= q) {
            if ((q >= 3) && (nscon >= q + 5)) {
                const int ldflag = sldet();
                if (ldflag > 3) {
                    qprime = q - 1;
                    eta = etaqm1;
                    eta = fmin(eta, etamax);
                    hprime = h * eta;
                    nor++;
                }
            }
        } else {
            nscon = 0;
        }
    }

    /// cvStep (cvodes.c:5800-6094)
    PB_HD int step() {
        int ncf = 0, nef = 0, ncfS = 0, nefS = 0;
        double dsm = 0.0;
        if ((nst > 0) && (hprime != h)) adjust_params();
        const double saved_t = tn;
        int nflag = cv::FIRST_CALL;
        for (;;) {
            predict();
            cvset();
            nflag = nls_state(nflag);
            int kflag = handle_nflag(nflag, saved_t, ncf, ncfn);
            if (kflag == cv::PREDICT_AGAIN) continue;
            if (kflag != cv::DO_ERROR_TEST) return kflag;
            int eflag = do_error_test(nflag, saved_t, acnrm, nef, netf, dsm);
            if (eflag == cv::TRY_AGAIN) continue;
            if (eflag != 0) return eflag;
            if (sensi) {
                ncf = nef = 0;
                // f at the converged y (cvodes.c:5947-5955) -- AMICI's sensitivity RHS does
                // not use it, the evaluation is kept for the recoverable-failure semantics
                double ft[NX];
                const int fr = f_state(tn, tmp[0], ft);
                if (fr) {
                    nflag = cv::PREV_CONV_FAIL;
                    // cvodes.c:5951-5955 goes straight back to the prediction without
                    // restoring; the predictor then acts on the predicted array.
                    continue;
                }
                publish_state(tmp[0]);
                refresh_Jcur(tn);
                nflag = nls_sens();
                kflag = handle_nflag(nflag, saved_t, ncfS, ncfnS);
                if (kflag == cv::PREDICT_AGAIN) continue;
                if (kflag != cv::DO_ERROR_TEST) return kflag;
                if (opt.errconS) {
                    double dsmS;
                    eflag = do_error_test(nflag, saved_t, acnrmS, nefS, netfS, dsmS);
                    if (eflag == cv::TRY_AGAIN) continue;
                    if (eflag != 0) return eflag;
                    if (dsmS > dsm) dsm = dsmS;
                }
            }
            break;
        }
        complete_step();
        prepare_next_step(dsm);
        if (opt.stldet) bdf_stab();
        etamax = (nst <= cv::SMALL_NST) ? cv::ETA_MAX_ES : cv::ETA_MAX_GS;
        // acor *= tq[2] (local error estimate) is not used downstream and is skipped
        return 0;
    }

    // ---------------------------------------------------------------- initial step
    /// cvYddNorm (cvodes.c:5676-5773)
    PB_HD int ydd_norm(double hg, double& yddnrm) {
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) tmp[0][i] = hg * zn[1][0][i] + zn[0][0][i];
        }
        const bool withS = sensi && opt.errconS;
        if (withS) {
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                if (!act[m] || is_state(m)) continue;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) tmp[m][i] = hg * zn[1][m][i] + zn[0][m][i];
            }
        }
        double ft[NX];
        int r = f_state(tn + hg, tmp[0], ft);
        if (r) return r;
        double fs[CPL][NX];
        if (withS) {
            r = sens_rhs_at(tn + hg, tmp[0], tmp, fs);
            if (r) return r;
        }
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m]) continue;
            if (is_state(m)) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) b[m][i] = (1.0 / hg) * ft[i] + (-1.0 / hg) * zn[1][m][i];
            } else if (withS) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) b[m][i] = (1.0 / hg) * fs[m][i] + (-1.0 / hg) * zn[1][m][i];
            } else {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) b[m][i] = 0.0;
            }
        }
        yddnrm = norm_all(b);
        return 0;
    }

    /// cvUpperBoundH0 (cvodes.c:5558-5665)
    PB_HD double upper_bound_h0(double tdist) const {
        double hub_inv = 0.0;
        const bool withS = sensi && opt.errconS;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m]) continue;
            if (!is_state(m) && !withS) continue;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const double den = cv::HUB_FACTOR * fabs(zn[0][m][i]) + 1.0 / ewt[m][i];
                const double v = fabs(zn[1][m][i]) / den;
                if (v > hub_inv) hub_inv = v;
            }
        }
        hub_inv = g.max(hub_inv);
        double hub = cv::HUB_FACTOR * tdist;
        if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
        return hub;
    }

    /// cvHin (cvodes.c:5436-5549)
    PB_HD int hin(double tout) {
        const double tdiff = tout - tn;
        if (tdiff == 0.0) return SIM_TOO_CLOSE;
        const int sign = (tdiff > 0.0) ? 1 : -1;
        const double tdist = fabs(tdiff);
        const double tround = cv::UROUND * fmax(fabs(tn), fabs(tout));
        if (tdist < 2.0 * tround) return SIM_TOO_CLOSE;
        const double hlb = cv::HLB_FACTOR * tround;
        const double hub = upper_bound_h0(tdist);
        double hg = sqrt(hlb * hub);
        if (hub < hlb) {
            h = (sign == -1) ? -hg : hg;
            return 0;
        }
        double hs = hg, hnew = hg, yddnrm = 0.0;
        for (int count1 = 1; count1 <= cv::HIN_MAX_ITERS; ++count1) {
            bool hgOK = false;
            int retval = 0;
            for (int count2 = 1; count2 <= cv::HIN_MAX_ITERS; ++count2) {
                const double hgs = hg * sign;
                retval = ydd_norm(hgs, yddnrm);
                if (retval < 0) return retval;
                if (retval == 0) { hgOK = true; break; }
                hg *= 0.2;
            }
            if (!hgOK) {
                if (count1 <= 2) return retval == cv::RHSFUNC_RECVR ? SIM_REPTD_RHSFUNC_ERR : SIM_REPTD_SRHSFUNC_ERR;
                hnew = hs;
                break;
            }
            hs = hg;
            hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
            if (count1 == cv::HIN_MAX_ITERS) break;
            const double hrat = hnew / hg;
            if ((hrat > 0.5) && (hrat < 2.0)) break;
            if ((count1 > 1) && (hrat > 2.0)) { hnew = hg; break; }
            hg = hnew;
        }
        double h0 = cv::H_BIAS * hnew;
        if (h0 < hlb) h0 = hlb;
        if (h0 > hub) h0 = hub;
        if (sign == -1) h0 = -h0;
        h = h0;
        return 0;
    }

    // ---------------------------------------------------------------- CVode driver
    /// CVode(tout) in CV_NORMAL mode with tstop = tout, as AMICI drives it
    /// (cvodes.c:2890-3636; deps/AMICI/src/solver.cpp:108-161).  Returns TSTOP_RETURN
    /// (tn == tout on exit) or a negative failure code.
    PB_HD int cvode(double tout) {
        // Solver::run -> setStopTime(tout); apply_max_num_steps()
        tstop = tout;
        tstopset = true;
        if (opt.maxsteps <= nst) return SIM_ERROR;
        const int mxstep = opt.maxsteps - nst;
        if (nst == 0) {
            if (!ewt_set()) return SIM_ILL_INPUT;
            double ft[NX];
            int r = f_state(tn, zn[0][0], ft);
            if (r) return SIM_FIRST_RHSFUNC_ERR;
            if (g.lane() == 0) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[1][0][i] = ft[i];
            }
            if (sensi) {
                double fs[CPL][NX];
                r = sens_rhs_at(tn, zn[0][0], zn[0], fs);
                if (r) return SIM_FIRST_SRHSFUNC_ERR;
                PB_UNROLL
                for (int m = 0; m < CPL; ++m) {
                    if (!act[m] || is_state(m)) continue;
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) zn[1][m][i] = fs[m][i];
                }
            }
            if ((tstop - tn) * (tout - tn) <= 0.0) return SIM_ILL_INPUT;
            const int hr = hin(tout);
            if (hr != 0) return hr;
            if ((tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
            hscale = h;
            hprime = h;
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[1][m][i] *= h;
            }
        } else {
            const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
            if (fabs(tn - tstop) <= troundoff) {
                tstopset = false;
                tn = tstop;  // AMICI reads t_ = tret = tstop
                return cv::TSTOP_RETURN;
            } else if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                eta = hprime / h;
            }
        }
        int nstloc = 0;
        for (;;) {
            if (nst > 0) {
                if (!ewt_set()) return SIM_ILL_INPUT;
            }
            if (nstloc >= mxstep) return SIM_TOO_MUCH_WORK;
            const double nrm = norm_all(zn[0]);
            if (cv::UROUND * nrm > 1.0) return SIM_TOO_MUCH_ACC;
            const int kflag = step();
            if (kflag != 0) return kflag;
            nstloc++;
            const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
            if (fabs(tn - tstop) <= troundoff) {
                tn = tstop;
                tstopset = false;
                return cv::TSTOP_RETURN;
            } else if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                eta = hprime / h;
            }
            if ((tn - tout) * h >= 0.0) return cv::TSTOP_RETURN;
        }
    }

    // ---------------------------------------------------------------- likelihood
    /// One output time: llh and sllh contributions (Model::addObservableObjective /
    /// addObservableObjectiveSensitivity, deps/AMICI/src/model.cpp:1106-1152 with
    /// fsigmay :2087-2119, fdJydy/fdJydsigma/fdJydp/fdJydx :2184-2392; Jy formulas
    /// deps/AMICI/models/model_steadystate/{Jy,dJydy,dJydsigma}.cpp).
    /// Returns false if a sigma is not positive (Model::checkSigmaPositivity throws).
    PB_HD bool data_point(double t, const double* my, const double* sd, double& llh, double (&sllh)[CPL],
                          double* y_out, double* sigmay_out) {
        publish_state(zn[0][0]);
        double y[NY], sig[NY];
        M::y(y, t, sh->ysh, sh->p, sh->k);
        M::sigmay(sig, t, sh->p, sh->k);
        bool from_data[NY];
        bool ok = true;
        PB_UNROLL
        for (int iy = 0; iy < NY; ++iy) {
            from_data[iy] = !isnan(sd[iy]);
            if (from_data[iy]) sig[iy] = sd[iy];
            if (!isnan(my[iy])) ok = ok && (sig[iy] > 0.0) && isfinite(sig[iy]);
        }
        if (g.lane() == 0) {
            if (y_out) {
                PB_UNROLL
                for (int iy = 0; iy < NY; ++iy) y_out[iy] = y[iy];
            }
            if (sigmay_out) {
                PB_UNROLL
                for (int iy = 0; iy < NY; ++iy) sigmay_out[iy] = sig[iy];
            }
        }
        if (!ok) return false;
        PB_UNROLL
        for (int iy = 0; iy < NY; ++iy) {
            if (isnan(my[iy])) continue;
            const double s2 = sig[iy] * sig[iy];
            const double r = my[iy] - y[iy];
            llh -= log(s2 * 3.141592653589793 * 2.0) * 0.5 + 1.0 / s2 * (r * r) * 0.5;
        }
        if (!sensi) return true;
        double dJydy[NY], dJyds[NY];
        PB_UNROLL
        for (int iy = 0; iy < NY; ++iy) {
            if (isnan(my[iy])) {
                dJydy[iy] = 0.0;
                dJyds[iy] = 0.0;
                continue;
            }
            const double s2 = sig[iy] * sig[iy];
            const double r = my[iy] - y[iy];
            dJydy[iy] = 1.0 / s2 * (my[iy] * 2.0 - y[iy] * 2.0) * -0.5;
            dJyds[iy] = 1.0 / (s2 * sig[iy]) * (r * r) * -1.0 + 1.0 / sig[iy];
            if (from_data[iy]) dJyds[iy] = 0.0;  // dsigmaydp := 0 for data sigmas (model.cpp:2137-2145)
        }
        // dJydx = dJydy * dydx
        double dJydx[NX];
        {
            double dydx[NY * NX];
            PB_UNROLL
            for (int i = 0; i < NY * NX; ++i) dydx[i] = 0.0;
            M::dydx(dydx, t, sh->ysh, sh->p, sh->k);
            PB_UNROLL
            for (int ix = 0; ix < NX; ++ix) {
                double a = 0.0;
                PB_UNROLL
                for (int iy = 0; iy < NY; ++iy) a += dJydy[iy] * dydx[iy + ix * NY];
                dJydx[ix] = a;
            }
        }
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            if (!act[m] || is_state(m)) continue;
            double dydp[NY], dsdp[NY];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) { dydp[iy] = 0.0; dsdp[iy] = 0.0; }
            M::dydp(dydp, t, sh->ysh, sh->p, sh->k, ipar[m]);
            M::dsigmaydp(dsdp, t, sh->p, sh->k, ipar[m]);
            double dJydp = 0.0;
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJydy[iy] * dydp[iy];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJyds[iy] * dsdp[iy];
            double a = dJydp;
            PB_UNROLL
            for (int ix = 0; ix < NX; ++ix) a += dJydx[ix] * zn[0][m][ix];
            sllh[m] -= a;
        }
        return true;
    }

    // ---------------------------------------------------------------- one simulation
    /// ForwardProblem::workForwardProblem + ReturnData for one condition at the given
    /// tolerances.  sh->p / sh->k / sh->plist must be set.  Returns the AMICI status.
    PB_HD int simulate(const ConditionIO& io, double rtol, double atol, double& llh_out, double (&sllh_out)[CPL]) {
        opt.rtol = rtol;
        opt.atol = atol;
        ns = io.ns;
        sensi = ns > 0;
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            const int c = g.lane() + m * G;
            act[m] = c <= ns;
            ipar[m] = (c >= 1 && c <= ns) ? sh->plist[c - 1] : -1;
        }
        // Model::initialize: x0, sx0 (deps/AMICI/src/model.cpp:249-275)
        PB_UNROLL
        for (int j = 0; j < cv::L_MAX; ++j) {
            PB_UNROLL
            for (int m = 0; m < CPL; ++m) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) zn[j][m][i] = 0.0;
            }
        }
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) { acor[m][i] = 0.0; ewt[m][i] = 1.0; tmp[m][i] = 0.0; b[m][i] = 0.0; }
            if (!act[m]) continue;
            if (is_state(m)) M::x0(zn[0][m], opt.t0, sh->p, sh->k);
            else M::sx0(zn[0][m], opt.t0, sh->p, sh->k, ipar[m]);
        }
        // CVodeInit (cvodes.c:677-830)
        tn = opt.t0;
        q = 1; L = 2; qwait = L; qprime = 1;
        etamax = cv::ETA_MAX_FS;
        h = hprime = hscale = 0.0;
        eta = 1.0; etaq = etaqm1 = etaqp1 = 0.0;
        gamma = gammap = 0.0; gamrat = 1.0; rl1 = 1.0;
        crate = crateS = 1.0; delp = 0.0; acnrm = acnrmS = 0.0; saved_tq5 = 0.0;
        nst = nscon = nstlp = nstlj = 0;
        convfail = cv::NO_FAILURES;
        jcur = jcurNLS = jcurNLSS = false;
        tstopset = false; tstop = 0.0;
        nfe = nsetups = nje = nni = nniS = nfSe = netf = netfS = ncfn = ncfnS = nor = qsum = 0;
        g.sync();
        if (g.lane() == 0) {
            for (int i = 0; i <= cv::L_MAX; ++i) { sh->l[i] = 0.0; sh->tau[i] = 0.0; }
            for (int i = 0; i < 6; ++i) sh->tq[i] = 0.0;
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 4; ++k) sh->ssdat[i][k] = 0.0;
        }
        g.sync();

        double llh = 0.0;
        double sllh[CPL];
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) sllh[m] = 0.0;
        int status = SIM_SUCCESS;
        for (int it = 0; it < io.nt; ++it) {
            const double tout = io.t[it];
            if (isinf(tout)) break;  // post-equilibration time points are out of scope
            if (tout > opt.t0) {
                while (tn < tout) {
                    // is_next_t_too_close, forwardproblem.cpp:23-35
                    const double tdiff = tout - tn;
                    if (tdiff == 0.0) break;
                    if (fabs(tdiff) < 2.0 * cv::UROUND * fmax(fabs(tn), fabs(tout))) break;
                    const int r = cvode(tout);
                    if (r < 0) { status = r; break; }
                }
                if (status != SIM_SUCCESS) break;
            }
            double* yo = io.y_out ? io.y_out + (size_t)it * NY : nullptr;
            double* so = io.sigmay_out ? io.sigmay_out + (size_t)it * NY : nullptr;
#if defined(PB200_HOST_TRACE)
            if (io.trace) { io.trace[2 * it] = nst; io.trace[2 * it + 1] = q; }
#endif
            if (!data_point(tout, io.my + (size_t)it * NY, io.sd + (size_t)it * NY, llh, sllh, yo, so)) {
                status = SIM_ERROR;
                break;
            }
        }
        if (status != SIM_SUCCESS) {
            llh = NAN;  // ReturnData::invalidate (rdata.cpp)
        } else if (sensi) {
            // applyChainRuleFactorToSimulationResults: caller multiplies by the
            // parameter-scale factor (it owns pscale); nothing to do here.
        }
        llh_out = g.bcast0(llh);
        PB_UNROLL
        for (int m = 0; m < CPL; ++m) sllh_out[m] = sllh[m];
        return status;
    }

    PB_HD void export_stats(int64_t* s, int trials) const {
        s[STAT_NST] = nst; s[STAT_NFE] = nfe; s[STAT_NSETUPS] = nsetups; s[STAT_NJE] = nje;
        s[STAT_NNI] = nni; s[STAT_NNIS] = nniS; s[STAT_NFSE] = nfSe; s[STAT_NETF] = netf;
        s[STAT_NETFS] = netfS; s[STAT_NCFN] = ncfn; s[STAT_NCFNS] = ncfnS; s[STAT_NOR] = nor;
        s[STAT_TRIALS] = trials; s[STAT_QSUM] = qsum; s[14] = 0; s[15] = 0;
    }
};

}  // namespace pb200
#endif
