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
//     processed by lane (c mod G);
//   * the Nordsieck history zn[0..5], the accumulated correction and the error
//     weights of all columns live in the group's slice of shared memory as
//     [order][state index][column] with the column index fastest, so the lanes of
//     a group touch consecutive 8-byte words (conflict-free) and every history
//     operation is a short rolled loop: the first version of this kernel kept
//     the columns in registers with fully unrolled code and was instruction-fetch
//     bound (profiles/r01_a_summary.txt);
//   * norms are warp-shuffle reductions (max over the lanes of a group);
//   * the Jacobian, its LU factors and the group-uniform BDF coefficient arrays
//     live once per group in shared memory; the state lane factorises, every
//     lane back-substitutes its own columns against the same factors (the
//     staggered corrector's "same matrix, nplist right-hand sides");
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
    PB_HD bool any(bool p) const { return p; }
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
    __device__ __forceinline__ bool any(bool p) const { return (__ballot_sync(mask, p) & mask) != 0u; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};
#endif

/// Per-condition working set: one copy per lane group (shared memory on the device).
template <class M>
struct GroupShared {
    static constexpr int NX = M::NX;
    static constexpr int NC = M::NP + 1;  // columns: state + one per model parameter
    // column data, [..][state index][column], column fastest
    // element (state index i, column c) of a vector lives at [i * ncol + c], ncol = 1 + number of
    // sensitivity columns of the current condition: the active elements are contiguous, so the
    // element-wise history operations run over a flat index with all lanes busy
    double zn[cv::L_MAX][NX * NC];  // Nordsieck history
    double acor[NX * NC];           // accumulated corrector update
    double ewt[NX * NC];            // error weights
    double b[NX * NC];              // residual / Newton update
    double sllh[NC];               // d llh / d p accumulator per sensitivity column
    // group-uniform data
    double Jsaved[NX * NX];  // Jacobian at its last evaluation (cvLs savedJ)
    double LU[NX * NX];      // factors of I - gamma*J
    double Jcur[NX * NX];    // J(tn, y_n) for the sensitivity right-hand sides
    double dfdp[NX * (M::NP > 0 ? M::NP : 1)];  // d xdot / d p at (tn, y_n), column ip at [ip*NX]
    double ysh[NX];          // converged state, read by the sensitivity lanes
    double yj[NX];           // state handed to the out-of-line model functions (xdot, Jacobian)
    double fout[NX];         // xdot returned by eval_xdot_shared
    double p[M::NP > 0 ? M::NP : 1];  // unscaled model parameters
    double k[M::NK > 0 ? M::NK : 1];  // constants (fixed parameters)
    double l[cv::L_MAX + 1];
    double tau[cv::L_MAX + 1];
    double tq[6];
    double ssdat[6][4];      // STALD history (cvBDFStab)
    int piv[NX];
    int perm[NX];            // row order after all pivot swaps: (P b)[i] = b[perm[i]]
    int plist[M::NP > 0 ? M::NP : 1];
};

// ---- out-of-line helpers (one copy of the code each; they touch only group-shared
//      memory and scalars, so a real call costs a few registers) ------------------------
PB_HD_NOINLINE static double pb_pow(double x, double y) { return pow(x, y); }
PB_HD_NOINLINE static double pb_div(double a, double b) { return a / b; }
PB_HD_NOINLINE static double pb_rcp(double a) { return 1.0 / a; }

/// xdot(t, sh->yj) -> sh->fout; returns false if a component is not finite (AMICI checkFinite,
/// deps/AMICI/src/solver_cvodes.cpp:1149-1173)
template <class M>
PB_HD_NOINLINE static bool eval_xdot_shared(GroupShared<M>* sh, double t) {
    M::xdot(sh->fout, t, sh->yj, sh->p, sh->k);
    bool ok = true;
    PB_ROLL
    for (int i = 0; i < M::NX; ++i) ok = ok && isfinite(sh->fout[i]);
    return ok;
}

template <class Grp>
PB_HD_NOINLINE static double group_max(const Grp& g, double v) { return g.max(v); }

/// max over the lanes of a group of a local sum of squares, then sqrt(. / n): the WRMS norm
/// max_columns sqrt(mean((v*w)^2)) (N_VWrmsNorm; cvSensNorm cvodes.c:9427-9441)
template <class Grp>
PB_HD_NOINLINE static double group_wrms(const Grp& g, double ssq_local, int n) {
    return sqrt(pb_div(g.max(ssq_local), (double)n));
}

/// J(t, y) -> the structurally non-zero entries of Jdst (dense column-major, NX*NX), everything
/// in group-shared memory; the structural zeros are written once per condition (simulate()).
template <class M>
PB_HD_NOINLINE static void eval_jac_shared(double* Jdst, double t, const double* y, const GroupShared<M>* sh) {
    M::jac(Jdst, t, y, sh->p, sh->k);
}

/// d xdot / d p (all columns) at (t, y) -> sh->dfdp, structural non-zeros only
template <class M>
PB_HD_NOINLINE static void eval_dfdp_shared(GroupShared<M>* sh, double t, const double* y) {
    M::dxdotdp_dense(sh->dfdp, t, y, sh->p, sh->k);
}

/// LU = I - gamma*Jsaved, factorised in place with partial pivoting, column-oriented
/// (SUNDlsMat_denseGETRF, sundials_dense.c:82-148); the pivots are stored as reciprocals.
/// Returns 0, or 1 on a zero pivot.
template <class M>
PB_HD_NOINLINE static int lu_factor_shared(GroupShared<M>* sh, double gamma) {
    constexpr int NX = M::NX;
    PB_ROLL
    for (int c = 0; c < NX; ++c)
        PB_ROLL
        for (int r = 0; r < NX; ++r) sh->LU[r + c * NX] = (r == c ? 1.0 : 0.0) - gamma * sh->Jsaved[r + c * NX];
    PB_ROLL
    for (int k = 0; k < NX; ++k) {
        double* col_k = sh->LU + k * NX;
        int l = k;
        PB_ROLL
        for (int i = k + 1; i < NX; ++i)
            if (fabs(col_k[i]) > fabs(col_k[l])) l = i;
        sh->piv[k] = l;
        if (col_k[l] == 0.0) return 1;
        if (l != k) {
            PB_ROLL
            for (int i = 0; i < NX; ++i) {
                const double t_ = sh->LU[l + i * NX];
                sh->LU[l + i * NX] = sh->LU[k + i * NX];
                sh->LU[k + i * NX] = t_;
            }
        }
        const double mult = pb_rcp(col_k[k]);
        PB_ROLL
        for (int i = k + 1; i < NX; ++i) col_k[i] *= mult;
        PB_ROLL
        for (int j = k + 1; j < NX; ++j) {
            double* col_j = sh->LU + j * NX;
            const double a_kj = col_j[k];
            if (a_kj != 0.0)
                PB_ROLL
                for (int i = k + 1; i < NX; ++i) col_j[i] -= a_kj * col_k[i];
        }
        col_k[k] = mult;
    }
    PB_ROLL
    for (int i = 0; i < NX; ++i) sh->perm[i] = i;
    PB_ROLL
    for (int k = 0; k < NX; ++k) {
        const int t_ = sh->perm[k];
        sh->perm[k] = sh->perm[sh->piv[k]];
        sh->perm[sh->piv[k]] = t_;
    }
    return 0;
}

/// cvSLdet (cvodes.c:8256-8556): dominant characteristic root from the STALD history (out of line)
PB_HD_NOINLINE static int sldet_shared(const double (*ssdat)[4], int q) {
    int kmin = 0, kflag = 0;
    double rat[5][4], rav[4], qkr[4], sigsq[4], smax[4], ssmax[4];
    double drr[4], rrc[4], sqmx[4], qjk[4][4], vrat[5], qc[6][4], qco[6][4];
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
    double rr = 0.0, sqmin = 0.0;
    double ss[6][4];
    PB_ROLL
    for (int i = 1; i <= 5; ++i)
        PB_ROLL
        for (int k = 1; k <= 3; ++k) ss[i][k] = ssdat[i][k];
    PB_ROLL
    for (int k = 1; k <= 3; ++k) {
        double smink = ss[1][k], smaxk = 0.0;
        PB_ROLL
        for (int i = 1; i <= 5; ++i) {
            smink = fmin(smink, ss[i][k]);
            smaxk = fmax(smaxk, ss[i][k]);
        }
        if (smink < cv::TINY * smaxk) return -1;
        smax[k] = smaxk;
        ssmax[k] = smaxk * smaxk;
        double sumrat = 0.0, sumrsq = 0.0;
        PB_ROLL
        for (int i = 1; i <= 4; ++i) {
            rat[i][k] = pb_div(ss[i][k], ss[i + 1][k]);
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
        PB_ROLL
        for (int i = 1; i <= 5; ++i) qco[i][k] = qc[i][k];
    }
    const double vmin = fmin(vrat[1], fmin(vrat[2], vrat[3]));
    const double vmax = fmax(vrat[1], fmax(vrat[2], vrat[3]));
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        rr = pb_div(rav[1] + rav[2] + rav[3], 3.0);
        double drrmax = 0.0;
        PB_ROLL
        for (int k = 1; k <= 3; ++k) drrmax = fmax(drrmax, fabs(rav[k] - rr));
        if (drrmax > vrrt2) return -3;
        kflag = 1;
    } else {
        if (fabs(qco[1][1]) < cv::TINY * ssmax[1]) return -4;
        double tem = pb_div(qco[1][2], qco[1][1]);
        PB_ROLL
        for (int i = 2; i <= 5; ++i) qco[i][2] -= tem * qco[i][1];
        qco[1][2] = 0.0;
        tem = pb_div(qco[1][3], qco[1][1]);
        PB_ROLL
        for (int i = 2; i <= 5; ++i) qco[i][3] -= tem * qco[i][1];
        qco[1][3] = 0.0;
        if (fabs(qco[2][2]) < cv::TINY * ssmax[2]) return -4;
        tem = pb_div(qco[2][3], qco[2][2]);
        PB_ROLL
        for (int i = 3; i <= 5; ++i) qco[i][3] -= tem * qco[i][2];
        if (fabs(qco[4][3]) < cv::TINY * ssmax[3]) return -4;
        rr = pb_div(-qco[5][3], qco[4][3]);
        if (rr < cv::TINY || rr > 100.0) return -5;
        PB_ROLL
        for (int k = 1; k <= 3; ++k) qkr[k] = qc[5][k] + rr * (qc[4][k] + rr * rr * (qc[2][k] + rr * qc[1][k]));
        double sqmax = 0.0;
        PB_ROLL
        for (int k = 1; k <= 3; ++k) {
            const double saqk = pb_div(fabs(qkr[k]), ssmax[k]);
            if (saqk > sqmax) sqmax = saqk;
        }
        if (sqmax < sqtol) {
            kflag = 2;
        } else {
            PB_ROLL
            for (int it = 1; it <= 3; ++it) {
                PB_ROLL
                for (int k = 1; k <= 3; ++k) {
                    const double qp = qc[4][k] + rr * rr * (3.0 * qc[2][k] + rr * 4.0 * qc[1][k]);
                    drr[k] = 0.0;
                    if (fabs(qp) > cv::TINY * ssmax[k]) drr[k] = pb_div(-qkr[k], qp);
                    rrc[k] = rr + drr[k];
                }
                PB_ROLL
                for (int k = 1; k <= 3; ++k) {
                    const double s = rrc[k];
                    double sqmaxk = 0.0;
                    PB_ROLL
                    for (int j = 1; j <= 3; ++j) {
                        qjk[j][k] = qc[5][j] + s * (qc[4][j] + s * s * (qc[2][j] + s * qc[1][j]));
                        const double saqj = pb_div(fabs(qjk[j][k]), ssmax[j]);
                        if (saqj > sqmaxk) sqmaxk = saqj;
                    }
                    sqmx[k] = sqmaxk;
                }
                sqmin = sqmx[1] + 1.0;
                PB_ROLL
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
                    PB_ROLL
                    for (int j = 1; j <= 3; ++j) qkr[j] = qjk[j][kmin];
                }
            }
            if (sqmin > sqtol) return -6;
        }
    }
    PB_ROLL
    for (int k = 1; k <= 3; ++k) {
        const double rsa = ss[1][k];
        const double rsb = ss[2][k] * rr;
        const double rsc = ss[3][k] * rr * rr;
        const double rsd = ss[4][k] * rr * rr * rr;
        const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
        const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c;
        const double rd3a = rd2a - rd2b;
        if (fabs(rd1b) < cv::TINY * smax[k]) return -7;
        const double cest1 = pb_div(-rd3a, rd1b);
        if (cest1 < cv::TINY || cest1 > 4.0) return -7;
        const double corr1 = pb_div(pb_div(rd2b, cest1), rr * rr);
        sigsq[k] = ss[3][k] + corr1;
    }
    if (sigsq[2] < cv::TINY) return -8;
    const double ratp = pb_div(sigsq[3], sigsq[2]);
    const double ratm = pb_div(sigsq[1], sigsq[2]);
    const double qfac1 = 0.25 * (q * q - 1.0);
    const double qfac2 = pb_div(2.0, q - 1.0);
    const double bb = ratp * ratm - 1.0 - qfac1 * ratp;
    const double tem = 1.0 - qfac2 * bb;
    if (fabs(tem) < cv::TINY) return -8;
    const double rrb = pb_rcp(tem);
    if (fabs(rrb - rr) > rrtol) return -9;
    if (rr > rrcut) {
        if (kflag == 1) kflag = 4;
        if (kflag == 2) kflag = 5;
        if (kflag == 3) kflag = 6;
    }
    return kflag;
}



/// WRMS norm max_c sqrt(mean_i(((ca*A[i][c] + cb*B[i][c]) * ewt[i][c])^2)) over the columns
/// [cmin, cmax] of [NX][NC] arrays in group-shared memory: N_VWrmsNorm for the state column,
/// cvSensNorm for the sensitivities, cvSensUpdateNorm for both (cvodes.c:9427-9457).  One copy
/// of the code serves every call site.
template <class M, class Grp>
PB_HD_NOINLINE static double wnorm_shared(const GroupShared<M>* sh, const Grp& g, int NC, int cmin, int cmax,
                                          const double* A, double ca, const double* B, double cb) {
    constexpr int NX = M::NX;
    double loc = 0.0;
    for (int c = g.lane(); c <= cmax; c += Grp::G) {
        if (c < cmin) continue;
        double s = 0.0;
        PB_UNROLL_X
        for (int i = 0; i < NX; ++i) {
            const double prod = (ca * A[i * NC + c] + cb * B[i * NC + c]) * sh->ewt[i * NC + c];
            s += prod * prod;
        }
        if (s > loc) loc = s;
    }
    return sqrt(pb_div(g.max(loc), (double)NX));
}

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

template <class M, class Grp>
struct Bdf {
    static constexpr int NX = M::NX, NP = M::NP, NK = M::NK, NY = M::NY;
    static constexpr int G = Grp::G;
    enum { PH_STATE = 0, PH_SENS = 1 };

    Grp g;
    GroupShared<M>* sh;
    SolverOptions opt;
    int ns;      // sensitivity columns (columns 1..ns; column 0 is the state)
    int ncol;    // 1 + ns: row stride of the column data in shared memory
    bool sensi;  // ns > 0

    // integrator scalars (uniform within the group)
    double tn, h, hprime, hscale, eta, etamax, etaq, etaqm1, etaqp1;
    double gamma, gammap, gamrat, rl1, crate, crateS, delp, acnrm, acnrmS, saved_tq5, tstop;
    int q, qprime, L, qwait, nst, nscon, nstlp, nstlj, convfail;
    bool jcur, jcurNLS, jcurNLSS, tstopset;
    // counters
    int nfe, nsetups, nje, nni, nniS, nfSe, netf, netfS, ncfn, ncfnS, nor, qsum;
    // re-entrant driver state (begin()/advance(): one step attempt per call, so that the groups of
    // a warp -- and the warps of a CTA -- can be kept in lock-step per attempt)
    ConditionIO io_;
    double llh_, tout_, saved_t_, dsm_, dsmS_;
    int it_, sim_status_, nstloc_, mxstep_, ncf_, nef_, ncfS_, nefS_, nflag_, dq_;
    bool in_cvode_, in_step_, do_rescale_, undo_, restart_;

// columns of this lane in [cmin, cmax]
#define PB_FOR_COLS(c, cmin, cmax) \
    for (int c = g.lane(); c <= (cmax); c += G) \
        if (c >= (cmin))
#define PB_FOR_ALL_COLS(c) for (int c = g.lane(); c <= ns; c += G)
// all (state index, column) elements of the active columns, flat: every lane busy
#define PB_FOR_ELEMS(e) for (int e = g.lane(); e < NX * ncol; e += G)

    PB_HD int phase_cmin(int phase) const { return phase == PH_STATE ? 0 : 1; }
    PB_HD int phase_cmax(int phase) const { return phase == PH_STATE ? 0 : ns; }

    // ---------------------------------------------------------------- norms
    PB_HD double wnorm1(int cmin, int cmax, const double* A) const {
        return wnorm_shared<M, Grp>(sh, g, ncol, cmin, cmax, A, 1.0, A, 0.0);
    }
    PB_HD double wnorm2(int cmin, int cmax, const double* A, double ca, const double* B, double cb) const {
        return wnorm_shared<M, Grp>(sh, g, ncol, cmin, cmax, A, ca, B, cb);
    }
    /// columns in the error test: state, plus sensitivities if errconS
    PB_HD int errtest_cmax() const { return (sensi && opt.errconS) ? ns : 0; }

    /// cvEwtSetSS / cvSensEwtSetSS (cvodes.c:9118-9129, 9263-9279)
    PB_HD bool ewt_set() {
        bool ok = true;
        PB_FOR_ELEMS(e) {
            const double w = opt.rtol * fabs(sh->zn[0][e]) + opt.atol;
            ok = ok && (w > 0.0);
            sh->ewt[e] = pb_rcp(w);
        }
        return !g.any(!ok);
    }

    /// J(t, ysh) -> sh->Jcur by the state lane (Model_ODE::fsxdot at ip == 0,
    /// deps/AMICI/src/model_ode.cpp:492-498)
    PB_HD void refresh_Jcur(double t) {
        g.sync();
        if (g.lane() == 0) {
            eval_jac_shared<M>(sh->Jcur, t, sh->ysh, sh);
            eval_dfdp_shared<M>(sh, t, sh->ysh);
        }
        g.sync();
    }

    // ---------------------------------------------------------------- right-hand sides
    /// THE right-hand-side evaluation (single copy in the instruction stream).  For the columns
    /// of `phase`:  v = zn[0] + acor;  F = f(t, v)  (state: xdot; sensitivities:
    /// dxdotdp(ip) + Jcur*v, Model_ODE::fsxdot model_ode.cpp:486-515, with sh->ysh / sh->Jcur
    /// prepared by the caller);  b = raw ? F : rl1*zn[1] + acor - gamma*F  (cvNlsResidual
    /// cvodes_nls.c:344-371; cvNlsResidualSensStg cvodes_nls_stg.c:377-421).
    /// Returns 0, RHSFUNC_RECVR or SRHSFUNC_RECVR (non-finite result).
    PB_HD int rhs(int phase, double t, bool raw) {
        bool bad = false;
        PB_FOR_COLS(c, phase_cmin(phase), phase_cmax(phase)) {
            double F[NX];
            if (c == 0) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) sh->yj[i] = sh->zn[0][(i) * ncol + (0)] + sh->acor[(i) * ncol + (0)];
                if (!eval_xdot_shared<M>(sh, t)) bad = true;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) F[i] = sh->fout[i];
            } else {
                const double* dcol = sh->dfdp + sh->plist[c - 1] * NX;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) F[i] = dcol[i];
                PB_ROLL
                for (int j = 0; j < NX; ++j) {
                    const double sj = sh->zn[0][(j) * ncol + (c)] + sh->acor[(j) * ncol + (c)];
                    PB_UNROLL
                    for (int r = 0; r < NX; ++r) F[r] += sh->Jcur[r + j * NX] * sj;
                }
                bool ok = true;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) ok = ok && isfinite(F[i]);
                if (!ok) bad = true;
            }
            PB_UNROLL
            for (int i = 0; i < NX; ++i)
                sh->b[(i) * ncol + (c)] = raw ? F[i] : (rl1 * sh->zn[1][(i) * ncol + (c)] + sh->acor[(i) * ncol + (c)]) - gamma * F[i];
        }
        if (phase == PH_STATE) nfe++;
        else nfSe += ns;
        if (g.any(bad)) return phase == PH_STATE ? cv::RHSFUNC_RECVR : cv::SRHSFUNC_RECVR;
        return 0;
    }

    /// state lane: sh->ysh = zn[0] + acor of the state column (the corrected state y_n)
    PB_HD void publish_state() {
        g.sync();
        if (g.lane() == 0) {
            PB_UNROLL_X
            for (int i = 0; i < NX; ++i) sh->ysh[i] = sh->zn[0][(i) * ncol + (0)] + sh->acor[(i) * ncol + (0)];
        }
        g.sync();
    }

    /// b = f(t, zn[0] + acor) for the state, then -- if sensitivities take part -- the
    /// sensitivity right-hand sides at that state with a fresh Jacobian (cvSensRhsWrapper with
    /// AMICI's fsxdot).  Used by the cold paths: initialisation, cvHin, order-1 restart.
    PB_HD int rhs_all_raw(double t, bool with_sens) {
        int r = 0;
        for (int phase_ = PH_STATE; phase_ <= PH_SENS; ++phase_) {
            int phase = phase_;
            PB_OPAQUE_INT(phase);
            if (phase == PH_SENS) {
                if (!with_sens) break;
                publish_state();
                refresh_Jcur(t);
            }
            r = rhs(phase, t, true);
            if (r) break;
        }
        return r;
    }

    // ---------------------------------------------------------------- linear algebra
    /// cvLsSetup + cvLsLinSys + dense GETRF (cvodes_ls.c:1611-1708,1386-1470); the state the
    /// Jacobian is evaluated at is sh->yj.
    PB_HD int lsetup() {
        const double dgamma = fabs(pb_div(gamma, gammap) - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + cv::MSBJ) ||
                          ((convfail == cv::FAIL_BAD_J) && (dgamma < cv::DGMAX_JBAD)) ||
                          (convfail == cv::FAIL_OTHER);
        int ret = 0;
        g.sync();
        if (jbad) {
            if (g.lane() == 0) eval_jac_shared<M>(sh->Jsaved, tn, sh->yj, sh);
            g.sync();
            // AMICI's fJ callback returns checkFinite(J): recoverable if not finite
            bool bad = false;
            for (int e = g.lane(); e < NX * NX; e += G)
                if (!isfinite(sh->Jsaved[e])) bad = true;
            if (g.any(bad)) ret = 1;
        }
        if (g.lane() == 0 && ret == 0) ret = lu_factor_shared<M>(sh, gamma);
        if (jbad) { nje++; nstlj = nst; }
        jcur = jbad;
        nsetups++;
        g.sync();
        return g.bcast0(ret);
    }

    /// dense GETRS for column c (sundials_dense.c:150-182) + cvLsSolve's gamma-ratio scaling
    /// (cvodes_ls.c:1849-1854): x = -(P b)[.] gathered through the stored row order, forward and
    /// backward substitution in registers against the group's factors.
    PB_HD void lsolve_neg_b(int c, double* x) const {
        PB_UNROLL
        for (int i = 0; i < NX; ++i) x[i] = -sh->b[sh->perm[i] * ncol + c];
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
            const double sc = pb_div(2.0, 1.0 + gamrat);
            PB_UNROLL
            for (int i = 0; i < NX; ++i) x[i] *= sc;
        }
    }

    // ---------------------------------------------------------------- BDF coefficients
    /// cvSetBDF + cvSetTqBDF + cvSet (cvodes.c:6788-6879, 6617-6629); computed redundantly by
    /// every lane (a few dozen flops), the arrays are kept in shared memory.
    PB_HD void cvset() {
        double l_[cv::L_MAX + 1];
        double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
        l_[0] = l_[1] = xi_inv = xistar_inv = 1.0;
        PB_ROLL
        for (int i = 2; i <= cv::L_MAX; ++i) l_[i] = 0.0;
        alpha0 = alpha0_hat = -1.0;
        hsum = h;
        if (q > 1) {
            PB_ROLL
            for (int j = 2; j < q; ++j) {
                hsum += sh->tau[j - 1];
                xi_inv = pb_div(h, hsum);
                alpha0 -= pb_rcp((double)j);
                PB_ROLL
                for (int i = j; i >= 1; --i) l_[i] += l_[i - 1] * xi_inv;
            }
            alpha0 -= pb_rcp((double)q);
            xistar_inv = -l_[1] - alpha0;
            hsum += sh->tau[q - 1];
            xi_inv = pb_div(h, hsum);
            alpha0_hat = -l_[1] - xi_inv;
            PB_ROLL
            for (int i = q; i >= 1; --i) l_[i] += l_[i - 1] * xistar_inv;
        }
        double tq1 = sh->tq[1], tq3 = sh->tq[3];
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        const double tq2 = fabs(pb_div(A1, alpha0 * A2));
        const double tq5 = fabs(pb_div(A2 * xistar_inv, l_[q] * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double C = pb_div(xistar_inv, l_[q]);
                const double A3 = alpha0 + pb_rcp((double)q);
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = pb_div(1.0 - A4 + A3, A3);
                tq1 = fabs(C * Cpinv);
            } else {
                tq1 = 1.0;
            }
            hsum += sh->tau[q];
            xi_inv = pb_div(h, hsum);
            const double A5 = alpha0 - pb_rcp((double)(q + 1));
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = pb_div(1.0 - A6 + A5, A2);
            tq3 = fabs(pb_div(Cppinv, xi_inv * (q + 2) * A5));
        }
        g.sync();
        if (g.lane() == 0) {
            PB_ROLL
            for (int i = 0; i <= cv::L_MAX; ++i) sh->l[i] = l_[i];
            sh->tq[1] = tq1;
            sh->tq[2] = tq2;
            sh->tq[3] = tq3;
            sh->tq[4] = pb_div(cv::CORTES, tq2);
            sh->tq[5] = tq5;
        }
        g.sync();
        rl1 = pb_rcp(l_[1]);
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? pb_div(gamma, gammap) : 1.0;
    }

    // ---------------------------------------------------------------- history array
    /// cvPredict (sgn = +1, cvodes.c:6507-6597) and cvRestore (sgn = -1, cvodes.c:7373-7424):
    /// the Pascal-triangle sweep zn[j-1] += sgn*zn[j], k = 1..q, j = q..k, done per element in
    /// registers (one load and one store of the q+1 history entries).
    PB_HD void shift_history(double sgn) {
        PB_FOR_ELEMS(e) {
            double z[cv::L_MAX];
            PB_UNROLL
            for (int j = 0; j < cv::L_MAX; ++j) z[j] = (j <= q) ? sh->zn[j][e] : 0.0;
            PB_UNROLL
            for (int k = 1; k <= cv::Q_MAX; ++k) {
                PB_UNROLL
                for (int j = cv::Q_MAX; j >= k; --j) {
                    if (j <= q && k <= q) z[j - 1] += sgn * z[j];
                }
            }
            PB_UNROLL
            for (int j = 0; j < cv::Q_MAX; ++j)
                if (j < q) sh->zn[j][e] = z[j];
        }
    }
    /// cvRescale (cvodes.c:6425-6495)
    PB_HD void rescale() {
        PB_FOR_ELEMS(e) {
            double factor = eta;
            for (int j = 1; j <= q; ++j) {
                sh->zn[j][e] *= factor;
                factor *= eta;
            }
        }
        h = hscale * eta;
        hscale = h;
        nscon = 0;
    }
    /// cvAdjustOrder -> cvIncreaseBDF / cvDecreaseBDF (cvodes.c:6134-6143, 6256-6351, 6363-6415)
    PB_HD void adjust_order(int deltaq) {
        if ((q == 2) && (deltaq != 1)) return;
        double lc[cv::L_MAX + 1];
        for (int i = 0; i <= cv::L_MAX; ++i) lc[i] = 0.0;
        lc[2] = 1.0;
        int src, jmax;
        double A1 = 0.0;
        if (deltaq == 1) {
            double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0;
            double hsum = hscale;
            for (int j = 1; j < q; ++j) {
                hsum += sh->tau[j + 1];
                const double xi = pb_div(hsum, hscale);
                prod *= xi;
                alpha0 -= pb_rcp((double)(j + 1));
                alpha1 += pb_rcp(xi);
                for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xiold + lc[i - 1];
                xiold = xi;
            }
            A1 = pb_div(-alpha0 - alpha1, prod);
            src = L;   // new column zn[L] = A1 * (saved correction in zn[qmax]); L = q+1 <= qmax
            jmax = q;  // N_VScaleAddMulti(q-1, l+2, zn[L], zn+2, zn+2)
        } else {
            double hsum = 0.0;
            for (int j = 1; j <= q - 2; ++j) {
                hsum += sh->tau[j];
                const double xi = pb_div(hsum, hscale);
                for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xi + lc[i - 1];
            }
            for (int j = 2; j < q; ++j) lc[j] = -lc[j];
            src = q;
            jmax = q - 1;  // zn[j] -= l[j]*zn[q], j = 2..q-1
        }
        PB_FOR_ELEMS(e) {
            if (deltaq == 1) sh->zn[src][e] = A1 * sh->zn[cv::Q_MAX][e];
            const double zs = sh->zn[src][e];
            for (int j = 2; j <= jmax; ++j) sh->zn[j][e] += lc[j] * zs;
        }
    }

    // ---------------------------------------------------------------- nonlinear solve
    /// cvNls / cvStgrNls + SUNNonlinSolSolve_Newton + cvNlsConvTest{,SensStg} for the columns of
    /// `phase` (cvodes.c:6894-7007, 7167-7210; sunnonlinsol_newton.c:176-335; cvodes_nls.c:
    /// 298-342; cvodes_nls_stg.c:238-375).  The state corrector and the staggered sensitivity
    /// corrector are the same Newton iteration on different columns with the same factors, so
    /// they share one copy of the code.  On success the corrected vector is zn[0]+acor.
    PB_HD int nls(int phase, int nflag) {
        const int cmin = phase_cmin(phase), cmax = phase_cmax(phase);
        bool setup_pending = false;
        if (phase == PH_STATE) {
            convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::NO_FAILURES : cv::FAIL_OTHER;
            setup_pending = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                            (nst >= nstlp + cv::MSBP) || (fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
        }
        PB_FOR_COLS(c, cmin, cmax) {
            PB_UNROLL_X
            for (int i = 0; i < NX; ++i) sh->acor[(i) * ncol + (c)] = 0.0;
        }
        bool jbad = false;
        bool top = true;  // residual at the top of a solve attempt: failures there are final
        int m_it = 0;
        for (;;) {
            int ret = rhs(phase, tn, false);
            if (ret != 0 && top) return ret;
            if (ret == 0 && setup_pending) {
                if (jbad) convfail = cv::FAIL_BAD_J;
                if (phase == PH_SENS) {  // Jacobian at the converged state (cv_y), cvodes_nls_stg.c:252
                    g.sync();
                    if (g.lane() == 0) {
                        PB_UNROLL_X
                        for (int i = 0; i < NX; ++i) sh->yj[i] = sh->ysh[i];
                    }
                }  // state phase: rhs() has just put zn[0]+acor into sh->yj
                const int lr = lsetup();
                setup_pending = false;
                if (phase == PH_STATE) jcurNLS = jcur;
                else jcurNLSS = jcur;
                gamrat = 1.0;
                gammap = gamma;
                crate = 1.0;
                crateS = 1.0;
                nstlp = nst;
                if (lr != 0) return cv::CONV_RECVR;
            }
            top = false;
            if (ret == 0) {
                if (phase == PH_STATE) nni++;
                else nniS++;
                double loc = 0.0;  // max over own columns of sum((delta*ewt)^2)
                PB_FOR_COLS(c, cmin, cmax) {
                    double x[NX];
                    lsolve_neg_b(c, x);
                    double s = 0.0;
                    PB_UNROLL
                    for (int i = 0; i < NX; ++i) {
                        sh->acor[(i) * ncol + (c)] += x[i];
                        const double prod = x[i] * sh->ewt[(i) * ncol + (c)];
                        s += prod * prod;
                    }
                    if (s > loc) loc = s;
                }
                const double del = group_wrms(g, loc, NX);
                double cr = (phase == PH_STATE) ? crate : crateS;
                if (m_it > 0) cr = fmax(cv::CRDOWN * cr, pb_div(del, delp));
                if (phase == PH_STATE) crate = cr;
                else crateS = cr;
                const double dcon = pb_div(del * fmin(1.0, cr), sh->tq[4]);
                if (dcon <= 1.0) {
                    const double an = (m_it == 0) ? del : wnorm1(cmin, cmax, sh->acor);
                    if (phase == PH_STATE) {
                        acnrm = an;
                        jcurNLS = false;
                    } else {
                        if (opt.errconS) acnrmS = an;
                        jcurNLSS = false;
                    }
                    jcur = false;
                    return 0;
                }
                if ((m_it >= 1) && (del > cv::RDIV * delp)) {
                    ret = cv::CONV_RECVR;
                } else {
                    delp = del;
                    m_it++;
                    if (m_it >= cv::NLS_MAXCOR) ret = cv::CONV_RECVR;
                    else continue;
                }
            }
            // failure inside the iteration: retry once with a fresh Jacobian if it was stale
            if ((ret > 0) && !((phase == PH_STATE) ? jcurNLS : jcurNLSS)) {
                setup_pending = true;
                jbad = true;
                top = true;
                m_it = 0;
                PB_FOR_COLS(c, cmin, cmax) {
                    PB_UNROLL_X
                    for (int i = 0; i < NX; ++i) sh->acor[(i) * ncol + (c)] = 0.0;
                }
                continue;
            }
            return ret;
        }
    }

    // ---------------------------------------------------------------- step control
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
        qwait--;
        const bool save = (qwait == 1) && (q != cv::Q_MAX);
        PB_FOR_ELEMS(e) {
            const double a = sh->acor[e];
            for (int j = 0; j <= q; ++j) sh->zn[j][e] += sh->l[j] * a;
            if (save) sh->zn[cv::Q_MAX][e] = a;
        }
        if (save) saved_tq5 = sh->tq[5];
    }

    /// cvPrepareNextStep with cvSetEta, cvComputeEtaqm1 / cvComputeEtaqp1 / cvChooseEta
    /// (cvodes.c:7747-8000)
    PB_HD void prepare_next_step(double dsm) {
        if (etamax == 1.0) {
            qwait = qwait > 2 ? qwait : 2;
            qprime = q;
            hprime = h;
            eta = 1.0;
            return;
        }
        etaq = pb_rcp(pb_pow(cv::BIAS2 * dsm, pb_rcp((double)L)) + cv::ADDON);
        qprime = q;
        eta = etaq;
        if (qwait == 0) {
            qwait = 2;
            const int ecmax = errtest_cmax();
            etaqm1 = 0.0;
            if (q > 1) {
                const int qq = q;
                double ddn = wnorm1(0, ecmax, sh->zn[qq]);
                ddn = ddn * sh->tq[1];
                etaqm1 = pb_rcp(pb_pow(cv::BIAS1 * ddn, pb_rcp((double)q)) + cv::ADDON);
            }
            etaqp1 = 0.0;
            if (q != cv::Q_MAX) {
                if (saved_tq5 != 0.0) {
                    const double hr = pb_div(h, sh->tau[2]);
                    double hp = 1.0;
                    for (int i = 0; i < L; ++i) hp *= hr;  // SUNRpowerI
                    const double cquot = pb_div(sh->tq[5], saved_tq5) * hp;
                    double dup = wnorm2(0, ecmax, sh->zn[cv::Q_MAX], -cquot, sh->acor, 1.0);
                    dup = dup * sh->tq[3];
                    etaqp1 = pb_rcp(pb_pow(cv::BIAS3 * dup, pb_rcp((double)(L + 1))) + cv::ADDON);
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
                PB_FOR_COLS(c, 0, ecmax) {
                    PB_UNROLL_X
                    for (int i = 0; i < NX; ++i) sh->zn[cv::Q_MAX][(i) * ncol + (c)] = sh->acor[(i) * ncol + (c)];
                }
            }
        }
        // cvSetEta (hmax_inv = 0, hmin = 0)
        if ((eta > cv::ETA_MIN_FX) && (eta < cv::ETA_MAX_FX)) {
            eta = 1.0;
            hprime = h;
        } else {
            if (eta >= cv::ETA_MAX_FX) eta = fmin(eta, etamax);
            else eta = fmax(eta, cv::ETA_MIN);
            hprime = h * eta;
            if (qprime < q) nscon = 0;
        }
    }

    /// cvBDFStab (cvodes.c:8158-8218)
    PB_HD void bdf_stab() {
        if (q >= 3) {
            double sqv[2];
            for (int w = 0; w < 2; ++w) {
                const int jj = q - w;
                sqv[w] = wnorm1(0, 0, sh->zn[jj]);
            }
            int factorial = 1;
            for (int i = 1; i <= q - 1; ++i) factorial *= i;
            const double sq = pb_div(factorial * q * (q + 1) * acnrm, fmax(sh->tq[5], cv::TINY));
            const double sqm1 = sqv[0] * (factorial * q);
            const double sqm2 = sqv[1] * factorial;
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
                const int ldflag = sldet_shared(sh->ssdat, q);
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

    /// zn[1] = h * f(tn, zn[0]) for the state and (if any) the sensitivities; acor is scratch
    PB_HD int reload_zn1() {
        PB_FOR_ELEMS(e) sh->acor[e] = 0.0;
        const int r = rhs_all_raw(tn, sensi);
        if (r) return r;
        PB_FOR_ELEMS(e) sh->zn[1][e] = h * sh->b[e];
        return 0;
    }

    /// cvStep (cvodes.c:5800-6094) with cvHandleNFlag (:7309-7363) and cvDoErrorTest
    /// (:7461-7614) folded into ONE failure handler, and cvAdjustParams (:6112-6122) folded into
    /// the top of the attempt loop, so that every history operation (Pascal sweep, order change,
    /// rescale) and both correctors exist once in the instruction stream.  The sequence of
    /// operations per attempt is the reference's.
    PB_HD void step_begin() {
        ncf_ = nef_ = ncfS_ = nefS_ = 0;
        dsm_ = dsmS_ = 0.0;
        saved_t_ = tn;
        nflag_ = cv::FIRST_CALL;
        do_rescale_ = (nst > 0) && (hprime != h);
        dq_ = do_rescale_ ? (qprime - q) : 0;
        undo_ = restart_ = false;
    }

    /// One attempt of the step (after a failed attempt: first the restore pass).  Returns 0 when
    /// the step was completed (history updated, next step/order chosen), 1 when another attempt
    /// is needed, or a negative failure code.
    PB_HD int step_attempt() {
        for (;;) {
            if (undo_) {
                tn = saved_t_;  // cvRestore
            } else {
                if (restart_) {
                    // already at order 1 after MXNEF1 error-test failures: reload zn[1] from
                    // scratch at the reduced step (cvodes.c:7536-7613)
                    restart_ = false;
                    h *= eta;
                    hscale = h;
                    qwait = cv::LONG_WAIT;
                    nscon = 0;
                    if (reload_zn1()) return sensi ? SIM_UNREC_SRHSFUNC_ERR : SIM_UNREC_RHSFUNC_ERR;
                }
                if (dq_ != 0) {
                    adjust_order(dq_);
                    q += dq_;
                    L = q + 1;
                    qwait = L;
                    dq_ = 0;
                }
                if (do_rescale_) {
                    rescale();
                    do_rescale_ = false;
                }
                tn += h;  // cvPredict
                if (tstopset && ((tn - tstop) * h > 0.0)) tn = tstop;
            }
            double sgn = undo_ ? -1.0 : 1.0;
            PB_OPAQUE_F64(sgn);
            shift_history(sgn);
            if (!undo_) break;
            undo_ = false;
        }
        cvset();
        int fail = 0;  // 1/3: corrector failure (state/sens); 2/4: error-test failure (state/sens)
        int ret = 0;
        for (int phase_ = PH_STATE; phase_ <= PH_SENS; ++phase_) {
            int phase = phase_;
            PB_OPAQUE_INT(phase);
            if (phase == PH_SENS) {
                if (fail || !sensi) break;
                ncf_ = nef_ = 0;
                publish_state();
                // f at the converged y (cvodes.c:5947-5955); AMICI's sensitivity RHS does not
                // use it -- kept for the recoverable-failure semantics
                int bad = 0;
                if (g.lane() == 0) {
                    PB_UNROLL_X
                    for (int i = 0; i < NX; ++i) sh->yj[i] = sh->ysh[i];
                    bad = eval_xdot_shared<M>(sh, tn) ? 0 : 1;
                }
                nfe++;
                if (g.bcast0(bad)) { fail = 5; break; }
                refresh_Jcur(tn);
            }
            ret = nls(phase, nflag_);
            if (ret != 0) {
                fail = 1 + 2 * phase;
            } else if (phase == PH_STATE) {
                dsm_ = acnrm * sh->tq[2];
                if (dsm_ > 1.0) fail = 2;
            } else if (opt.errconS) {
                dsmS_ = acnrmS * sh->tq[2];
                if (dsmS_ > 1.0) fail = 4;
                else if (dsmS_ > dsm_) dsm_ = dsmS_;
            }
        }
        if (fail == 0) {
            complete_step();
            prepare_next_step(dsm_);
            if (opt.stldet) bdf_stab();
            etamax = (nst <= cv::SMALL_NST) ? cv::ETA_MAX_ES : cv::ETA_MAX_GS;
            // acor *= tq[2] (local error estimate) is not used downstream and is skipped
            return 0;
        }
        if (fail == 5) {  // cvodes.c:5951-5955: straight back to the prediction, no restore
            nflag_ = cv::PREV_CONV_FAIL;
            return 1;
        }
        undo_ = true;
        if (fail == 1 || fail == 3) {  // cvHandleNFlag
            if (fail == 1) { ncfn++; ncf_++; } else { ncfnS++; ncfS_++; }
            if (ret < 0) return ret;
            etamax = 1.0;
            if ((fabs(h) <= 0.0) || ((fail == 1 ? ncf_ : ncfS_) == cv::MXNCF)) {
                if (ret == cv::CONV_RECVR) return SIM_CONV_FAILURE;
                if (ret == cv::RHSFUNC_RECVR) return SIM_REPTD_RHSFUNC_ERR;
                return SIM_REPTD_SRHSFUNC_ERR;
            }
            eta = cv::ETA_CF;  // max(eta_cf, hmin/|h|), hmin = 0
            nflag_ = cv::PREV_CONV_FAIL;
            do_rescale_ = true;
        } else {  // cvDoErrorTest
            int e;
            if (fail == 2) { e = ++nef_; netf++; } else { e = ++nefS_; netfS++; }
            nflag_ = cv::PREV_ERR_FAIL;
            if ((fabs(h) <= 0.0) || (e == cv::MXNEF)) return SIM_ERR_FAILURE;
            etamax = 1.0;
            if (e <= cv::MXNEF1) {
                eta = pb_rcp(pb_pow(cv::BIAS2 * (fail == 2 ? dsm_ : dsmS_), pb_rcp((double)L)) + cv::ADDON);
                eta = fmax(cv::ETA_MIN_EF, eta);
                if (e >= cv::SMALL_NEF) eta = fmin(eta, cv::ETA_MAX_EF);
                do_rescale_ = true;
            } else if (q > 1) {
                eta = cv::ETA_MIN_EF;
                dq_ = -1;
                do_rescale_ = true;
            } else {
                eta = cv::ETA_MIN_EF;
                restart_ = true;
            }
        }
        return 1;
    }

    // ---------------------------------------------------------------- initial step
    /// cvHin with cvUpperBoundH0 and cvYddNorm (cvodes.c:5436-5773); zn[1] holds y'(t0) unscaled.
    PB_HD int hin(double tout) {
        const double tdiff = tout - tn;
        if (tdiff == 0.0) return SIM_TOO_CLOSE;
        const int sign = (tdiff > 0.0) ? 1 : -1;
        const double tdist = fabs(tdiff);
        const double tround = cv::UROUND * fmax(fabs(tn), fabs(tout));
        if (tdist < 2.0 * tround) return SIM_TOO_CLOSE;
        const double hlb = cv::HLB_FACTOR * tround;
        const bool withS = sensi && opt.errconS;
        const int ecmax = errtest_cmax();
        // cvUpperBoundH0
        double hub_inv = 0.0;
        PB_FOR_COLS(c, 0, ecmax) {
            PB_UNROLL_X
            for (int i = 0; i < NX; ++i) {
                const double den = cv::HUB_FACTOR * fabs(sh->zn[0][(i) * ncol + (c)]) + pb_rcp(sh->ewt[(i) * ncol + (c)]);
                const double v = pb_div(fabs(sh->zn[1][(i) * ncol + (c)]), den);
                if (v > hub_inv) hub_inv = v;
            }
        }
        hub_inv = group_max(g, hub_inv);
        double hub = cv::HUB_FACTOR * tdist;
        if (hub * hub_inv > 1.0) hub = pb_rcp(hub_inv);
        double hg = sqrt(hlb * hub);
        if (hub < hlb) {
            h = (sign == -1) ? -hg : hg;
            return 0;
        }
        double hs = hg, hnew = hg;
        for (int count1 = 1; count1 <= cv::HIN_MAX_ITERS; ++count1) {
            bool hgOK = false;
            int retval = 0;
            double yddnrm = 0.0;
            for (int count2 = 1; count2 <= cv::HIN_MAX_ITERS; ++count2) {
                // cvYddNorm: f at (tn + hg, zn0 + hg*zn1), difference quotient, WRMS norm
                const double hgs = hg * sign;
                PB_FOR_ELEMS(e) sh->acor[e] = hgs * sh->zn[1][e];
                retval = rhs_all_raw(tn + hgs, withS);
                if (retval == 0) {
                    const double hinv = pb_rcp(hgs);
                    yddnrm = wnorm2(0, ecmax, sh->b, hinv, sh->zn[1], -hinv);
                    hgOK = true;
                    break;
                }
                hg *= 0.2;
            }
            PB_FOR_ELEMS(e) sh->acor[e] = 0.0;
            if (!hgOK) {
                if (count1 <= 2) return retval == cv::RHSFUNC_RECVR ? SIM_REPTD_RHSFUNC_ERR : SIM_REPTD_SRHSFUNC_ERR;
                hnew = hs;
                break;
            }
            hs = hg;
            hnew = (yddnrm * hub * hub > 2.0) ? sqrt(pb_div(2.0, yddnrm)) : sqrt(hg * hub);
            if (count1 == cv::HIN_MAX_ITERS) break;
            const double hrat = pb_div(hnew, hg);
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
    /// Prologue of CVode(tout) in CV_NORMAL mode with tstop = tout, as AMICI drives it
    /// (cvodes.c:2890-3347; deps/AMICI/src/solver.cpp:108-161): first-call initialisation (ewt,
    /// zn[1], initial step) or the stop tests of a later call.  Returns 0 to enter the step loop,
    /// TSTOP_RETURN if tn is already at tout, or a negative failure code.
    PB_HD int cvode_begin(double tout) {
        // Solver::run -> setStopTime(tout); apply_max_num_steps()
        tstop = tout;
        tstopset = true;
        if (opt.maxsteps <= nst) return SIM_ERROR;
        mxstep_ = opt.maxsteps - nst;
        nstloc_ = 0;
        if (nst == 0) {
            if (!ewt_set()) return SIM_ILL_INPUT;
            h = 1.0;  // zn[1] = f(t0, y0), znS[1] = fS(t0, y0, yS0), unscaled
            const int r = reload_zn1();
            if (r) return r == cv::RHSFUNC_RECVR ? SIM_FIRST_RHSFUNC_ERR : SIM_FIRST_SRHSFUNC_ERR;
            if ((tstop - tn) * (tout - tn) <= 0.0) return SIM_ILL_INPUT;
            const int hr = hin(tout);
            if (hr != 0) return hr;
            if ((tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
            hscale = h;
            hprime = h;
            PB_FOR_ELEMS(e) sh->zn[1][e] *= h;
        } else {
            const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
            if (fabs(tn - tstop) <= troundoff) {
                tstopset = false;
                tn = tstop;  // AMICI reads t_ = tret = tstop
                return cv::TSTOP_RETURN;
            } else if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                eta = pb_div(hprime, h);
            }
        }
        return 0;
    }

    // ---------------------------------------------------------------- likelihood
    /// One output time: llh and sllh contributions (Model::addObservableObjective /
    /// addObservableObjectiveSensitivity, deps/AMICI/src/model.cpp:1106-1152 with
    /// fsigmay :2087-2119, fdJydy/fdJydsigma/fdJydp/fdJydx :2184-2392; Jy formulas
    /// deps/AMICI/models/model_steadystate/{Jy,dJydy,dJydsigma}.cpp).  The state at the output
    /// time is zn[0] of column 0.  Returns false if a sigma is not positive
    /// (Model::checkSigmaPositivity throws).
    PB_HD bool data_point(double t, const double* my, const double* sd, double& llh, double* y_out,
                          double* sigmay_out) {
        g.sync();
        if (g.lane() == 0) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) sh->ysh[i] = sh->zn[0][(i) * ncol + (0)];
        }
        g.sync();
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
        PB_FOR_COLS(c, 1, ns) {
            double dydp[NY], dsdp[NY];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) { dydp[iy] = 0.0; dsdp[iy] = 0.0; }
            const int ip = sh->plist[c - 1];
            M::dydp(dydp, t, sh->ysh, sh->p, sh->k, ip);
            M::dsigmaydp(dsdp, t, sh->p, sh->k, ip);
            double dJydp = 0.0;
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJydy[iy] * dydp[iy];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJyds[iy] * dsdp[iy];
            double a = dJydp;
            PB_UNROLL
            for (int ix = 0; ix < NX; ++ix) a += dJydx[ix] * sh->zn[0][(ix) * ncol + (c)];
            sh->sllh[c] -= a;
        }
        return true;
    }

    // ---------------------------------------------------------------- one simulation
    /// Start one simulation (ForwardProblem::workForwardProblem, forwardproblem.cpp:67-203) at
    /// the given tolerances: x0/sx0 (Model::initialize, model.cpp:249-275), CVodeInit
    /// (cvodes.c:677-830).  sh->p / sh->k / sh->plist must be set.
    PB_HD void begin(const ConditionIO& io, double rtol, double atol) {
        io_ = io;
        opt.rtol = rtol;
        opt.atol = atol;
        ns = io.ns;
        ncol = ns + 1;
        sensi = ns > 0;
        g.sync();
        PB_FOR_ALL_COLS(c) {
            double v[NX];
            PB_UNROLL
            for (int i = 0; i < NX; ++i) v[i] = 0.0;
            if (c == 0) M::x0(v, opt.t0, sh->p, sh->k);
            else M::sx0(v, opt.t0, sh->p, sh->k, sh->plist[c - 1]);
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                sh->zn[0][(i) * ncol + (c)] = v[i];
                for (int j = 1; j < cv::L_MAX; ++j) sh->zn[j][(i) * ncol + (c)] = 0.0;
                sh->acor[(i) * ncol + (c)] = 0.0;
                sh->ewt[(i) * ncol + (c)] = 1.0;
                sh->b[(i) * ncol + (c)] = 0.0;
            }
            sh->sllh[c] = 0.0;
        }
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
        for (int e = g.lane(); e < NX * NX; e += G) {
            sh->Jsaved[e] = 0.0;
            sh->Jcur[e] = 0.0;
        }
        for (int e = g.lane(); e < NX * NP; e += G) sh->dfdp[e] = 0.0;
        if (g.lane() == 0) {
            for (int i = 0; i <= cv::L_MAX; ++i) { sh->l[i] = 0.0; sh->tau[i] = 0.0; }
            for (int i = 0; i < 6; ++i) sh->tq[i] = 0.0;
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 4; ++k) sh->ssdat[i][k] = 0.0;
        }
        g.sync();
        llh_ = 0.0;
        it_ = 0;
        sim_status_ = SIM_SUCCESS;
        in_cvode_ = in_step_ = false;
        step_begin();
    }

    /// Advance the simulation by at most ONE step attempt (plus whatever bookkeeping precedes it:
    /// output times that need no integration, the CVode prologue, the per-step checks of the CVode
    /// loop cvodes.c:3366-3490, and after a completed step the tstop tests cvodes.c:3504-3605).
    /// Returns true when the simulation has ended; status() / llh() then hold the result and
    /// d llh / d p (w.r.t. the unscaled parameter) is in sh->sllh[1..ns].
    PB_HD bool advance() {
        if (!in_step_) {
            if (!in_cvode_) {
                for (;;) {
                    if (it_ >= io_.nt) return finish(SIM_SUCCESS);
                    tout_ = io_.t[it_];
                    if (isinf(tout_)) return finish(SIM_SUCCESS);  // post-equilibration: out of scope
                    bool integrate = (tout_ > opt.t0) && (tn < tout_);
                    if (integrate) {  // is_next_t_too_close, forwardproblem.cpp:23-35
                        const double tdiff = tout_ - tn;
                        if (fabs(tdiff) < 2.0 * cv::UROUND * fmax(fabs(tn), fabs(tout_))) integrate = false;
                    }
                    if (integrate) break;
                    double* yo = io_.y_out ? io_.y_out + (size_t)it_ * NY : nullptr;
                    double* so = io_.sigmay_out ? io_.sigmay_out + (size_t)it_ * NY : nullptr;
#if defined(PB200_HOST_TRACE)
                    if (io_.trace) { io_.trace[2 * it_] = nst; io_.trace[2 * it_ + 1] = q; }
#endif
                    if (!data_point(tout_, io_.my + (size_t)it_ * NY, io_.sd + (size_t)it_ * NY, llh_, yo, so))
                        return finish(SIM_ERROR);
                    it_++;
                }
                const int r = cvode_begin(tout_);
                if (r < 0) return finish(r);
                if (r == cv::TSTOP_RETURN) return false;  // already at tout: data point on the next call
                in_cvode_ = true;
            }
            // top of the CVode step loop
            if (nst > 0) {
                if (!ewt_set()) return finish(SIM_ILL_INPUT);
            }
            if (nstloc_ >= mxstep_) return finish(SIM_TOO_MUCH_WORK);
            const double nrm = wnorm1(0, errtest_cmax(), sh->zn[0]);
            if (cv::UROUND * nrm > 1.0) return finish(SIM_TOO_MUCH_ACC);
            step_begin();
            in_step_ = true;
        }
        const int r = step_attempt();
        if (r == 1) return false;
        in_step_ = false;
        if (r < 0) return finish(r);
        nstloc_++;
        const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
        if (fabs(tn - tstop) <= troundoff) {
            tn = tstop;
            tstopset = false;
            in_cvode_ = false;
        } else {
            if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                eta = pb_div(hprime, h);
            }
            if ((tn - tout_) * h >= 0.0) in_cvode_ = false;
        }
        return false;
    }

    PB_HD bool finish(int status) {
        sim_status_ = status;
        if (status != SIM_SUCCESS) llh_ = NAN;  // ReturnData::invalidate (rdata.cpp)
        g.sync();
        llh_ = g.bcast0(llh_);
        return true;
    }
    PB_HD int status() const { return sim_status_; }
    PB_HD double llh() const { return llh_; }

    /// Whole simulation in one call (host harness, and the non-lock-step kernel variant).
    PB_HD int simulate(const ConditionIO& io, double rtol, double atol, double& llh_out) {
        begin(io, rtol, atol);
        while (!advance()) {}
        llh_out = llh_;
        return sim_status_;
    }

    PB_HD void export_stats(int64_t* s, int trials) const {
        s[STAT_NST] = nst; s[STAT_NFE] = nfe; s[STAT_NSETUPS] = nsetups; s[STAT_NJE] = nje;
        s[STAT_NNI] = nni; s[STAT_NNIS] = nniS; s[STAT_NFSE] = nfSe; s[STAT_NETF] = netf;
        s[STAT_NETFS] = netfS; s[STAT_NCFN] = ncfn; s[STAT_NCFNS] = ncfnS; s[STAT_NOR] = nor;
        s[STAT_TRIALS] = trials; s[STAT_QSUM] = qsum; s[14] = 0; s[15] = 0;
    }
#undef PB_FOR_COLS
#undef PB_FOR_ALL_COLS
#undef PB_FOR_ELEMS
};

}  // namespace pb200
#endif
