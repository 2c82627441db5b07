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
//   * a condition is owned by a group of G lanes of one warp (G = 1..32, normally the
//     power of two >= nx), several conditions per warp, kept in step per step attempt;
//   * the state vector and every sensitivity vector is a *column* of an nx x (1+nplist)
//     array stored column by column (element (i, c) at [c*nx + i]).  Lane l owns the ROWS
//     i = l, l+G, ... of every column: residuals, J*s products, Newton solves and updates are
//     row-parallel loops over the columns, element-wise history operations (prediction,
//     rescaling, order change, correction) run over the flat element index with every lane
//     busy, and the lanes of a group always touch consecutive 8-byte words;
//   * the Nordsieck history zn[0..5], the accumulated correction and the error weights of
//     all columns live in the group's slice of shared memory (the first version of this
//     kernel kept them in registers with fully unrolled code and was instruction-fetch
//     bound, profiles/r01_a_summary.txt);
//   * the Newton matrix I - gamma*J is INVERTED at every linear-solver setup (in-place
//     Gauss-Jordan with partial pivoting, one matrix column per lane), so that each
//     Newton iteration of the state and of all nplist sensitivities (the staggered
//     corrector's "same matrix, nplist right-hand sides") is a row-parallel mat-vec with
//     the lane's rows of the inverse in registers instead of two triangular solves;
//   * WRMS norms: squared weighted entries go through a shared scratch array (odd column
//     stride: conflict-free), one lane sums a column in CVODES' order, max over lanes by
//     warp shuffles;
//   * BDF coefficients (l, tq) live in registers, computed redundantly by every lane;
//   * model functions are the generated __host__ __device__ code of csrc/models/<name>.cuh,
//     evaluated by the group's first lane.
//
// The same source compiles for the host with the SingleLane group (G = 1); that
// instantiation exists only in tests/host_harness to diff the algorithm against
// the CVODES oracle without a GPU.  The product library has no host integrator.
//
// ATTRIBUTION.  The scalar step-control logic in this file (initial step size `hin`, BDF coefficient
// setup `set_bdf` / `cvset`, error-test / convergence-failure handling, order and step-size selection
// `prepare_next_step`, stability-limit detection `sldet`) restates the algorithm of CVODES
// (SUNDIALS 7.1.1, src/cvodes/cvodes.c: cvHin, cvSetBDF, cvSetTqBDF, cvDoErrorTest, cvHandleNFlag,
// cvPrepareNextStep, cvSLdet) with its constants and, where that helps a reader compare the two,
// its local names -- it has to, so that the step sequence stays that of the reference.  SUNDIALS is
// Copyright (c) 2002-2024, Lawrence Livermore National Security and Southern Methodist University,
// all rights reserved, and distributed under the BSD 3-Clause License
// (deps/AMICI/ThirdParty/sundials/LICENSE in the reference tree):
//   Redistribution and use in source and binary forms, with or without modification, are permitted
//   provided that the following conditions are met: (1) redistributions of source code must retain the
//   above copyright notice, this list of conditions and the following disclaimer; (2) redistributions in
//   binary form must reproduce the above copyright notice, this list of conditions and the following
//   disclaimer in the documentation and/or other materials provided with the distribution; (3) neither
//   the name of the copyright holder nor the names of its contributors may be used to endorse or promote
//   products derived from this software without specific prior written permission.
//   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR
//   IMPLIED WARRANTIES, INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND
//   FITNESS FOR A PARTICULAR PURPOSE ARE DISCLAIMED.  IN NO EVENT SHALL THE COPYRIGHT HOLDER OR
//   CONTRIBUTORS BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL
//   DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR SERVICES; LOSS OF USE,
//   DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER
//   IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF
//   THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
// Data layout, lane mapping, staging, linear algebra and norms in this file are this project's own.
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
    STAT_NCFN, STAT_NCFNS, STAT_NOR, STAT_TRIALS, STAT_QSUM, STAT_NFALLBACK = 16, STAT_CYC_SPARSE = 17, NSTATS = 20
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
    // bit 0: sensitivities take part in the error test (AMICI forces 1, solver.cpp:646).  Bit 1 (steady-state
    // instantiation only, set by pb200_create together with bit 0): steadyStateSensitivityMode newtonOnly -- a
    // simulation to steady state runs WITHOUT the sensitivities, which are solved for at the steady state
    // (steadystateproblem.cpp:728-736, 91-103)
    // Bit 2 (steady-state instantiation only, with bit 0): the SIMULTANEOUS corrector (CV_SIMULTANEOUS, AMICI's
    // InternalSensitivityMethod::simultaneous) instead of the staggered one.
    int errconS;
    PB_HD bool ss_sens_by_solve() const { return (errconS & 2) != 0; }
    PB_HD bool simultaneous() const { return (errconS & 4) != 0; }
    // steady states: > 0 Newton iterations tried before / after the simulation (AMICI default 0 = none);
    // < 0: Newton's method ONLY, with -newton_maxsteps iterations at most and no simulation if it fails
    // (steadyStateComputationMode newtonOnly, steadystateproblem.cpp:143-146)
    int newton_maxsteps;
    PB_HD int newton_steps() const { return newton_maxsteps < 0 ? -newton_maxsteps : newton_maxsteps; }
    PB_HD bool newton_only() const { return newton_maxsteps < 0; }
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
/// Column data: element (state index i, column c) of a vector lives at [c*NX + i]; column 0 is
/// the state, columns 1..ns the sensitivities of the current condition.
template <class M>
struct GroupShared {
    static constexpr int NX = M::NX;
    // columns: state + one per DYNAMIC model parameter (one the state depends on).  For the other
    // parameters (noise / observable parameters) the sensitivity is identically zero -- sx0 = 0 and
    // sxdot = J*0 + 0 -- so their columns contribute nothing to any norm or sum the reference
    // computes and are not stored or integrated; d llh / d p for them comes from dydp / dsigmaydp.
    static constexpr int NPD = M::NPD > 0 ? M::NPD : 0;
    static constexpr int NC = NPD + 1;
    static constexpr int NXS = NX | 1;    // odd column stride of the norm scratch: conflict-free column sums
    alignas(16) double zn[cv::L_MAX][NX * NC];  // Nordsieck history
    alignas(16) double acor[NX * NC];           // accumulated corrector update
    alignas(16) double ewt[NX * NC];            // error weights
    alignas(16) double b[NX * NC];              // right-hand side / Newton residual (read by every lane)
    // scratch: corrected sensitivities (stride NX), squared norm terms (stride NXS), STALD work arrays (45)
    alignas(16) double scr[NC * (NX + 1) > 48 ? NC * (NX + 1) : 48];
    // group-uniform data
    alignas(16) double Jsaved[NX * NX];  // Jacobian at its last evaluation (cvLs savedJ), column-major
    alignas(16) double Ainv[NX * NX];    // inverse of I - gamma*J, column-major
    alignas(16) double Jcur[NX * NX];    // J(tn, y_n) for the sensitivity right-hand sides
    // d xdot / d p at (tn, y_n): the columns of the dynamic parameters only, column of parameter ip at
    // [rank(ip)*NX], rank = its position among the dynamic parameters
    alignas(16) double dfdp[NX * (NPD > 0 ? NPD : 1)];
    double sllh[(M::NP > 0 ? M::NP : 0) + 1];  // d llh / d p accumulator per plist position (index j + 1)
    double ysh[NX];          // converged state, argument of the Jacobian / observable functions
    double yj[NX];           // state handed to the out-of-line model functions (xdot, Jacobian)
    double fout[NX];         // xdot returned by eval_xdot_shared
    double p[M::NP > 0 ? M::NP : 1];  // unscaled model parameters
    double k[M::NK > 0 ? M::NK : 1];  // constants (fixed parameters)
    double tau[cv::L_MAX + 1];
    double ssdat[6][4];      // STALD history (cvBDFStab)
    int perm[NX];            // row order after the pivot swaps of the inversion
    int plist[M::NP > 0 ? M::NP : 1];   // sensitivity parameter list of the condition (all entries)
    int cplist[NPD > 0 ? NPD : 1];      // model parameter of sensitivity column c (index c - 1)
    int cdcol[NPD > 0 ? NPD : 1];       // its column in dfdp
    int colof[M::NP > 0 ? M::NP : 1];   // column of plist position j, 0 = not integrated (sx == 0)
    int nsd;                            // number of integrated sensitivity columns
};

// ---- out-of-line helpers (one copy of the code each; they touch only group-shared
//      memory and scalars, so a real call costs a few registers) ------------------------
PB_HD_NOINLINE static double pb_pow(double x, double y) { return pow(x, y); }
PB_HD_NOINLINE static double pb_div(double a, double b) { return a / b; }
PB_HD double pb_rcp(double a) { return pb_div(1.0, a); }
/// 1.0 / n for the small integers of the BDF formulas
PB_HD double pb_inv_int(int n) { return pb_rcp((double)n); }

/// xdot(t, sh->yj) -> sh->fout; returns false if a component is not finite (AMICI checkFinite,
/// deps/AMICI/src/solver_cvodes.cpp:1149-1173)
template <class M>
PB_HD_NOINLINE static bool eval_xdot_shared(GroupShared<M>* sh, double t) {
    PB_ASSUME_SHARED(sh);
    M::xdot(sh->fout, t, sh->yj, sh->p, sh->k);
    bool ok = true;
    PB_UNROLL
    for (int i = 0; i < M::NX; ++i) ok = ok && isfinite(sh->fout[i]);
    return ok;
}

/// J(t, y) -> the structurally non-zero entries of Jdst (dense column-major, NX*NX), everything
/// in group-shared memory; the structural zeros are written once per condition (begin()).
template <class M>
PB_HD_NOINLINE static void eval_jac_shared(double* Jdst, double t, const double* y, const GroupShared<M>* sh) {
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(Jdst);
    PB_ASSUME_SHARED(y);
    M::jac(Jdst, t, y, sh->p, sh->k);
}

/// d xdot / d p (all columns) at (t, y) -> sh->dfdp, structural non-zeros only
template <class M>
PB_HD_NOINLINE static void eval_dfdp_shared(GroupShared<M>* sh, double t, const double* y) {
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(y);
    M::dxdotdp_dense(sh->dfdp, t, y, sh->p, sh->k);
}

/// cvSLdet (cvodes.c:8256-8556): dominant characteristic root from the STALD history.  The
/// three data columns k = 1..3 (the scaled norms of the order q-2, q-1, q terms) are processed by
/// three lanes of the group side by side; what one column needs from the others goes through the
/// group's scratch array.  Every branch is taken on group-uniform values.  Out of line.
template <class M, class Grp>
PB_HD_NOINLINE static int sldet_shared(GroupShared<M>* sh, const Grp g, int q) {
    PB_ASSUME_SHARED(sh);
    constexpr int G = Grp::G;
    const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
    const double (*ss)[4] = sh->ssdat;
    // scratch layout, three entries each (index k-1); qjk is [k-1][j-1]
    double* const w = sh->scr;
    double* const rav = w, * const vrat = w + 3, * const qc1 = w + 6, * const qc2 = w + 9, * const qc4 = w + 12;
    double* const qc5 = w + 15, * const ssmax = w + 18, * const smax = w + 21, * const qkr = w + 24;
    double* const rrc = w + 27, * const sqmx = w + 30, * const qjk = w + 33, * const sigsq = w + 42;
#define PB_FOR_K(k) for (int k = 1 + g.lane(); k <= 3; k += G)
    bool bad = false;
    PB_FOR_K(k) {
        double smink = ss[1][k], smaxk = 0.0;
        PB_UNROLL
        for (int i = 1; i <= 5; ++i) {
            smink = fmin(smink, ss[i][k]);
            smaxk = fmax(smaxk, ss[i][k]);
        }
        if (smink < cv::TINY * smaxk) bad = true;
        smax[k - 1] = smaxk;
        ssmax[k - 1] = smaxk * smaxk;
        double sumrat = 0.0, sumrsq = 0.0;
        PB_UNROLL
        for (int i = 1; i <= 4; ++i) {
            const double rat = pb_div(ss[i][k], ss[i + 1][k]);
            sumrat += rat;
            sumrsq += rat * rat;
        }
        const double ravk = 0.25 * sumrat;
        rav[k - 1] = ravk;
        vrat[k - 1] = fabs(0.25 * sumrsq - ravk * ravk);
        qc5[k - 1] = ss[1][k] * ss[3][k] - ss[2][k] * ss[2][k];
        qc4[k - 1] = ss[2][k] * ss[3][k] - ss[1][k] * ss[4][k];
        qc2[k - 1] = ss[2][k] * ss[5][k] - ss[3][k] * ss[4][k];
        qc1[k - 1] = ss[4][k] * ss[4][k] - ss[3][k] * ss[5][k];
    }
    if (g.any(bad)) return -1;  // (also orders the scratch writes before the reads below)
    g.sync();
    int kflag = 0;
    double rr = 0.0;
    const double vmin = fmin(vrat[0], fmin(vrat[1], vrat[2]));
    const double vmax = fmax(vrat[0], fmax(vrat[1], vrat[2]));
    if (vmin < vrrtol * vrrtol) {
        if (vmax > vrrt2 * vrrt2) return -2;
        rr = pb_div(rav[0] + rav[1] + rav[2], 3.0);
        double drrmax = 0.0;
        PB_UNROLL
        for (int k = 0; k < 3; ++k) drrmax = fmax(drrmax, fabs(rav[k] - rr));
        if (drrmax > vrrt2) return -3;
        kflag = 1;
    } else {
        // Gaussian elimination on the 3 x 4 system of the quartics' coefficients (rows i = 1,2,4,5
        // of qc; row 3 is zero), uniform
        if (fabs(qc1[0]) < cv::TINY * ssmax[0]) return -4;
        double tem = pb_div(qc1[1], qc1[0]);
        const double q22 = qc2[1] - tem * qc2[0];
        double q32 = 0.0 - tem * 0.0;
        const double q42 = qc4[1] - tem * qc4[0], q52 = qc5[1] - tem * qc5[0];
        tem = pb_div(qc1[2], qc1[0]);
        double q23 = qc2[2] - tem * qc2[0];
        double q33 = 0.0 - tem * 0.0;
        double q43 = qc4[2] - tem * qc4[0], q53 = qc5[2] - tem * qc5[0];
        if (fabs(q22) < cv::TINY * ssmax[1]) return -4;
        tem = pb_div(q23, q22);
        q33 -= tem * q32;
        q43 -= tem * q42;
        q53 -= tem * q52;
        (void)q33;
        if (fabs(q43) < cv::TINY * ssmax[2]) return -4;
        rr = pb_div(-q53, q43);
        if (rr < cv::TINY || rr > 100.0) return -5;
        double sq_own = 0.0;
        PB_FOR_K(k) {
            const double v = qc5[k - 1] + rr * (qc4[k - 1] + rr * rr * (qc2[k - 1] + rr * qc1[k - 1]));
            qkr[k - 1] = v;
            sq_own = fmax(sq_own, pb_div(fabs(v), ssmax[k - 1]));
        }
        const double sqmax = g.max(sq_own);
        double sqmin = 0.0;
        if (sqmax < sqtol) {
            kflag = 2;
        } else {
            PB_ROLL
            for (int it = 1; it <= 3; ++it) {
                g.sync();
                PB_FOR_K(k) {
                    const double qp = qc4[k - 1] + rr * rr * (3.0 * qc2[k - 1] + rr * 4.0 * qc1[k - 1]);
                    double drr = 0.0;
                    if (fabs(qp) > cv::TINY * ssmax[k - 1]) drr = pb_div(-qkr[k - 1], qp);
                    const double s = rr + drr;
                    rrc[k - 1] = s;
                    double sqmaxk = 0.0;
                    PB_UNROLL
                    for (int j = 0; j < 3; ++j) {
                        const double v = qc5[j] + s * (qc4[j] + s * s * (qc2[j] + s * qc1[j]));
                        qjk[(k - 1) * 3 + j] = v;
                        const double saqj = pb_div(fabs(v), ssmax[j]);
                        if (saqj > sqmaxk) sqmaxk = saqj;
                    }
                    sqmx[k - 1] = sqmaxk;
                }
                g.sync();
                int kmin = 0;
                sqmin = sqmx[0] + 1.0;
                PB_UNROLL
                for (int k = 0; k < 3; ++k) {
                    if (sqmx[k] < sqmin) {
                        kmin = k;
                        sqmin = sqmx[k];
                    }
                }
                rr = rrc[kmin];
                if (sqmin < sqtol) {
                    kflag = 3;
                    break;
                }
                double nq = 0.0;
                PB_FOR_K(k) nq = qjk[kmin * 3 + (k - 1)];
                g.sync();
                PB_FOR_K(k) qkr[k - 1] = nq;
            }
            if (sqmin > sqtol) return -6;
        }
    }
    g.sync();
    bad = false;
    PB_FOR_K(k) {
        const double rsa = ss[1][k];
        const double rsb = ss[2][k] * rr;
        const double rsc = ss[3][k] * rr * rr;
        const double rsd = ss[4][k] * rr * rr * rr;
        const double rd1a = rsa - rsb, rd1b = rsb - rsc, rd1c = rsc - rsd;
        const double rd2a = rd1a - rd1b, rd2b = rd1b - rd1c;
        const double rd3a = rd2a - rd2b;
        if (fabs(rd1b) < cv::TINY * smax[k - 1]) bad = true;
        const double cest1 = pb_div(-rd3a, rd1b);
        if (cest1 < cv::TINY || cest1 > 4.0) bad = true;
        const double corr1 = pb_div(pb_div(rd2b, cest1), rr * rr);
        sigsq[k - 1] = ss[3][k] + corr1;
    }
    if (g.any(bad)) return -7;
    g.sync();
#undef PB_FOR_K
    if (sigsq[1] < cv::TINY) return -8;
    const double ratp = pb_div(sigsq[2], sigsq[1]);
    const double ratm = pb_div(sigsq[0], sigsq[1]);
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


// ---- out-of-line group kernels: static functions of (shared working set, lane group, scalars)
//      only -- a noinline MEMBER function would force the integrator object into local memory ----
/// position of flat element e in the norm scratch (column stride NXS)
template <class M>
PB_HD int scr_pos(int e) { return (GroupShared<M>::NXS == M::NX) ? e : e + e / M::NX; }

/// sqrt(max_c(sum_i scr[c][i]) / NX) over the columns [cmin, cmax] of the norm scratch: the
/// WRMS norm (N_VWrmsNorm for the state column; cvSensNorm / cvSensUpdateNorm take the max
/// over the columns, cvodes.c:9427-9457).  One lane sums a column in CVODES' order.
template <class M, class Grp>
PB_HD_NOINLINE static double colnorm_shared(GroupShared<M>* sh, const Grp g, int cmin, int cmax) {
    constexpr int NX = M::NX, NXS = GroupShared<M>::NXS, G = Grp::G;
    PB_ASSUME_SHARED(sh);
    g.sync();
    double loc = 0.0;
    for (int c = cmin + g.lane(); c <= cmax; c += G) {
        const double* col = sh->scr + c * NXS;
        double s = 0.0;
        PB_UNROLL
        for (int i = 0; i < NX; ++i) s += col[i];
        if (s > loc) loc = s;
    }
    const double m = g.max(loc);
    g.sync();  // the scratch may be overwritten by its next user
    // mean over the NX entries: a power-of-two NX divides exactly, so the multiplication is the division
    if ((NX & (NX - 1)) == 0) return sqrt(m * (1.0 / (double)NX));
    return sqrt(pb_div(m, (double)NX));
}
/// WRMS norm of A over the columns [cmin, cmax]
template <class M, class Grp>
PB_HD_NOINLINE static double wnorm1_shared(GroupShared<M>* sh, const Grp g, int cmin, int cmax, const double* A) {
    constexpr int NX = M::NX, G = Grp::G;
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(A);
    for (int e = cmin * NX + g.lane(); e < (cmax + 1) * NX; e += G) {
        const double prod = A[e] * sh->ewt[e];
        sh->scr[scr_pos<M>(e)] = prod * prod;
    }
    return colnorm_shared<M, Grp>(sh, g, cmin, cmax);
}
/// WRMS norm of ca*A + cb*B
template <class M, class Grp>
PB_HD_NOINLINE static double wnorm2_shared(GroupShared<M>* sh, const Grp g, int cmin, int cmax, const double* A,
                                           double ca, const double* B, double cb) {
    constexpr int NX = M::NX, G = Grp::G;
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(A);
    PB_ASSUME_SHARED(B);
    for (int e = cmin * NX + g.lane(); e < (cmax + 1) * NX; e += G) {
        const double prod = (ca * A[e] + cb * B[e]) * sh->ewt[e];
        sh->scr[scr_pos<M>(e)] = prod * prod;
    }
    return colnorm_shared<M, Grp>(sh, g, cmin, cmax);
}

/// Ainv = (I - gamma*Jsaved)^-1 by in-place Gauss-Jordan elimination with partial pivoting
/// (the pivot order of SUNDlsMat_denseGETRF, sundials_dense.c:82-148), one matrix column per
/// lane.  Returns 0, or 1 on a zero pivot (group-uniform).
template <class M, class Grp>
PB_HD_NOINLINE static int invert_newton_matrix(GroupShared<M>* sh, const Grp g, const double gamma) {
constexpr int NX = M::NX, G = Grp::G, RPL = (NX + G - 1) / G;
    double* A = sh->Ainv;
    for (int e = g.lane(); e < NX * NX; e += G) {
        const int r = e % NX, c = e / NX;
        A[e] = (r == c ? 1.0 : 0.0) - gamma * sh->Jsaved[e];
    }
    for (int i = g.lane(); i < NX; i += G) sh->perm[i] = i;
    g.sync();
    int fail = 0;
    PB_ROLL
    for (int k = 0; k < NX; ++k) {
        // every lane: column k, pivot row l (first largest magnitude among rows >= k)
        double colk[NX];
        PB_UNROLL
        for (int i = 0; i < NX; ++i) colk[i] = A[i + k * NX];
        int lp = k;
        double pv = A[k + k * NX];
        PB_UNROLL
        for (int i = 1; i < NX; ++i) {
            if (i > k && fabs(colk[i]) > fabs(pv)) { pv = colk[i]; lp = i; }
        }
        if (pv == 0.0) { fail = 1; break; }
        const double pinv = pb_rcp(pv);
        // column k after the row swap k <-> lp
        const double ck = A[k + k * NX];
        PB_UNROLL
        for (int i = 1; i < NX; ++i)
            if (i == lp && i > k) colk[i] = ck;
        g.sync();  // every lane has read column k
        for (int j = g.lane(); j < NX; j += G) {
            double* cj = A + j * NX;
            if (lp != k) {
                const double t_ = cj[lp];
                cj[lp] = cj[k];
                cj[k] = t_;
            }
            if (j == k) {
                PB_UNROLL
                for (int i = 0; i < NX; ++i) cj[i] = (i == k) ? pinv : -colk[i] * pinv;
            } else {
                const double akj = cj[k] * pinv;
                PB_UNROLL
                for (int i = 0; i < NX; ++i) {
                    if (i == k) cj[i] = akj;
                    else cj[i] -= colk[i] * akj;
                }
            }
        }
        if (g.lane() == 0 && lp != k) {
            const int t_ = sh->perm[k];
            sh->perm[k] = sh->perm[lp];
            sh->perm[lp] = t_;
        }
        g.sync();
    }
    if (fail) return 1;
    // A holds (P (I - gamma J))^-1 = (I - gamma J)^-1 P^T: its column j belongs at perm[j]
    double mine[RPL][NX];
    PB_UNROLL
    for (int r = 0; r < RPL; ++r) {
        const int j = g.lane() + r * G;
        if (j < NX) {
            PB_UNROLL
            for (int i = 0; i < NX; ++i) mine[r][i] = A[i + j * NX];
        }
    }
    g.sync();
    PB_UNROLL
    for (int r = 0; r < RPL; ++r) {
        const int j = g.lane() + r * G;
        if (j < NX) {
            double* dst = A + sh->perm[j] * NX;
            PB_UNROLL
            for (int i = 0; i < NX; ++i) dst[i] = mine[r][i];
        }
    }
    g.sync();
    return 0;
}


/// cvAdjustOrder -> cvIncreaseBDF / cvDecreaseBDF (cvodes.c:6134-6143, 6256-6351, 6363-6415)
template <class M, class Grp>
PB_HD_NOINLINE static void adjust_order_shared(GroupShared<M>* sh, const Grp g, int ncol, int q, int L, double hscale,
                                            int deltaq) {
constexpr int NX = M::NX, G = Grp::G;
PB_ASSUME_SHARED(sh);
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
            alpha0 -= pb_inv_int(j + 1);
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
    for (int e = g.lane(); e < NX * ncol; e += G) {
        if (deltaq == 1) sh->zn[src][e] = A1 * sh->zn[cv::Q_MAX][e];
        const double zs = sh->zn[src][e];
        for (int j = 2; j <= jmax; ++j) sh->zn[j][e] += lc[j] * zs;
    }
}


/// Convergence test of a simulation to steady state (SteadystateProblem::runSteadystateSimulation,
/// deps/AMICI/src/steadystateproblem.cpp:711-794, with getWrms :553-591, getWrmsNorm :519-551 and
/// getWrmsFSA :593-626): wrms(xdot(t, x)) < 1 with the weights 1 / (rtol*|x| + atol) and -- when
/// sensitivities are integrated -- wrms(J*sx_ip + dxdotdp_ip) < 1 with 1 / (rtol*|sx_ip| + atol)
/// for the parameters of plist in order, stopping at the first one above the threshold.  x / sx are
/// column 0 / the sensitivity columns of zn[0] (what the solver returns at tn).  The tolerances are
/// the steady-state ones (AMICI: 100 x the integration tolerances, solver.h:1908-1923).
/// Scratch: yj, ysh, fout, Jcur, dfdp, scr -- all rewritten by the next step attempt.  Out of line.
template <class M, class Grp>
PB_HD_NOINLINE static bool ss_converged_shared(GroupShared<M>* sh, const Grp g, double t, double rtol, double atol,
                                               int ns, int nsall, const double* xsrc) {
    constexpr int NX = M::NX, G = Grp::G;
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(xsrc);
    g.sync();
    // xsrc: the state the test is made at -- zn[0] (the solver's state), or the last trial point of a failed Newton
    // attempt before the first step of a post-equilibration (SteadystateProblem::state_.x is what getWrms reads)
    for (int i = g.lane(); i < NX; i += G) {
        const double v = xsrc[i];
        sh->yj[i] = v;
        sh->ysh[i] = v;
    }
    g.sync();
    int conv = 0;
    if (g.lane() == 0) {
        eval_xdot_shared<M>(sh, t);
        double s = 0.0;
        for (int i = 0; i < NX; ++i) {
            const double w = pb_rcp(rtol * fabs(sh->ysh[i]) + atol);
            const double prod = sh->fout[i] * w;
            s += prod * prod;
        }
        const double wrms = sqrt(pb_div(s, (double)NX));
        conv = (wrms < 1.0) ? 1 : 0;
    }
    conv = g.bcast0(conv);
    if (!conv || ns == 0) return conv != 0;
    // sensitivities: Model_ODE::fsxdot (model_ode.cpp:486-515) at (t, x)
    if (g.lane() == 0) {
        eval_jac_shared<M>(sh->Jcur, t, sh->ysh, sh);
        eval_dfdp_shared<M>(sh, t, sh->ysh);
    }
    g.sync();
    for (int c = 1 + g.lane(); c <= ns; c += G) {
        const double* dcol = sh->dfdp + sh->cdcol[c - 1] * NX;
        const double* sx = sh->zn[0] + c * NX;
        double s = 0.0;
        for (int i = 0; i < NX; ++i) {
            double F = dcol[i];
            for (int j = 0; j < NX; ++j) F += sh->Jcur[i + j * NX] * sx[j];
            const double w = pb_rcp(rtol * fabs(sx[i]) + atol);
            const double prod = F * w;
            s += prod * prod;
        }
        sh->scr[c] = sqrt(pb_div(s, (double)NX));
    }
    g.sync();
    if (g.lane() == 0) {
        double wrms = 0.0;
        for (int j = 0; j < nsall; ++j) {
            const int c = sh->colof[j];
            wrms = c > 0 ? sh->scr[c] : 0.0;  // a parameter the state does not depend on: sx == 0, sxdot == 0
            if (wrms > 1.0) break;
        }
        conv = (wrms < 1.0) ? 1 : 0;
    }
    conv = g.bcast0(conv);
    g.sync();
    return conv != 0;
}

/// J(t, x) delta = rhs for nrhs right-hand sides by LU with partial pivoting (the pivot order of
/// SUNDlsMat_denseGETRF), serially by the group's first lane: the linear system of the steady-state Newton solver
/// (NewtonSolver::prepareLinearSystem / solveLinearSystem, deps/AMICI/src/newton_solver.cpp:86-125; the reference
/// factorises with KLU by default).  x = sh->ysh; the factors overwrite sh->Jcur, which is zeroed again afterwards
/// (the Jacobian functions write structural non-zeros only).  rhs: nrhs columns of stride NX, solved in place.
/// Returns 1 if the matrix is singular or not finite.  Lane 0 only (callers synchronise).
template <class M>
PB_HD_NOINLINE static int ss_newton_linsolve_lane0(GroupShared<M>* sh, double t, double* rhs, int nrhs) {
    constexpr int NX = M::NX;
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(rhs);
    double* A = sh->Jcur;
    for (int e = 0; e < NX * NX; ++e) A[e] = 0.0;
    eval_jac_shared<M>(A, t, sh->ysh, sh);
    int piv[NX];
    int bad = 0;
    for (int k = 0; k < NX && !bad; ++k) {
        int l = k;
        for (int i = k + 1; i < NX; ++i)
            if (fabs(A[i + k * NX]) > fabs(A[l + k * NX])) l = i;
        piv[k] = l;
        const double d = A[l + k * NX];
        if (d == 0.0 || !PB_FINITE(d)) { bad = 1; break; }
        if (l != k)
            for (int j = 0; j < NX; ++j) {
                const double tmp = A[l + j * NX];
                A[l + j * NX] = A[k + j * NX];
                A[k + j * NX] = tmp;
            }
        const double mult = pb_rcp(A[k + k * NX]);
        for (int i = k + 1; i < NX; ++i) A[i + k * NX] *= mult;
        for (int j = k + 1; j < NX; ++j) {
            const double akj = A[k + j * NX];
            if (akj != 0.0)
                for (int i = k + 1; i < NX; ++i) A[i + j * NX] -= akj * A[i + k * NX];
        }
    }
    if (!bad) {
        for (int c = 0; c < nrhs; ++c) {
            double* b = rhs + c * NX;
            for (int k = 0; k < NX; ++k) {  // row permutation, then forward substitution (unit lower triangle)
                const int l = piv[k];
                if (l != k) { const double tmp = b[l]; b[l] = b[k]; b[k] = tmp; }
            }
            for (int k = 0; k < NX - 1; ++k)
                for (int i = k + 1; i < NX; ++i) b[i] -= b[k] * A[i + k * NX];
            for (int k = NX - 1; k >= 0; --k) {
                b[k] = pb_div(b[k], A[k + k * NX]);
                for (int i = 0; i < k; ++i) b[i] -= b[k] * A[i + k * NX];
            }
        }
    }
    for (int e = 0; e < NX * NX; ++e) A[e] = 0.0;
    return bad;
}

/// SteadystateProblem::applyNewtonsMethod (deps/AMICI/src/steadystateproblem.cpp:628-683) with getWrms (:553-591),
/// getNewtonStep (:889-896) and updateDampingFactor (:697-715; damping on, lower bound 1e-8: AMICI's defaults):
/// damped Newton iteration x <- x_old + gamma * delta, J(x_old) delta = -xdot(x_old), accepted while the weighted
/// RMS norm of xdot decreases, converged when it is below 1 at the steady-state tolerances.  x (NX values in
/// group-shared memory) is iterated in place and holds the LAST TRIAL POINT on return, as the reference's state_.x
/// does.  Scratch: yj, ysh (x_old), fout, b (delta), Jcur.  Executed by the group's first lane.
/// Returns the number of Newton steps taken (>= 0) on convergence, -1 otherwise (too many steps, damping factor at
/// its lower bound, singular Jacobian).
template <class M, class Grp>
PB_HD_NOINLINE static int ss_newton_shared(GroupShared<M>* sh, const Grp g, double t, double rtol, double atol,
                                           int max_steps, bool retry, double* x) {
    constexpr int NX = M::NX;
    PB_ASSUME_SHARED(sh);
    PB_ASSUME_SHARED(x);
    g.sync();
    int result = -1;
    if (g.lane() == 0) {
        double* xold = sh->ysh;
        double* delta = sh->b;
        auto wrms_at = [&](const double* xs) {  // getWrms: xdot at xs, weights 1 / (rtol |xs| + atol)
            for (int i = 0; i < NX; ++i) sh->yj[i] = xs[i];
            eval_xdot_shared<M>(sh, t);
            double s = 0.0;
            for (int i = 0; i < NX; ++i) {
                const double prod = sh->fout[i] * pb_rcp(rtol * fabs(xs[i]) + atol);
                s += prod * prod;
            }
            return sqrt(pb_div(s, (double)NX));
        };
        for (int i = 0; i < NX; ++i) xold[i] = x[i];
        double wrms = wrms_at(x);
        bool converged = retry ? false : (wrms < 1.0);
        bool update_direction = true, failed = false;
        double gamma = 1.0;
        int it = 0;
        while (!converged && it < max_steps) {
            if (update_direction) {
                // xdot at the accepted state (x == xold whenever a new direction is taken)
                for (int i = 0; i < NX; ++i) sh->yj[i] = xold[i];
                eval_xdot_shared<M>(sh, t);
                for (int i = 0; i < NX; ++i) delta[i] = -sh->fout[i];
                if (ss_newton_linsolve_lane0<M>(sh, t, delta, 1)) { failed = true; break; }
            }
            for (int i = 0; i < NX; ++i) x[i] = xold[i] + gamma * delta[i];
            const double wrms_tmp = wrms_at(x);
            const bool ok = wrms_tmp < wrms;
            if (ok) {
                wrms = wrms_tmp;
                converged = wrms < 1.0;
                for (int i = 0; i < NX; ++i) xold[i] = x[i];
            }
            gamma = ok ? fmin(1.0, 2.0 * gamma) : gamma / 4.0;
            if (gamma < 1e-8) { failed = true; break; }
            update_direction = ok;
            ++it;
        }
        result = (converged && !failed) ? it : -1;
    }
    result = g.bcast0(result);
    g.sync();
    return result;
}

/// NewtonSolver::computeNewtonSensis (newton_solver.cpp:52-84): sx_ip = -J(x)^-1 dxdotdp_ip at the steady state x
/// (column 0 of zn[0]) into the sensitivity columns of zn[0].  Returns 1 if the Jacobian is singular.
template <class M, class Grp>
PB_HD_NOINLINE static int ss_newton_sensis_shared(GroupShared<M>* sh, const Grp g, double t, int ns) {
    constexpr int NX = M::NX;
    PB_ASSUME_SHARED(sh);
    g.sync();
    int bad = 0;
    if (g.lane() == 0) {
        for (int i = 0; i < NX; ++i) sh->ysh[i] = sh->zn[0][i];
        eval_dfdp_shared<M>(sh, t, sh->ysh);
        for (int c = 1; c <= ns; ++c) {
            const double* dcol = sh->dfdp + sh->cdcol[c - 1] * NX;
            for (int i = 0; i < NX; ++i) sh->zn[0][c * NX + i] = -dcol[i];
        }
        bad = ss_newton_linsolve_lane0<M>(sh, t, sh->zn[0] + NX, ns);
    }
    bad = g.bcast0(bad);
    g.sync();
    return bad;
}

/// SS: with support for steady states -- output times t = inf (post-equilibration) and a
/// pre-equilibration phase (begin() modes 1 / 2).  The SS = false instantiation is the code without
/// any of it (an infinite output time is an error there).
template <class M, class Grp, bool SS = false>
struct Bdf {
    static constexpr int NX = M::NX, NP = M::NP, NK = M::NK, NY = M::NY;
    static constexpr int G = Grp::G;
    static_assert(!SS || NX <= 32, "steady-state support uses a 32-bit mask of re-initialised states");
    // 2-D lane mapping of the row-parallel kernels: the G lanes of a group are LR "row lanes"
    // (the power of two >= NX, at most G) times CG "column groups"; lane l works on the rows
    // (l % LR) + r*LR, r < RPL, of the columns c = cmin + (l / LR), + CG, ...
    static constexpr int NXP2 = NX <= 1 ? 1 : NX <= 2 ? 2 : NX <= 4 ? 4 : NX <= 8 ? 8 : NX <= 16 ? 16 : 32;
    static constexpr int LR = G < NXP2 ? G : NXP2;
    static constexpr int CG = G / LR;
    static constexpr int RPL = (NX + LR - 1) / LR;  // rows (state indices) per lane
    PB_HD int row_lane() const { return CG == 1 ? g.lane() : g.lane() % LR; }
    PB_HD int col_group() const { return CG == 1 ? 0 : g.lane() / LR; }
    static constexpr int NXS = GroupShared<M>::NXS;
    enum { PH_STATE = 0, PH_SENS = 1 };

    Grp g;
    GroupShared<M>* sh;  // host harness: the working set; device: unused, see S()
    /// The group's working set.  On the device it is re-derived from the dynamic shared memory
    /// base wherever it is used, so that the compiler always knows the address space (a pointer
    /// that has been spilled and reloaded is a generic pointer: LD/ST instead of LDS/STS).
    PB_HD GroupShared<M>* S() const {
#if defined(__CUDA_ARCH__)
        extern __shared__ __align__(16) unsigned char pb_smem_raw[];
        return reinterpret_cast<GroupShared<M>*>(pb_smem_raw) + (threadIdx.x / G);
#else
        return sh;
#endif
    }
    SolverOptions opt;
    int ns;      // integrated sensitivity columns (columns 1..ns; column 0 is the state)
    int nsall;   // plist entries of the condition (ns of them integrated, the others identically zero)
    int ncol;    // 1 + ns
    bool sensi;  // ns > 0

    // integrator scalars (uniform within the group)
    double tn, h, hprime, hscale, eta, etamax, etaq, etaqm1, etaqp1;
    double gamma, gammap, gamrat, rl1, crate, crateS, delp, acnrm, acnrmS, saved_tq5, tstop;
    double l[cv::L_MAX];          // BDF polynomial coefficients l[0..q] (compile-time indices only)
    double tq1, tq2, tq3, tq4, tq5;  // test quantities
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
    // stage of the attempt the state machine is in (stage_pre / stage_nls / stage_post)
    enum { ST_PRE = 0, ST_NLS = 1, ST_DONE = 2 };
    int stage_, fail_, ret_;
    // steady-state support (SS only): what the simulation is doing
    enum { SS_OFF = 0, SS_PREEQ = 1, SS_POSTEQ = 2 };
    int ss_mode_;
    int ss_newton_;  // Newton attempts of the current steady-state phase: 0 none yet, 1 first done, 2 retry done,
                     // 3 retry requested (the simulation has just failed)
    // steadyStateSensitivityMode newtonOnly: the sensitivity columns are switched off for the current steady-state
    // phase (Solver::switchForwardSensisOff -> CVodeSensToggleOff); ns_full_ = their number
    bool sens_off_;
    int ns_full_;
    PB_HD void sens_toggle_off() {
        if (!SS || !opt.ss_sens_by_solve() || !sensi) return;
        ns_full_ = ns;
        ns = 0;
        ncol = 1;
        sensi = false;
        sens_off_ = true;
    }
    /// true if the columns had been off (and are integrated / evaluated again from here on)
    PB_HD bool sens_back_on() {
        if (!SS || !sens_off_) return false;
        ns = ns_full_;
        ncol = ns + 1;
        sensi = true;
        sens_off_ = false;
        return true;
    }
    /// true while CVode is driven in CV_ONE_STEP mode without a stop time (Solver::step,
    /// deps/AMICI/src/solver.cpp:161-175): simulation to steady state
    PB_HD bool one_step() const { return SS && ss_mode_ != SS_OFF; }
    /// AMICI always sets a stop time in CV_NORMAL mode; it is absent only in one-step mode
    PB_HD bool tstop_active() const { return SS ? tstopset : true; }

// columns of this lane in [cmin, cmax] (column-parallel cold paths: initial values, data points)
#define PB_FOR_COLS(c, cmin, cmax) \
    for (int c = g.lane(); c <= (cmax); c += G) \
        if (c >= (cmin))
// flat element range [e0, e1) of the column data: every lane busy
#define PB_FOR_RANGE(e, e0, e1) for (int e = (e0) + g.lane(); e < (e1); e += G)
// all elements of the active columns
#define PB_FOR_ELEMS(e) PB_FOR_RANGE(e, 0, NX * ncol)
// the rows (state indices) owned by this lane: i = lane + r*G
#define PB_FOR_ROWS(r, i) \
    PB_UNROLL                \
    for (int r = 0; r < RPL; ++r) \
        for (int i = row_lane() + r * LR, once_ = 1; once_ && i < NX; once_ = 0)
// the same, first column group only (single-column data)
#define PB_FOR_ROWS0(r, i) \
    if (col_group() == 0) PB_FOR_ROWS(r, i)

    PB_HD int phase_cmin(int phase) const { return phase == PH_STATE ? 0 : 1; }
    PB_HD int phase_cmax(int phase) const { return phase == PH_STATE ? 0 : ns; }
    /// columns in the error test: state, plus sensitivities if errconS
    PB_HD int errtest_cmax() const { return (sensi && opt.errconS) ? ns : 0; }
    /// position of flat element e in the norm scratch (column stride NXS)

    // ---------------------------------------------------------------- norms
    PB_HD double wnorm1(int cmin, int cmax, const double* A) const { return wnorm1_shared<M, Grp>(S(), g, cmin, cmax, A); }
    PB_HD double wnorm2(int cmin, int cmax, const double* A, double ca, const double* B, double cb) const {
        return wnorm2_shared<M, Grp>(S(), g, cmin, cmax, A, ca, B, cb);
    }
    PB_HD double colnorm(int cmin, int cmax) const { return colnorm_shared<M, Grp>(S(), g, cmin, cmax); }

    /// cvEwtSetSS / cvSensEwtSetSS (cvodes.c:9118-9129, 9263-9279)
    PB_HD bool ewt_set() {
        bool ok = true;
        PB_FOR_ELEMS(e) {
            const double w = opt.rtol * fabs(S()->zn[0][e]) + opt.atol;
            ok = ok && (w > 0.0);
            S()->ewt[e] = 1.0 / w;
        }
        return !g.any(!ok);
    }

    /// J(t, ysh) and d xdot / d p -> S()->Jcur, S()->dfdp by the first lane (Model_ODE::fsxdot at
    /// ip == 0, deps/AMICI/src/model_ode.cpp:492-498)
    PB_HD void refresh_Jcur(double t) {
        g.sync();
        if (g.lane() == 0) {
            eval_jac_shared<M>(S()->Jcur, t, S()->ysh, S());
            eval_dfdp_shared<M>(S(), t, S()->ysh);
        }
        g.sync();
    }

    // ---------------------------------------------------------------- right-hand sides
    /// THE right-hand-side evaluation.  For the columns of `phase`:  v = zn[0] + acor;
    /// F = f(t, v)  (state: xdot; sensitivities: dxdotdp(ip) + Jcur*v, Model_ODE::fsxdot
    /// model_ode.cpp:486-515, with S()->ysh / S()->Jcur / S()->dfdp prepared by the caller);
    /// b = raw ? F : rl1*zn[1] + acor - gamma*F  (cvNlsResidual cvodes_nls.c:344-371;
    /// cvNlsResidualSensStg cvodes_nls_stg.c:377-421).  Row-parallel: every lane produces its
    /// rows of every column.  Returns 0, RHSFUNC_RECVR or SRHSFUNC_RECVR (non-finite result).
    PB_HD int rhs(int phase, double t, bool raw) {
        bool bad = false;
        g.sync();  // element-wise writers of zn / acor may have been other lanes
        if (phase == PH_STATE) {
            PB_FOR_ROWS0(r, i) S()->yj[i] = S()->zn[0][i] + S()->acor[i];
            g.sync();
            int ok = 1;
            if (g.lane() == 0) ok = eval_xdot_shared<M>(S(), t) ? 1 : 0;
            g.sync();
            bad = g.bcast0(ok) == 0;
            PB_FOR_ROWS0(r, i) {
                const double F = S()->fout[i];
                S()->b[i] = raw ? F : (rl1 * S()->zn[1][i] + S()->acor[i]) - gamma * F;
            }
            nfe++;
        } else {
            // corrected sensitivities of all columns -> scratch (read by every lane below)
            PB_FOR_RANGE(e, NX, NX * ncol) S()->scr[e] = S()->zn[0][e] + S()->acor[e];
            double jrow[RPL][NX];
            PB_FOR_ROWS(r, i) {
                PB_UNROLL
                for (int j = 0; j < NX; ++j) jrow[r][j] = S()->Jcur[i + j * NX];
            }
            g.sync();
            PB_UNROLL_COLS
            for (int c = 1 + col_group(); c <= ns; c += CG) {
                const double* dcol = S()->dfdp + S()->cdcol[c - 1] * NX;
                const double* sc = S()->scr + c * NX;
                double sv[NX];
                PB_UNROLL
                for (int j = 0; j < NX; ++j) sv[j] = sc[j];
                PB_FOR_ROWS(r, i) {
                    double F = dcol[i];
                    PB_UNROLL
                    for (int j = 0; j < NX; ++j) F += jrow[r][j] * sv[j];
                    if (!isfinite(F)) bad = true;
                    const int e = c * NX + i;
                    S()->b[e] = raw ? F : (rl1 * S()->zn[1][e] + S()->acor[e]) - gamma * F;
                }
            }
            nfSe += nsall;  // counted as the reference does: one evaluation per plist entry
        }
        g.sync();
        if (g.any(bad)) return phase == PH_STATE ? cv::RHSFUNC_RECVR : cv::SRHSFUNC_RECVR;
        return 0;
    }

    /// S()->ysh = zn[0] + acor of the state column (the corrected state y_n)
    PB_HD void publish_state() {
        g.sync();
        PB_FOR_ROWS0(r, i) S()->ysh[i] = S()->zn[0][i] + S()->acor[i];
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
    /// cvLsSetup + cvLsLinSys (cvodes_ls.c:1611-1708,1386-1470) + the inversion; the state the
    /// Jacobian is evaluated at is S()->yj.
    PB_HD int lsetup() {
        const double dgamma = fabs(pb_div(gamma, gammap) - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + cv::MSBJ) ||
                          ((convfail == cv::FAIL_BAD_J) && (dgamma < cv::DGMAX_JBAD)) ||
                          (convfail == cv::FAIL_OTHER);
        int ret = 0;
        g.sync();
        if (jbad) {
            if (g.lane() == 0) eval_jac_shared<M>(S()->Jsaved, tn, S()->yj, S());
            g.sync();
            // AMICI's fJ callback returns checkFinite(J): recoverable if not finite
            bool bad = false;
            for (int e = g.lane(); e < NX * NX; e += G)
                if (!isfinite(S()->Jsaved[e])) bad = true;
            if (g.any(bad)) ret = 1;
        }
        if (ret == 0) ret = invert_newton_matrix<M, Grp>(S(), g, gamma);
        if (jbad) { nje++; nstlj = nst; }
        jcur = jbad;
        nsetups++;
        return ret;
    }

    /// Newton update of the columns [cmin, cmax]: delta = -(I - gamma J)^-1 b (dense GETRS
    /// sundials_dense.c:150-182 as a mat-vec with the inverse) with cvLsSolve's gamma-ratio
    /// scaling (cvodes_ls.c:1849-1854); acor += delta; returns the WRMS norm of delta.
    PB_HD double newton_update(int cmin, int cmax) {
        double arow[RPL][NX];
        PB_FOR_ROWS(r, i) {
            PB_UNROLL
            for (int j = 0; j < NX; ++j) arow[r][j] = S()->Ainv[i + j * NX];
        }
        const double nsc = (gamrat != 1.0) ? pb_div(-2.0, 1.0 + gamrat) : -1.0;
        PB_UNROLL_COLS
        for (int c = cmin + col_group(); c <= cmax; c += CG) {
            const double* bc = S()->b + c * NX;
            double bv[NX];
            PB_UNROLL
            for (int j = 0; j < NX; ++j) bv[j] = bc[j];
            PB_FOR_ROWS(r, i) {
                double acc = 0.0;
                PB_UNROLL
                for (int j = 0; j < NX; ++j) acc += arow[r][j] * bv[j];
                const double x = nsc * acc;
                const int e = c * NX + i;
                S()->acor[e] += x;
                const double prod = x * S()->ewt[e];
                S()->scr[c * NXS + i] = prod * prod;
            }
        }
        return colnorm(cmin, cmax);
    }

    // ---------------------------------------------------------------- BDF coefficients
    /// cvSetBDF + cvSetTqBDF + cvSet (cvodes.c:6788-6879, 6617-6629); computed redundantly by
    /// every lane, fully unrolled so that l[] stays in registers.
    PB_HD void cvset() {
        double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
        l[0] = l[1] = xi_inv = xistar_inv = 1.0;
        PB_UNROLL
        for (int i = 2; i < cv::L_MAX; ++i) l[i] = 0.0;
        alpha0 = alpha0_hat = -1.0;
        hsum = h;
        if (q > 1) {
            PB_UNROLL
            for (int j = 2; j < cv::Q_MAX; ++j) {
                if (j < q) {
                    hsum += S()->tau[j - 1];
                    xi_inv = pb_div(h, hsum);
                    alpha0 -= 1.0 / (double)j;
                    PB_UNROLL
                    for (int i = j; i >= 1; --i) l[i] += l[i - 1] * xi_inv;
                }
            }
            alpha0 -= pb_inv_int(q);
            xistar_inv = -l[1] - alpha0;
            hsum += S()->tau[q - 1];
            xi_inv = pb_div(h, hsum);
            alpha0_hat = -l[1] - xi_inv;
            PB_UNROLL
            for (int i = cv::Q_MAX; i >= 1; --i)
                if (i <= q) l[i] += l[i - 1] * xistar_inv;
        }
        // l[q] with compile-time indices only (written out: inside `if (i == q)` of a loop the
        // optimiser rewrites l[i] as l[q], and one dynamic index puts the whole integrator object
        // into local memory)
        double lq = l[1];
        if (q == 2) lq = l[2];
        if (q == 3) lq = l[3];
        if (q == 4) lq = l[4];
        if (q == 5) lq = l[5];
        static_assert(cv::Q_MAX == 5, "l[q] selection is written out for q <= 5");
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        tq2 = fabs(pb_div(A1, alpha0 * A2));
        tq5 = fabs(pb_div(A2 * xistar_inv, lq * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double C = pb_div(xistar_inv, lq);
                const double A3 = alpha0 + pb_inv_int(q);
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = pb_div(1.0 - A4 + A3, A3);
                tq1 = fabs(C * Cpinv);
            } else {
                tq1 = 1.0;
            }
            hsum += S()->tau[q];
            xi_inv = pb_div(h, hsum);
            const double A5 = alpha0 - pb_inv_int(q + 1);
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = pb_div(1.0 - A6 + A5, A2);
            tq3 = fabs(pb_div(Cppinv, xi_inv * (q + 2) * A5));
        }
        tq4 = pb_div(cv::CORTES, tq2);
        rl1 = pb_rcp(l[1]);
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
            for (int j = 0; j < cv::L_MAX; ++j) z[j] = (j <= q) ? S()->zn[j][e] : 0.0;
            PB_UNROLL
            for (int k = 1; k <= cv::Q_MAX; ++k) {
                PB_UNROLL
                for (int j = cv::Q_MAX; j >= k; --j) {
                    if (j <= q && k <= q) z[j - 1] += sgn * z[j];
                }
            }
            PB_UNROLL
            for (int j = 0; j < cv::Q_MAX; ++j)
                if (j < q) S()->zn[j][e] = z[j];
        }
    }
    /// cvRescale (cvodes.c:6425-6495)
    PB_HD void rescale() {
        PB_FOR_ELEMS(e) {
            double factor = eta;
            for (int j = 1; j <= q; ++j) {
                S()->zn[j][e] *= factor;
                factor *= eta;
            }
        }
        h = hscale * eta;
        hscale = h;
        nscon = 0;
    }
    /// cvAdjustOrder (out of line, one copy)
    PB_HD void adjust_order(int deltaq) { adjust_order_shared<M, Grp>(S(), g, ncol, q, L, hscale, deltaq); }

    // ---------------------------------------------------------------- nonlinear solve
    /// cvNls / cvStgrNls + SUNNonlinSolSolve_Newton + cvNlsConvTest{,SensStg} for the columns of
    /// `phase` (cvodes.c:6894-7007, 7167-7210; sunnonlinsol_newton.c:176-335; cvodes_nls.c:
    /// 298-342; cvodes_nls_stg.c:238-375).  The state corrector and the staggered sensitivity
    /// corrector are the same Newton iteration on different columns with the same inverse, so
    /// they share one copy of the code.  On success the corrected vector is zn[0]+acor.
    PB_HD int nls(int phase, int nflag) {
        const int cmin = phase_cmin(phase), cmax = phase_cmax(phase);
        bool setup_pending = false;
        if (phase == PH_STATE) {
            convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::NO_FAILURES : cv::FAIL_OTHER;
            setup_pending = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                            (nst >= nstlp + cv::MSBP) || (fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
        }
        PB_FOR_RANGE(e, cmin * NX, (cmax + 1) * NX) S()->acor[e] = 0.0;
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
                    PB_FOR_ROWS0(r, i) S()->yj[i] = S()->ysh[i];
                }  // state phase: rhs() has just put zn[0]+acor into S()->yj
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
                const double del = newton_update(cmin, cmax);
                double cr = (phase == PH_STATE) ? crate : crateS;
                if (m_it > 0) cr = fmax(cv::CRDOWN * cr, pb_div(del, delp));
                if (phase == PH_STATE) crate = cr;
                else crateS = cr;
                const double dcon = pb_div(del * fmin(1.0, cr), tq4);
                if (dcon <= 1.0) {
                    const double an = (m_it == 0) ? del : wnorm1(cmin, cmax, S()->acor);
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
                g.sync();
                PB_FOR_RANGE(e, cmin * NX, (cmax + 1) * NX) S()->acor[e] = 0.0;
                continue;
            }
            return ret;
        }
    }

    /// The simultaneous corrector (steady-state instantiation only): cvNls with do_sensi_sim (cvodes.c:6894-7007),
    /// SUNNonlinSolSolve_Newton on the wrapped vector (state, sensitivities) with cvNlsResidualSensSim /
    /// cvNlsLSolveSensSim / cvNlsConvTestSensSim (cvodes_nls_sim.c:374-560): every iteration evaluates f at the
    /// current state iterate and the sensitivity right-hand sides at the current (state, sensitivity) iterates --
    /// AMICI's fsxdot re-evaluates J and dxdotdp there (model_ode.cpp:486-515) --, solves all columns with the same
    /// matrix, and tests the MAXIMUM of the columns' WRMS norms (cvSensUpdateNorm); one convergence rate, one
    /// iteration counter (nni), acnrm = the joint norm when the sensitivities take part in the error test.
    PB_HD int nls_simultaneous(int nflag) {
        const int cmax = ns;
        convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::NO_FAILURES : cv::FAIL_OTHER;
        bool setup_pending = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                             (nst >= nstlp + cv::MSBP) || (fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
        PB_FOR_RANGE(e, 0, (cmax + 1) * NX) S()->acor[e] = 0.0;
        bool jbad = false;
        bool top = true;
        int m_it = 0;
        for (;;) {
            int ret = rhs(PH_STATE, tn, false);  // (S()->yj = the state iterate)
            if (ret == 0) {
                g.sync();
                PB_FOR_ROWS0(r, i) S()->ysh[i] = S()->yj[i];
                refresh_Jcur(tn);
                ret = rhs(PH_SENS, tn, false);
            }
            if (ret != 0 && top) return ret;
            if (ret == 0 && setup_pending) {
                if (jbad) convfail = cv::FAIL_BAD_J;
                const int lr = lsetup();  // Jacobian at S()->yj
                setup_pending = false;
                jcurNLS = jcur;
                gamrat = 1.0;
                gammap = gamma;
                crate = 1.0;
                crateS = 1.0;
                nstlp = nst;
                if (lr != 0) return cv::CONV_RECVR;
            }
            top = false;
            if (ret == 0) {
                nni++;
                const double del = newton_update(0, cmax);
                if (m_it > 0) crate = fmax(cv::CRDOWN * crate, pb_div(del, delp));
                const double dcon = pb_div(del * fmin(1.0, crate), tq4);
                if (dcon <= 1.0) {
                    const int ecmax = errtest_cmax();
                    acnrm = (m_it == 0 && ecmax == cmax) ? del : wnorm1(0, ecmax, S()->acor);
                    jcurNLS = false;
                    jcur = false;
                    publish_state();
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
            if ((ret > 0) && !jcurNLS) {  // retry once with a fresh Jacobian if it was stale
                setup_pending = true;
                jbad = true;
                top = true;
                m_it = 0;
                g.sync();
                PB_FOR_RANGE(e, 0, (cmax + 1) * NX) S()->acor[e] = 0.0;
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
            for (int i = q; i >= 2; --i) S()->tau[i] = S()->tau[i - 1];
            if ((q == 1) && (nst > 1)) S()->tau[2] = S()->tau[1];
            S()->tau[1] = h;
        }
        qwait--;
        const bool save = (qwait == 1) && (q != cv::Q_MAX);
        PB_FOR_ELEMS(e) {
            const double a = S()->acor[e];
            PB_UNROLL
            for (int j = 0; j < cv::L_MAX; ++j)
                if (j <= q) S()->zn[j][e] += l[j] * a;
            if (save) S()->zn[cv::Q_MAX][e] = a;
        }
        g.sync();
        if (save) saved_tq5 = tq5;
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
        etaq = pb_rcp(pb_pow(cv::BIAS2 * dsm, pb_inv_int(L)) + cv::ADDON);
        qprime = q;
        eta = etaq;
        if (qwait == 0) {
            qwait = 2;
            const int ecmax = errtest_cmax();
            etaqm1 = 0.0;
            if (q > 1) {
                const int qq = q;
                double ddn = wnorm1(0, ecmax, S()->zn[qq]);
                ddn = ddn * tq1;
                etaqm1 = pb_rcp(pb_pow(cv::BIAS1 * ddn, pb_inv_int(q)) + cv::ADDON);
            }
            etaqp1 = 0.0;
            if (q != cv::Q_MAX) {
                if (saved_tq5 != 0.0) {
                    const double hr = pb_div(h, S()->tau[2]);
                    double hp = 1.0;
                    for (int i = 0; i < L; ++i) hp *= hr;  // SUNRpowerI
                    const double cquot = pb_div(tq5, saved_tq5) * hp;
                    double dup = wnorm2(0, ecmax, S()->zn[cv::Q_MAX], -cquot, S()->acor, 1.0);
                    dup = dup * tq3;
                    etaqp1 = pb_rcp(pb_pow(cv::BIAS3 * dup, pb_inv_int(L + 1)) + cv::ADDON);
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
                PB_FOR_RANGE(e, 0, (ecmax + 1) * NX) S()->zn[cv::Q_MAX][e] = S()->acor[e];
                g.sync();
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
                sqv[w] = wnorm1(0, 0, S()->zn[jj]);
            }
            int factorial = 1;
            for (int i = 1; i <= q - 1; ++i) factorial *= i;
            const double sq = pb_div(factorial * q * (q + 1) * acnrm, fmax(tq5, cv::TINY));
            const double sqm1 = sqv[0] * (factorial * q);
            const double sqm2 = sqv[1] * factorial;
            g.sync();
            if (g.lane() == 0) {
                for (int k = 1; k <= 3; ++k)
                    for (int i = 5; i >= 2; --i) S()->ssdat[i][k] = S()->ssdat[i - 1][k];
                S()->ssdat[1][1] = sqm2 * sqm2;
                S()->ssdat[1][2] = sqm1 * sqm1;
                S()->ssdat[1][3] = sq * sq;
            }
            g.sync();
        }
        if (qprime >= q) {
            if ((q >= 3) && (nscon >= q + 5)) {
                const int ldflag = sldet_shared<M, Grp>(S(), g, q);
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
        g.sync();
        PB_FOR_ELEMS(e) S()->acor[e] = 0.0;
        g.sync();
        const int r = rhs_all_raw(tn, sensi);
        if (r) return r;
        PB_FOR_ELEMS(e) S()->zn[1][e] = h * S()->b[e];
        g.sync();
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

    /// First part of a step attempt (after a failed attempt: first the restore pass): order /
    /// step-size adjustment, prediction, BDF coefficients.  Returns 0 or a negative failure code.
    PB_HD int attempt_predict() {
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
        fail_ = 0;  // 1/3: corrector failure (state/sens); 2/4: error-test failure (state/sens); 5: f(y_n) not finite
        ret_ = 0;
        return 0;
    }

    /// The corrector of one phase of the attempt (state, then -- staggered -- all sensitivities)
    /// with its error test (cvodes.c:5905-6008).
    PB_HD void attempt_correct(int phase) {
        if (SS && sensi && opt.simultaneous()) {
            // CV_SIMULTANEOUS: one Newton iteration on state and sensitivities together, ONE error test on the
            // joint norm (cvStep cvodes.c:5869-5905 without the staggered block :5938-6010)
            if (phase == PH_SENS) return;
            ret_ = nls_simultaneous(nflag_);
            if (ret_ != 0) {
                fail_ = 1;
            } else {
                dsm_ = acnrm * tq2;
                if (dsm_ > 1.0) fail_ = 2;
            }
            return;
        }
        if (phase == PH_SENS) {
            if (fail_ || !sensi) return;
            ncf_ = nef_ = 0;
            publish_state();
            // f at the converged y (cvodes.c:5947-5955); AMICI's sensitivity RHS does not
            // use it -- kept for the recoverable-failure semantics
            int bad = 0;
            PB_FOR_ROWS0(r, i) S()->yj[i] = S()->ysh[i];
            g.sync();
            if (g.lane() == 0) bad = eval_xdot_shared<M>(S(), tn) ? 0 : 1;
            nfe++;
            if (g.bcast0(bad)) { fail_ = 5; return; }
            refresh_Jcur(tn);
        }
        ret_ = nls(phase, nflag_);
        if (ret_ != 0) {
            fail_ = 1 + 2 * phase;
        } else if (phase == PH_STATE) {
            dsm_ = acnrm * tq2;
            if (dsm_ > 1.0) fail_ = 2;
        } else if (opt.errconS) {
            dsmS_ = acnrmS * tq2;
            if (dsmS_ > 1.0) fail_ = 4;
            else if (dsmS_ > dsm_) dsm_ = dsmS_;
        }
    }

    /// End of the attempt: cvCompleteStep / cvPrepareNextStep / cvBDFStab on success, otherwise
    /// cvHandleNFlag (:7309-7363) and cvDoErrorTest (:7461-7614) folded into ONE failure handler.
    /// Returns 0 when the step was completed, 1 when another attempt is needed, or a negative
    /// failure code.
    PB_HD int attempt_finish() {
        const int fail = fail_;
        const int ret = ret_;
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
                eta = pb_rcp(pb_pow(cv::BIAS2 * (fail == 2 ? dsm_ : dsmS_), pb_inv_int(L)) + cv::ADDON);
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
        PB_FOR_RANGE(e, 0, (ecmax + 1) * NX) {
            const double den = cv::HUB_FACTOR * fabs(S()->zn[0][e]) + pb_rcp(S()->ewt[e]);
            const double v = pb_div(fabs(S()->zn[1][e]), den);
            if (v > hub_inv) hub_inv = v;
        }
        hub_inv = g.max(hub_inv);
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
                g.sync();
                PB_FOR_ELEMS(e) S()->acor[e] = hgs * S()->zn[1][e];
                retval = rhs_all_raw(tn + hgs, withS);
                if (retval == 0) {
                    const double hinv = pb_rcp(hgs);
                    yddnrm = wnorm2(0, ecmax, S()->b, hinv, S()->zn[1], -hinv);
                    hgOK = true;
                    break;
                }
                hg *= 0.2;
            }
            g.sync();
            PB_FOR_ELEMS(e) S()->acor[e] = 0.0;
            g.sync();
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
        // Solver::run -> setStopTime(tout); apply_max_num_steps().  Solver::step (one-step mode,
        // simulation to steady state) sets no stop time.
        if (!one_step()) {
            tstop = tout;
            tstopset = true;
        }
        if (opt.maxsteps <= nst) return SIM_ERROR;
        mxstep_ = opt.maxsteps - nst;
        nstloc_ = 0;
        if (nst == 0) {
            if (!ewt_set()) return SIM_ILL_INPUT;
            h = 1.0;  // zn[1] = f(t0, y0), znS[1] = fS(t0, y0, yS0), unscaled
            const int r = reload_zn1();
            if (r) return r == cv::RHSFUNC_RECVR ? SIM_FIRST_RHSFUNC_ERR : SIM_FIRST_SRHSFUNC_ERR;
            if (tstop_active() && (tstop - tn) * (tout - tn) <= 0.0) return SIM_ILL_INPUT;
            const int hr = hin(tout);
            if (hr != 0) return hr;
            if (tstop_active() && (tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
            hscale = h;
            hprime = h;
            PB_FOR_ELEMS(e) S()->zn[1][e] *= h;
        } else if (tstop_active()) {
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
        PB_FOR_ROWS0(r, i) S()->ysh[i] = S()->zn[0][i];
        g.sync();
        double y[NY], sig[NY];
        M::y(y, t, S()->ysh, S()->p, S()->k);
        M::sigmay(sig, t, S()->p, S()->k);
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
        if (nsall == 0) return true;
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
            M::dydx(dydx, t, S()->ysh, S()->p, S()->k);
            PB_UNROLL
            for (int ix = 0; ix < NX; ++ix) {
                double a = 0.0;
                PB_UNROLL
                for (int iy = 0; iy < NY; ++iy) a += dJydy[iy] * dydx[iy + ix * NY];
                dJydx[ix] = a;
            }
        }
        PB_FOR_COLS(pj, 1, nsall) {
            double dydp[NY], dsdp[NY];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) { dydp[iy] = 0.0; dsdp[iy] = 0.0; }
            const int ip = S()->plist[pj - 1];
            const int c = S()->colof[pj - 1];
            M::dydp(dydp, t, S()->ysh, S()->p, S()->k, ip);
            M::dsigmaydp(dsdp, t, S()->p, S()->k, ip);
            double dJydp = 0.0;
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJydy[iy] * dydp[iy];
            PB_UNROLL
            for (int iy = 0; iy < NY; ++iy) dJydp += dJyds[iy] * dsdp[iy];
            double a = dJydp;
            if (c > 0) {
                PB_UNROLL
                for (int ix = 0; ix < NX; ++ix) a += dJydx[ix] * S()->zn[0][c * NX + ix];
            }
            S()->sllh[pj] -= a;
        }
        return true;
    }

    // ---------------------------------------------------------------- one simulation
    /// Start one simulation (ForwardProblem::workForwardProblem, forwardproblem.cpp:67-203) at
    /// the given tolerances: x0/sx0 (Model::initialize, model.cpp:249-275), CVodeInit
    /// (cvodes.c:677-830).  S()->p / S()->k / S()->plist must be set.
    ///
    /// SS only -- pre-equilibration (amici.cpp:95-108; SteadystateProblem::workSteadyStateProblem with
    /// it = -1, steadystateproblem.cpp:82-106,257-289):
    ///   mode SS_PREEQ   S()->k holds the constants of the PRE-EQUILIBRATION condition; the simulation
    ///                   runs from x0 / sx0 to steady state (a solver of its own in the reference:
    ///                   createSteadystateSimSolver :796-839) and ends with SIM_SUCCESS, the steady state
    ///                   in zn[0];
    ///   mode SS_POSTEQ  (after a successful SS_PREEQ run, S()->k now the simulation condition's
    ///                   constants) start the main simulation from the state and sensitivities in zn[0]
    ///                   (ForwardProblem::ForwardProblem forwardproblem.cpp:59-63); the states of
    ///                   `reinit_mask` are re-initialised from x0 / sx0
    ///                   (Solver::updateAndReinitStatesAndSensitivities -> Model::reinitialize,
    ///                   model.cpp:277-291, fx0_fixedParameters / fsx0_fixedParameters :1849-1923).
    PB_HD void begin(const ConditionIO& io, double rtol, double atol, int mode = 0, unsigned reinit_mask = 0u) {
        io_ = io;
        ss_mode_ = (SS && mode == SS_PREEQ) ? SS_PREEQ : SS_OFF;
        ss_newton_ = 0;
        const bool keep = SS && mode == SS_POSTEQ;  // initial values: the pre-equilibrated state
        if (SS) reinit_mask &= M::X0K_MASK;         // only states whose x0 depends on the constants are re-initialised
        opt.rtol = rtol;
        opt.atol = atol;
        nsall = io.ns;
        g.sync();
        if (g.lane() == 0) {  // which plist entries get a sensitivity column
            int c = 0;
            for (int j = 0; j < nsall; ++j) {
                const int ip = S()->plist[j];
                if ((M::DYN_MASK >> ip) & 1u) {
                    int rank = 0;
                    for (int b_ = 0; b_ < ip; ++b_) rank += (int)((M::DYN_MASK >> b_) & 1u);
                    S()->cdcol[c] = rank;
                    S()->cplist[c] = ip;
                    S()->colof[j] = ++c;
                } else {
                    S()->colof[j] = 0;
                }
            }
            S()->nsd = c;
        }
        g.sync();
        ns = S()->nsd;
        ncol = ns + 1;
        sensi = ns > 0;
        for (int j = g.lane(); j <= nsall; j += G) S()->sllh[j] = 0.0;
        PB_FOR_COLS(c, 0, ns) {
            double v[NX];
            PB_UNROLL
            for (int i = 0; i < NX; ++i) v[i] = 0.0;
            if (c == 0) M::x0(v, opt.t0, S()->p, S()->k);
            else M::sx0(v, opt.t0, S()->p, S()->k, S()->cplist[c - 1]);
            PB_UNROLL
            for (int i = 0; i < NX; ++i) {
                const int e = c * NX + i;
                if (!keep || ((reinit_mask >> i) & 1u)) S()->zn[0][e] = v[i];
                for (int j = 1; j < cv::L_MAX; ++j) S()->zn[j][e] = 0.0;
                S()->acor[e] = 0.0;
                S()->ewt[e] = 1.0;
                S()->b[e] = 0.0;
            }
        }
        if (SS) {
            // (pre-equilibration in sensitivity mode newtonOnly: the simulation's solver is created without
            // sensitivities, createSteadystateSimSolver steadystateproblem.cpp:796-839; sx0 is set above because a
            // Newton attempt that starts in the steady state leaves the sensitivities at their initial values)
            sens_off_ = false;
            if (ss_mode_ == SS_PREEQ) sens_toggle_off();
        }
        PB_UNROLL
        for (int j = 0; j < cv::L_MAX; ++j) l[j] = 0.0;
        tq1 = tq2 = tq3 = tq4 = tq5 = 0.0;
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
            S()->Jsaved[e] = 0.0;
            S()->Jcur[e] = 0.0;
        }
        for (int e = g.lane(); e < NX * GroupShared<M>::NPD; e += G) S()->dfdp[e] = 0.0;
        if (g.lane() == 0) {
            for (int i = 0; i <= cv::L_MAX; ++i) S()->tau[i] = 0.0;
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 4; ++k) S()->ssdat[i][k] = 0.0;
        }
        g.sync();
        llh_ = 0.0;
        it_ = 0;
        sim_status_ = SIM_SUCCESS;
        in_cvode_ = in_step_ = false;
        stage_ = ST_PRE;
        fail_ = ret_ = 0;
        step_begin();
    }

    /// The simulation advances by at most ONE step attempt per round of
    ///     stage_pre();  stage_nls(PH_STATE);  stage_nls(PH_SENS);  stage_post();
    /// Each stage returns to the caller, so that the kernel can re-converge the lane groups of a
    /// warp (and the warps of a CTA) between the stages: every group then walks through the same
    /// code at the same time, whatever happened to it in the previous stage.
    ///
    /// stage_pre: whatever bookkeeping precedes the attempt -- output times that need no
    /// integration, the CVode prologue, the per-step checks of the CVode loop cvodes.c:3366-3490
    /// -- then order/step adjustment, prediction and BDF coefficients.
    PB_HD void stage_pre() {
        if (stage_ != ST_PRE) return;
        if (SS && ss_newton_ == 3) {  // (requested by finish())
            if (!steady_retry()) finish(SIM_ERROR);
            return;
        }
        if (!in_step_) {
            if (!in_cvode_) {
                for (;;) {
                    if (SS && ss_mode_ == SS_PREEQ) break;
                    if (it_ >= io_.nt) { finish(SIM_SUCCESS); return; }
                    tout_ = io_.t[it_];
                    if (!PB_FINITE(tout_)) {
                        // t = inf: post-equilibration (amici.cpp:109-114; ForwardProblem stops at the first infinite
                        // output time, forwardproblem.cpp:149-150).  Without steady-state support an error, never a
                        // silently dropped point.
                        if (!SS) { finish(SIM_ERROR); return; }
                        if (ss_mode_ != SS_POSTEQ) {
                            ss_newton_ = 0;
                            sens_toggle_off();  // (the solver carries on without them, :728-736)
                        }
                        ss_mode_ = SS_POSTEQ;
                        break;
                    }
                    bool integrate = (tout_ > opt.t0) && (tn < tout_);
                    if (integrate) {  // is_next_t_too_close, forwardproblem.cpp:23-35
                        const double tdiff = tout_ - tn;
                        if (fabs(tdiff) < 2.0 * cv::UROUND * fmax(fabs(tn), fabs(tout_))) integrate = false;
                    }
                    if (integrate) break;
                    if (!output_point()) return;
                }
                if (one_step()) {
                    const double rtol_ss = 100.0 * opt.rtol, atol_ss = 100.0 * opt.atol;  // steady-state tolerances
                    const double* xchk = S()->zn[0];
                    if (opt.newton_maxsteps != 0 && ss_newton_ == 0) {
                        // SteadystateProblem::findSteadyState (steadystateproblem.cpp:120-166): Newton's method first
                        // (newton_maxsteps > 0; AMICI's default 0 goes straight to the simulation).  Pre-equilibration:
                        // in place -- a simulation after a failed attempt starts from the last trial point, as the
                        // reference's does (createSteadystateSimSolver sets the new solver up at state_.x).
                        // Post-equilibration: on a copy (acor is free between steps), the solver keeps its state.
                        ss_newton_ = 1;
                        double* x = S()->zn[0];
                        if (ss_mode_ == SS_POSTEQ) {
                            x = S()->acor;
                            g.sync();
                            PB_FOR_RANGE(e, 0, NX) x[e] = S()->zn[0][e];
                        }
                        const int it = ss_newton_shared<M, Grp>(S(), g, tn, rtol_ss, atol_ss, opt.newton_steps(), false, x);
                        if (it >= 0) { steady_found(it > 0, x); return; }
                        if (opt.newton_only()) {  // turnOffSimulation: nothing else is tried (handleSteadyStateFailure)
                            ss_newton_ = 2;
                            finish(SIM_ERROR);
                            return;
                        }
                        xchk = x;  // the simulation's first convergence test reads the last Newton trial point
                    }
                    // one trip of SteadystateProblem::runSteadystateSimulation (steadystateproblem.cpp:711-794):
                    // convergence test at the solver's current state, then ONE internal step towards
                    // max(t, 1) * 10.  Steady-state tolerances: 100 x the integration tolerances.
                    if (ss_converged_shared<M, Grp>(S(), g, tn, rtol_ss, atol_ss, ns, nsall, xchk)) {
                        steady_success();
                        return;
                    }
                    if (nst >= opt.maxsteps) { finish(SIM_ERROR); return; }
                    tout_ = fmax(tn, 1.0) * 10.0;
                }
                const int r = cvode_begin(tout_);
                if (r < 0) { finish(r); return; }
                if (r == cv::TSTOP_RETURN) return;  // already at tout: data point in the next round
                in_cvode_ = true;
            }
            // top of the CVode step loop
            if (nst > 0) {
                if (!ewt_set()) { finish(SIM_ILL_INPUT); return; }
            }
            if (nstloc_ >= mxstep_) { finish(SIM_TOO_MUCH_WORK); return; }
            const double nrm = wnorm1(0, errtest_cmax(), S()->zn[0]);
            if (cv::UROUND * nrm > 1.0) { finish(SIM_TOO_MUCH_ACC); return; }
            step_begin();
            in_step_ = true;
        }
        const int r = attempt_predict();
        if (r < 0) { finish(r); return; }
        stage_ = ST_NLS;
    }

    /// The steady state is reached: every t = inf output time is evaluated there
    /// (ReturnData::processPostEquilibration, rdata.cpp:234-258); a pre-equilibration just ends.
    PB_HD void steady_success() {
        if (sens_back_on()) {
            // the simulation ran without the sensitivities: sx = -J^-1 dxdotdp at the steady state
            // (getSensitivityFlag(newtonSensi), workSteadyStateProblem :91-103)
            if (ss_newton_sensis_shared<M, Grp>(S(), g, tn, ns)) {
                ss_newton_ = 2;
                finish(SIM_ERROR);
                return;
            }
        }
        if (ss_mode_ == SS_POSTEQ) {
            while (it_ < io_.nt) {
                tout_ = io_.t[it_];
                if (!output_point()) return;
            }
        }
        finish(SIM_SUCCESS);
    }

    /// Newton's method found the steady state x: it becomes the state, and -- unless the iteration started in the
    /// steady state (no step taken: getSensitivityFlag, steadystateproblem.cpp:453-517) -- the sensitivities are
    /// those of the steady state, sx = -J^-1 dxdotdp (workSteadyStateProblem :91-103; a singular Jacobian there is
    /// an AmiException -> AMICI_ERROR).
    PB_HD void steady_found(bool newton_sensis, const double* x) {
        g.sync();
        if (x != S()->zn[0]) PB_FOR_RANGE(e, 0, NX) S()->zn[0][e] = x[e];
        g.sync();
        sens_back_on();
        if (sensi && newton_sensis) {
            if (ss_newton_sensis_shared<M, Grp>(S(), g, tn, ns)) {
                ss_newton_ = 2;
                finish(SIM_ERROR);
                return;
            }
        }
        steady_success();
    }

    /// The simulation to steady state failed: Newton's method once more, from the last state the simulation reached
    /// and with at least one step (findSteadyStateByNewtonsMethod(model, true), steadystateproblem.cpp:160-162).
    /// Returns true if that found the steady state.
    PB_HD bool steady_retry() {
        ss_newton_ = 2;
        if (in_step_ && undo_) {  // the failure came in the middle of a step: back to the last completed one (cvRestore)
            tn = saved_t_;
            shift_history(-1.0);
            undo_ = false;
        }
        in_step_ = false;
        g.sync();
        const int it = ss_newton_shared<M, Grp>(S(), g, tn, 100.0 * opt.rtol, 100.0 * opt.atol, opt.newton_maxsteps, true,
                                                S()->zn[0]);
        if (it < 0) return false;
        steady_found(true, S()->zn[0]);
        return true;
    }

    /// The data point of output time it_ (tout_) at the current state; false = the simulation has ended
    /// with an error (sigma not positive).
    PB_HD bool output_point() {
        double* yo = io_.y_out ? io_.y_out + (size_t)it_ * NY : nullptr;
        double* so = io_.sigmay_out ? io_.sigmay_out + (size_t)it_ * NY : nullptr;
#if defined(PB200_HOST_TRACE)
        if (io_.trace) { io_.trace[2 * it_] = nst; io_.trace[2 * it_ + 1] = q; }
#endif
        if (!data_point(tout_, io_.my + (size_t)it_ * NY, io_.sd + (size_t)it_ * NY, llh_, yo, so)) {
            finish(SIM_ERROR);
            return false;
        }
        it_++;
        return true;
    }

    /// stage_nls: the Newton corrector of one phase (state, then all sensitivities) + error test
    PB_HD void stage_nls(int phase) {
        if (stage_ != ST_NLS) return;
        attempt_correct(phase);
    }

    /// stage_post: step completion or failure handling, then the tstop tests after a completed
    /// step (cvodes.c:3504-3605).  Returns true when the simulation has ended; status() / llh()
    /// then hold the result and d llh / d p (w.r.t. the unscaled parameter) is in S()->sllh[1..ns].
    PB_HD bool stage_post() {
        if (stage_ == ST_DONE) return true;
        if (stage_ != ST_NLS) return false;
        stage_ = ST_PRE;
        const int r = attempt_finish();
        if (r == 1) return false;
        in_step_ = false;
        if (r < 0) return finish(r);
        nstloc_++;
        if (one_step()) {  // CV_ONE_STEP: return after every internal step (cvodes.c:3590-3600)
            in_cvode_ = false;
            return false;
        }
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

    /// One round of the four stages (host harness)
    PB_HD bool advance() {
        stage_pre();
        for (int phase = PH_STATE; phase <= PH_SENS; ++phase) stage_nls(phase);
        return stage_post();
    }

    PB_HD bool finish(int status) {
        if (SS && one_step() && status != SIM_SUCCESS && opt.newton_maxsteps > 0 && ss_newton_ < 2) {
            // the simulation to steady state failed: Newton's method once more, at the top of the next stage_pre
            // (ONE copy of that code instead of one per call site of finish())
            ss_newton_ = 3;
            stage_ = ST_PRE;
            return false;
        }
        stage_ = ST_DONE;
        // any failure while simulating to steady state ends as AmiException -> AMICI_ERROR
        // (SteadystateProblem::findSteadyStateBySimulation / handleSteadyStateFailure,
        // steadystateproblem.cpp:198-255,397-425; amici.cpp:186-192)
        if (one_step() && status != SIM_SUCCESS) status = SIM_ERROR;
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
        s[STAT_TRIALS] = trials; s[STAT_QSUM] = qsum;
        for (int j = 14; j < NSTATS; ++j) s[j] = 0;
    }
#undef PB_FOR_COLS
#undef PB_FOR_RANGE
#undef PB_FOR_ROWS
#undef PB_FOR_ROWS0
#undef PB_FOR_ELEMS
};

}  // namespace pb200
#endif
