// bdf_team.cuh -- the large-model integrator: stiff BDF with Newton iteration and staggered
// forward sensitivities, ONE TEAM (thread block) per experimental condition.
//
// Same reference rows as bdf_core.cuh (ForwardProblem +
// CVODES cvStep/cvNls/cvStgrNls/cvDoErrorTest/... + dense linear algebra + likelihood; the
// file:line citations are on the corresponding functions of bdf_core.cuh and are not repeated
// at every function here), for models whose working set does not fit an SM: nx in the hundreds
// to low thousands (BASELINE configs 3-4, "Chen-shape" nx = 500, "Froehlich-shape" nx = 1200).
// CVODES' heuristics and constants are kept; the execution model is different from both CVODES
// and the small-model kernel:
//
//   * the Nordsieck history zn[0..5], accumulated correction, error weights and right-hand
//     side of ALL 1+nplist columns (element (i, c) at [c*nx + i]) live in the team's slice of an
//     HBM workspace; every vector operation is a coalesced team-wide sweep over the flat index;
//   * model functions are evaluated by the whole team from tables (net_model.cuh): one thread
//     per reaction / species / Jacobian entry; J and the sensitivity right-hand sides
//     (dxdotdp + J*s for all columns at once) use the sparse pattern;
//   * the Newton matrix I - gamma*J is dense and INVERTED at every linear-solver setup by a
//     blocked in-place Gauss-Jordan elimination with partial pivoting: the n x NB panel is
//     factored in shared memory, the rest of the matrix gets one rank-NB update per block step
//     (each thread owns a matrix row, the panel row in registers, the pivot rows broadcast from
//     shared memory).  Every Newton iteration -- of the state and of all nplist sensitivity
//     columns together, the staggered corrector's "same matrix, nplist right-hand sides" -- is
//     then ONE GEMM  delta = -(I - gamma J)^-1 * B  (rows owned by threads, streamed from
//     HBM/L2 once per 26 columns, B tile broadcast from shared memory);
//   * the integrator's scalars (step size, order, BDF coefficients, counters) are computed
//     redundantly and identically by every thread; reductions are fixed-order.
//
// The same source compiles for the host with HostTeam (one thread); that instantiation exists
// only in tests/host_harness to diff the algorithm against the CVODES oracle without a GPU.
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
#ifndef PB200_BDF_TEAM_CUH
#define PB200_BDF_TEAM_CUH

#include "bdf_core.cuh"
#include "net_model.cuh"
#include "sparse_lu.h"
#include "team.cuh"

namespace pb200 {

constexpr int TEAM_NB_MAX = 32;     // panel width of the blocked inversion
constexpr int TEAM_CW = 64;         // columns per chunk of the rank-NB update
constexpr int TEAM_RBUF = TEAM_NB_MAX * TEAM_CW;  // doubles in the broadcast tile
constexpr int TEAM_TN = 26;         // right-hand-side columns per pass of the solve GEMM
constexpr int TEAM_MAXNY = 16;
constexpr int TEAM_JW_MAX = 24;     // widest Jacobian row (ELL width) the sensitivity right-hand side holds in registers
/// shared-memory layout of the team kernel (doubles): reduction scratch | pivot row | displaced row |
/// pivots | row-permutation lists | two broadcast tiles (double buffered) | panel (generic path only)
/// (tiles read as tensor-core B fragments are padded: 64 x 36 / 32 x 68 doubles)
constexpr int TEAM_RTILE = 3808, TEAM_RS = 36, TEAM_BS = 68;
constexpr int TEAM_SM_RED = 0, TEAM_SM_ROWK = 64, TEAM_SM_XBUF2 = 96, TEAM_SM_PIV = 128, TEAM_SM_TL = 144,
              TEAM_SM_NT = 208, TEAM_SM_PERM = 224 /* 512 ints */, TEAM_SM_SLOT = 480 /* 512 ints */, TEAM_SM_RBUF = 736, TEAM_SM_RBUF1 = TEAM_SM_RBUF + TEAM_RTILE,
              TEAM_SM_PANEL = TEAM_SM_RBUF1 + TEAM_RTILE;
/// rows path (nx <= 512): ring of matrix elements in flight (cp.async), TEAM_ST stages of TEAM_GC
/// columns x 512 rows; aliases the panel of the generic path (a launch uses one of the two)
constexpr int TEAM_GC = 8, TEAM_ST = 4, TEAM_SM_XSTAGE = TEAM_SM_PANEL,
              TEAM_XCS = 514,  // column stride of a ring slot read as tensor-core C fragments (conflict-free)
              TEAM_SM_ROWS_END = TEAM_SM_XSTAGE + TEAM_ST * TEAM_GC * TEAM_XCS;

#if defined(__CUDA_ARCH__)
/// 8-byte asynchronous global -> shared copy (LDGSTS); nbytes = 0 writes zeros without reading
__device__ __forceinline__ void pb_cp_async8(double* smem_dst, const double* gsrc, int nbytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void pb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
/// 16-byte variant (both addresses 16-byte aligned)
__device__ __forceinline__ void pb_cp_async16(double* smem_dst, const double* gsrc, int nbytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(nbytes) : "memory");
}
template <int N>
__device__ __forceinline__ void pb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
/// L2 prefetch of the line holding *p (no register, no shared memory: extends the reach of the ring)
__device__ __forceinline__ void pb_prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
/// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  Fragments: a = A[lane>>2][lane&3],
/// b = B[lane&3][lane>>2], d0/d1 = D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void pb_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

/// Per-team working set: pointers into the team's slice of the HBM workspace plus the staging
/// buffers (shared memory on the device).
struct TeamWork {
    double* znbase;  // Nordsieck history zn[0..5], each nxc doubles
    size_t nxc;
    double *acor, *ewt, *b, *scr;
    double *dfdp, *A, *Jsaved, *Jcur, *ysh, *yj, *fout, *vb, *v, *p, *k, *sllh;
    int *plist, *perm;
    // sparse path (sparse_lu.h): factor values V of the sparse part, and the row gather of the dense
    // Schur solve (original row index of the k-th pivot row of the block)
    double* luV;
    int* gidx;
    double* stage;      // host harness: staging area of the sparse triangular solves
    int stage_doubles;  // its capacity (device: shared memory from TEAM_SM_RBUF to the end)
    int panel_doubles;  // capacity of the generic dense path's panel (device: from TEAM_SM_PANEL to the end)
    // staging buffers: host harness only -- on the device they are derived from the dynamic
    // shared memory base at every use (BdfTeam::panel() ...), so that the compiler knows the
    // address space (LDS/STS instead of generic LD/ST)
    double* panel;  // nx * NB
    double* rbuf;   // TEAM_RBUF
    double* rowk;   // TEAM_NB_MAX
    int* piv;       // TEAM_NB_MAX
    int NB;
    PB_HD double* zn(int j) const { return znbase + (size_t)j * nxc; }

    /// doubles of HBM workspace one team needs
    static size_t hbm_doubles(const NetTables& T, int nV = 0) {
        const size_t nxc = (size_t)T.nx * (T.np + 1);
        size_t n = (cv::L_MAX + 4) * nxc + (size_t)T.nx * T.np + (size_t)T.nx * T.nx + 2 * (size_t)T.jell_w * T.nx + 3 * (size_t)T.nx +
                   2 * (size_t)T.nr + T.np + T.nk + (T.np + 1);
        n += (size_t)nV + 2;
        n += ((size_t)T.np + 2 * (size_t)T.nx + 1) / 2 + 1;  // plist, perm, gidx (ints)
        return (n + 15) & ~(size_t)15;
    }
    PB_HD void carve(double* base, const NetTables& T, int nV = 0) {
        const size_t nxc = (size_t)T.nx * (T.np + 1);
        double* q = base;
        znbase = q; q += cv::L_MAX * nxc;
        this->nxc = nxc;
        acor = q; q += nxc;
        ewt = q; q += nxc;
        b = q; q += nxc;
        scr = q; q += nxc;
        dfdp = q; q += (size_t)T.nx * T.np;
        A = q; q += (size_t)T.nx * T.nx;
        Jsaved = q; q += (size_t)T.jell_w * T.nx;  // Jacobians in ELL layout
        Jcur = q; q += (size_t)T.jell_w * T.nx;
        ysh = q; q += T.nx;
        yj = q; q += T.nx;
        fout = q; q += T.nx;
        vb = q; q += T.nr;
        v = q; q += T.nr;
        p = q; q += T.np;
        k = q; q += T.nk;
        sllh = q; q += T.np + 1;
        luV = q; q += nV + 2;
        plist = reinterpret_cast<int*>(q);
        perm = plist + T.np;
        gidx = perm + T.nx;
    }
    /// panel width that fits `smem_doubles` of staging memory (after rbuf / rowk / piv)
    PB_HD static int panel_width(int nx, size_t panel_doubles) {
        int nb = TEAM_NB_MAX;
        while (nb > 1 && (size_t)nx * nb > panel_doubles) nb >>= 1;
        return nb;
    }
};

/// SS: with support for steady states -- output times t = inf (post-equilibration) and a pre-equilibration
/// phase (preequilibrate() + simulate(..., keep_state)); the SS = false instantiation is the code without any of it.
template <class Team, bool SS = false>
struct BdfTeam {
    NetTables T;
    SparseLU SP{};   // symbolic factorisation (enabled = 0: dense path only)
    Team tm;
    TeamWork w;
    SolverOptions opt;
    int nx, ns, ncol;
    int dn;          // dimension (= leading dimension) of the dense matrix in w.A: nx, or the Schur block's m
    bool use_sparse = true;  // linear-solver choice of this launch (SP.enabled permitting)
    bool lin_sparse; // the current factorisation is sparse part + dense Schur block
    int nfallback;   // setups that fell back to the dense path (weak diagonal pivot)
    bool sensi;
    enum { PH_STATE = 0, PH_SENS = 1 };

    // integrator scalars (identical in every thread)
    double tn, h, hprime, hscale, eta, etamax, etaq, etaqm1, etaqp1;
    double gamma, gammap, gamrat, rl1, crate, crateS, delp, acnrm, acnrmS, saved_tq5, tstop;
    double l[cv::L_MAX], tq1, tq2, tq3, tq4, tq5;
    double tau[cv::L_MAX + 1], ssdat[6][4];
    int q, qprime, L, qwait, nst, nscon, nstlp, nstlj, convfail;
    bool jcur, jcurNLS, jcurNLSS, tstopset;
    int nfe, nsetups, nje, nni, nniS, nfSe, netf, netfS, ncfn, ncfnS, nor, qsum;
    ConditionIO io_;
    double llh_, tout_, saved_t_, dsm_, dsmS_;
    int it_, sim_status_, nstloc_, mxstep_, ncf_, nef_, ncfS_, nefS_, nflag_, dq_;
    bool do_rescale_, undo_, restart_;
    bool save_, pending_complete_, ahead_;
    bool fuse_history = true;  // fused history pass between the steps of one CVode call (advance_fused)
    bool fuse_newton = true;   // sparse path: right-hand sides + forward substitution, backward substitution + update + norm
    bool b_forwarded_ = false; // w.b already holds the forward-substituted right-hand sides
    bool fwd_ok_ = false;      // the factors in w.luV are the ones the coming solve uses (no setup pending)
    bool acor_zero_ = false;   // the sensitivity columns of acor are zero but the zeros have not been written
    bool lazy_acor = true;     // (experiments: false = zero acor eagerly at the start of every sensitivity iteration)
    int fail_, ret_;
    // phase cycle counters (device: clock64 of thread 0's SM; exported in stats[14..15] packed as
    // two 32-bit fields of kilo-cycles each: [14] = inversion panel | inversion update, [15] = solve GEMM | model+rhs)
    long long cyc_panel, cyc_update, cyc_solve, cyc_rhs, cyc_spfactor, cyc_sptri, cyc_hist, cyc_state, cyc_sens, cyc_total;
    PB_HD static long long now() {
#if defined(__CUDA_ARCH__)
        return clock64();
#else
        return 0;
#endif
    }

    PB_HD explicit BdfTeam(const Team& t) : tm(t) {}

    // staging buffers: shared memory on the device (layout: TEAM_SM_* below), plain arrays on the host
#if defined(__CUDA_ARCH__)
    __device__ __forceinline__ double* smem_base() const {
        extern __shared__ __align__(16) double pb_team_smem[];
        return pb_team_smem;
    }
    __device__ __forceinline__ double* rowk() const { return smem_base() + TEAM_SM_ROWK; }
    __device__ __forceinline__ int* piv() const { return reinterpret_cast<int*>(smem_base() + TEAM_SM_PIV); }
    __device__ __forceinline__ double* rbuf() const { return smem_base() + TEAM_SM_RBUF; }
    __device__ __forceinline__ double* panel() const { return smem_base() + TEAM_SM_PANEL; }
#else
    double* rowk() const { return w.rowk; }
    int* piv() const { return w.piv; }
    double* rbuf() const { return w.rbuf; }
    double* panel() const { return w.panel; }
#endif

// Team-wide strided sweeps.  Every sweep streams from HBM/L2 with ~50 elements per thread, so what
// matters is the number of independent loads in flight: loops whose stores go to another array
// than their loads use __restrict__ pointers and are unrolled by 4 (PT_FOR); in-place updates load
// a batch of PT_U elements into registers before the first store (PT_FOR_BATCH).
#if defined(__CUDA_ARCH__)
#define PT_UNROLL4 _Pragma("unroll 8")
#else
#define PT_UNROLL4
#endif
    /// staging area for the corrected sensitivity vectors of a batch of columns: 16 columns of 512
    /// doubles in shared memory on the rows path (the ring, idle during right-hand sides), otherwise
    /// all columns at once in the scratch array in HBM
    PB_HD double* sens_stage(int& batch_cols, int& stride) const {
#if defined(__CUDA_ARCH__)
        if (nx <= 512) {
            batch_cols = (TEAM_ST * TEAM_GC * 512) / 512 / 2;  // 16
            stride = 512;
            return smem_base() + TEAM_SM_XSTAGE;
        }
#endif
        batch_cols = ns > 0 ? ns : 1;
        stride = nx;
        return w.scr;
    }

// out of line on the device: own register allocation for the sweeps that hold a batch in registers
#if defined(__CUDACC__)
#define PT_NOINLINE __host__ __device__ __noinline__
#else
#define PT_NOINLINE
#endif
#define PT_FOR(e, e0, e1) PT_UNROLL4 for (int e = (e0) + tm.tid(); e < (e1); e += tm.size())
#define PT_FOR_BATCH(eb, e0, e1, U) for (int eb = (e0) + tm.tid(); eb < (e1); eb += (U) * tm.size())
#define PT_FOR_ELEMS(e) PT_FOR(e, 0, nx * ncol)

    PB_HD int phase_cmin(int phase) const { return phase == PH_STATE ? 0 : 1; }
    PB_HD int phase_cmax(int phase) const { return phase == PH_STATE ? 0 : ns; }
    PB_HD int errtest_cmax() const { return (sensi && opt.errconS) ? ns : 0; }

    // ---------------------------------------------------------------- norms
    /// WRMS norm of ca*A + cb*B (B may be null) with weights ewt over the columns [cmin, cmax],
    /// maximum over the columns (N_VWrmsNorm / cvSensNorm / cvSensUpdateNorm)
    PB_HD double wnorm(int cmin, int cmax, const double* A, double ca, const double* B, double cb) const {
        tm.sync();
        double res;
        if (cmin == cmax) {
            double s = 0.0;
            PT_FOR(e, cmin * nx, (cmin + 1) * nx) {
                const double v = (B ? ca * A[e] + cb * B[e] : ca * A[e]) * w.ewt[e];
                s += v * v;
            }
            res = tm.sum(s);
        } else {
            double loc = 0.0;
            for (int c = cmin + tm.warp(); c <= cmax; c += tm.nwarps()) {
                double s = 0.0;
                PT_UNROLL4
                for (int i = tm.lane(); i < nx; i += Team::WARP) {
                    const int e = c * nx + i;
                    const double v = (B ? ca * A[e] + cb * B[e] : ca * A[e]) * w.ewt[e];
                    s += v * v;
                }
                s = tm.warp_sum(s);
                if (s > loc) loc = s;
            }
            res = tm.max(loc);
        }
        return sqrt(res / (double)nx);
    }
    PB_HD double wnorm1(int cmin, int cmax, const double* A) const { return wnorm(cmin, cmax, A, 1.0, nullptr, 0.0); }

    /// cvEwtSetSS / cvSensEwtSetSS
    PB_HD bool ewt_set() {
        bool ok = true;
        tm.sync();
        const double* __restrict__ z0 = w.zn(0);
        double* __restrict__ ew = w.ewt;
        PT_FOR_ELEMS(e) {
            const double wv = opt.rtol * fabs(z0[e]) + opt.atol;
            ok = ok && (wv > 0.0);
            ew[e] = 1.0 / wv;
        }
        return !tm.any(!ok);
    }

    // ---------------------------------------------------------------- model
    /// J(t, ysh) and d xdot / d p at the converged state (Model_ODE::fsxdot at ip == 0)
    PB_HD void refresh_Jcur() {
        tm.sync();
        net::rates(T, tm, w.ysh, w.p, w.vb, (double*)nullptr);
        net::jac_values(T, tm, w.ysh, w.p, w.Jcur, true);
        tm.sync();
        net::dxdotdp_dense(T, tm, w.vb, w.dfdp);
        tm.sync();
    }

    /// xdot(w.yj) -> w.fout; team-uniform result
    PB_HD bool eval_xdot() {
        tm.sync();
        net::rates(T, tm, w.yj, w.p, w.vb, w.v);
        tm.sync();
        const bool ok = net::xdot_from_rates(T, tm, w.v, w.fout);
        nfe++;
        return !tm.any(!ok);
    }

    /// Sensitivity right-hand sides of all columns (out of line: its own register allocation).
    /// Returns true if a value is not finite (thread-local; the caller reduces).
#if defined(__CUDACC__)
    __host__ __device__ __noinline__
#endif
    bool rhs_sens(bool raw) {
        bool bad = false;
        // Columns in batches: the corrected sensitivity vectors of a batch go to the staging area
        // (shared memory on the rows path), then every thread produces elements of
        // dxdotdp + J*s from the ELL Jacobian -- coalesced loads of the matrix entries shared by
        // all columns, gathers from the staged vectors, four elements per thread in lock-step.
        const double* __restrict__ z0 = w.zn(0);
        const double* __restrict__ z1 = w.zn(1);
        const double* __restrict__ ac = w.acor;
        const double* __restrict__ jc = w.Jcur;
        const double* __restrict__ dfdp = w.dfdp;
        const int* __restrict__ jcol = T.jell_col;
        double* __restrict__ bb = w.b;
        const int jw = T.jell_w;
        int CB, sstride;
        double* const stage = sens_stage(CB, sstride);
        for (int cb = 1; cb <= ns; cb += CB) {
            const int ncb = (ns + 1 - cb < CB) ? ns + 1 - cb : CB;
            const int nel = ncb * nx;
            tm.sync();
            PT_FOR(x, 0, nel) {
                const int cl = x / nx, i = x - cl * nx;
                const size_t e = (size_t)(cb + cl) * nx + i;
                stage[(size_t)cl * sstride + i] = z0[e] + ac[e];
            }
            tm.sync();
            // one matrix row per thread: its (padded) ELL entries in registers, reused for every
            // column of the batch; padding entries carry the value 0 and gather element 0
            for (int i0 = 0; i0 < nx; i0 += tm.size()) {  // (whole warps enter: warp_max_int below)
                const int i = i0 + tm.tid();
                const bool rowok = i < nx;
                double jv[TEAM_JW_MAX];
                int jn[TEAM_JW_MAX];
                int len = 0;
                PB_UNROLL
                for (int t = 0; t < TEAM_JW_MAX; ++t) {
                    const int j = (rowok && t < jw) ? jcol[t * nx + i] : -1;
                    jn[t] = j >= 0 ? j : 0;
                    jv[t] = j >= 0 ? jc[t * nx + i] : 0.0;
                    if (j >= 0) len = t + 1;
                }
                // rows are numbered by decreasing length (netgen): the warp's longest row bounds the loop
                const int wlen = tm.warp_max_int(len);
                if (!rowok) continue;
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
                for (int cl = 0; cl < ncb; ++cl) {
                    const size_t e = (size_t)(cb + cl) * nx + i;
                    const double* sv = stage + (size_t)cl * sstride;
                    double F = dfdp[i + (size_t)w.plist[cb + cl - 1] * nx];
                    PB_UNROLL
                    for (int t = 0; t < TEAM_JW_MAX; ++t)
                        if (t < wlen) F += jv[t] * sv[jn[t]];
                    if (!PB_FINITE(F)) bad = true;
                    bb[e] = raw ? F : (rl1 * z1[e] + ac[e]) - gamma * F;
                }
            }
        }
        return bad;
    }

    PB_HD void materialize_acor_zero() {
        tm.sync();
        PT_FOR(e, nx, (ns + 1) * nx) w.acor[e] = 0.0;
        tm.sync();
        acor_zero_ = false;
    }

    /// rhs_sens(false) fused with the forward substitution of the sparse solver: the right-hand sides of a
    /// batch of columns are formed in the staging area (permuted rows, Z[row][column]), substituted there,
    /// and reach w.b already as  y = L11^-1 b1 | b2' = b2 - L21 y  -- one write of w.b instead of a write, a
    /// read and a second write.  Returns true if a value is not finite (thread-local; the caller reduces).
#if defined(__CUDACC__)
    __host__ __device__ __noinline__
#endif
    bool rhs_sens_fwd() {
        bool bad = false;
        const double* __restrict__ z0 = w.zn(0);
        const double* __restrict__ z1 = w.zn(1);
        const double* __restrict__ ac = w.acor;
        const double* __restrict__ jc = w.Jcur;
        const double* __restrict__ dfdp = w.dfdp;
        const int* __restrict__ jcol = T.jell_col;
        const unsigned short* __restrict__ iord = SP.iord16;
        const int* __restrict__ plist = w.plist;
        double* __restrict__ bb = w.b;
        const int jw = T.jell_w;
        const int nx = this->nx;
        const double rl1 = this->rl1, gamma = this->gamma;
        const bool azero = acor_zero_;
        PB_ASSUME_GLOBAL(z0); PB_ASSUME_GLOBAL(z1); PB_ASSUME_GLOBAL(ac); PB_ASSUME_GLOBAL(jc); PB_ASSUME_GLOBAL(dfdp);
        PB_ASSUME_GLOBAL(jcol); PB_ASSUME_GLOBAL(plist); PB_ASSUME_GLOBAL(bb);
        // staging area: corrected sensitivity vectors of the batch S[column][row] | right-hand sides Z[row][column]
        // | factor values
        tm.sync();
        int zcap;
        const double* const V = sp_stage_values(zcap);
        int CB = zcap / (2 * nx);
        if (CB < 1) CB = 1;  // (the caller checked that two columns' worth of rows fit)
        if (CB > ns) CB = ns;
        double* const S = sp_stage();
        double* const Z = S + (size_t)CB * nx;
        for (int cb = 1; cb <= ns; cb += CB) {
            const int ncb = (ns + 1 - cb < CB) ? ns + 1 - cb : CB;
            const int nel = ncb * nx;
            tm.sync();
            // flat, coalesced: S = zn[0] + acor (what J multiplies), Z = rl1 zn[1] + acor (the rest of the residual)
            if (azero) {  // (acor = 0, not materialised: first iteration)
                PT_FOR(x, 0, nel) {
                    const int cl = x / nx, i = x - cl * nx;
                    const size_t e = (size_t)(cb + cl) * nx + i;
                    S[(size_t)cl * nx + i] = z0[e] + 0.0;
                    Z[(size_t)iord[i] * CB + cl] = rl1 * z1[e] + 0.0;
                }
            } else {
                PT_FOR(x, 0, nel) {
                    const int cl = x / nx, i = x - cl * nx;
                    const size_t e = (size_t)(cb + cl) * nx + i;
                    const double a = ac[e];
                    S[(size_t)cl * nx + i] = z0[e] + a;
                    Z[(size_t)iord[i] * CB + cl] = rl1 * z1[e] + a;
                }
            }
            tm.sync();
            // one matrix row per thread (ELL entries in registers); the only global loads of the loop over the
            // columns -- the entries of dxdotdp -- are issued four columns ahead of their use
            for (int i0 = 0; i0 < nx; i0 += tm.size()) {  // (whole warps enter: warp_max_int below)
                const int i = i0 + tm.tid();
                const bool rowok = i < nx;
                double jv[TEAM_JW_MAX];
                int jn[TEAM_JW_MAX];
                int len = 0;
                PB_UNROLL
                for (int t = 0; t < TEAM_JW_MAX; ++t) {
                    const int j = (rowok && t < jw) ? jcol[t * nx + i] : -1;
                    jn[t] = j >= 0 ? j : 0;
                    jv[t] = j >= 0 ? jc[t * nx + i] : 0.0;
                    if (j >= 0) len = t + 1;
                }
                const int wlen = tm.warp_max_int(len);
                if (!rowok) continue;
                const size_t zrow = (size_t)iord[i] * CB;
                constexpr int UC = 4;
                for (int cl0 = 0; cl0 < ncb; cl0 += UC) {
                    double d[UC];
                    PB_UNROLL
                    for (int u = 0; u < UC; ++u)
                        d[u] = (cl0 + u < ncb) ? dfdp[i + (size_t)plist[cb + cl0 + u - 1] * nx] : 0.0;
                    PB_UNROLL
                    for (int u = 0; u < UC; ++u) {
                        const int cl = cl0 + u;
                        if (cl < ncb) {
                            const double* sv = S + (size_t)cl * nx;
                            double F = d[u];
                            PB_UNROLL
                            for (int t = 0; t < TEAM_JW_MAX; ++t)
                                if (t < wlen) F += jv[t] * sv[jn[t]];
                            if (!PB_FINITE(F)) bad = true;
                            Z[zrow + cl] = Z[zrow + cl] - gamma * F;
                        }
                    }
                }
            }
            tm.sync();
            sparse_forward_staged(Z, CB, ncb, V);
            PT_FOR(x, 0, nel) {
                const int cl = x / nx, i = x - cl * nx;
                bb[(size_t)(cb + cl) * nx + i] = Z[(size_t)iord[i] * CB + cl];
            }
        }
        return bad;
    }

    /// Right-hand sides of the columns of `phase` at v = zn[0] + acor:
    /// b = raw ? F : rl1*zn[1] + acor - gamma*F  (cvNlsResidual / cvNlsResidualSensStg)
    PB_HD int rhs(int phase, bool raw) {
        bool bad = false;
        const long long t_r0 = now();
        tm.sync();
        if (phase == PH_STATE) {
            PT_FOR(i, 0, nx) w.yj[i] = w.zn(0)[i] + w.acor[i];
            bad = !eval_xdot();
            PT_FOR(i, 0, nx) {
                const double F = w.fout[i];
                w.b[i] = raw ? F : (rl1 * w.zn(1)[i] + w.acor[i]) - gamma * F;
            }
        } else {
            // (sparse solver, no setup pending: right-hand sides and forward substitution in one staging pass)
            const bool fwd = fuse_newton && lin_sparse && !raw && fwd_ok_ &&
                             w.stage_doubles - (2 * SP.nV > w.stage_doubles ? 0 : SP.nV) >= 2 * nx;
            if (acor_zero_ && !fwd) materialize_acor_zero();
            bad = fwd ? rhs_sens_fwd() : rhs_sens(raw);
            nfSe += ns;
            bad = tm.any(bad);
            b_forwarded_ = fwd && !bad;
        }
        tm.sync();
        cyc_rhs += now() - t_r0;
        if (bad) return phase == PH_STATE ? cv::RHSFUNC_RECVR : cv::SRHSFUNC_RECVR;
        return 0;
    }

    PB_HD void publish_state() {
        tm.sync();
        PT_FOR(i, 0, nx) w.ysh[i] = w.zn(0)[i] + w.acor[i];
        tm.sync();
    }

    PB_HD int rhs_all_raw(bool with_sens) {
        int r = rhs(PH_STATE, true);
        if (r || !with_sens) return r;
        publish_state();
        refresh_Jcur();
        return rhs(PH_SENS, true);
    }

    // ---------------------------------------------------------------- linear algebra
    /// A = (P (I - gamma*Jsaved))^-1 by blocked in-place Gauss-Jordan elimination with partial
    /// pivoting (pivot rule of SUNDlsMat_denseGETRF); w.perm records the row order P.
    /// Returns 0, or 1 on a zero / non-finite pivot.
    PB_HD int invert_newton_matrix() {
        double* const A = w.A;
        tm.sync();
        for (size_t e = tm.tid(); e < (size_t)nx * nx; e += tm.size()) A[e] = 0.0;
        tm.sync();
        PT_FOR(i, 0, nx) {
            for (int t = 0; t < T.jell_w; ++t) {
                const int j = T.jell_col[t * nx + i];
                if (j >= 0) A[i + (size_t)j * nx] = (i == j ? 1.0 : 0.0) - gamma * w.Jsaved[t * nx + i];
            }
        }
        tm.sync();
        dn = nx;
        lin_sparse = false;
        const int r = invert_dense();
        if (r) return r;
        PT_FOR(i, 0, nx) w.gidx[i] = w.perm[i];
        tm.sync();
        return 0;
    }

    /// In-place inversion (with row pivoting, order in w.perm) of the dn x dn column-major matrix in w.A
    PB_HD int invert_dense() {
        double* const A = w.A;
        const int nx = dn;  // (shadows the member: the dense dimension)
        // panel width: as wide as the staging memory allows (host harness: w.NB > 0 asks for a width)
        int NB = TeamWork::panel_width(nx, (size_t)w.panel_doubles);
        if (w.NB > 0 && w.NB < NB) NB = w.NB;
        PT_FOR(i, 0, nx) w.perm[i] = i;
        tm.sync();
#if defined(__CUDA_ARCH__)
        if (nx <= tm.size()) return invert_rows();
#endif
        for (int k0 = 0; k0 < nx; k0 += NB) {
            const int nb = (nx - k0 < NB) ? nx - k0 : NB;
            // 1. panel -> staging memory
            const long long t_p0 = now();
            PT_FOR(e, 0, nx * nb) panel()[e] = A[(size_t)k0 * nx + e];
            tm.sync();
            // 2. Gauss-Jordan on the panel columns
            for (int kk = 0; kk < nb; ++kk) {
                const int k = k0 + kk;
                double* const colk = panel() + (size_t)kk * nx;
                double best = -1.0;
                int bi = 0x7fffffff;
                for (int i = k + tm.tid(); i < nx; i += tm.size()) {
                    const double m = fabs(colk[i]);
                    if (m > best) { best = m; bi = i; }
                }
                best = tm.argmax(best, bi);
                if (!(best > 0.0) || !PB_FINITE(best)) return 1;
                const int lp = bi;
                if (lp != k) {
                    PT_FOR(c, 0, nb) {
                        double* col = panel() + (size_t)c * nx;
                        const double t_ = col[lp];
                        col[lp] = col[k];
                        col[k] = t_;
                    }
                }
                if (tm.tid() == 0) {
                    piv()[kk] = lp;
                    const int t_ = w.perm[k];
                    w.perm[k] = w.perm[lp];
                    w.perm[lp] = t_;
                }
                tm.sync();
                const double pinv = 1.0 / colk[k];
                PT_FOR(c, 0, nb) rowk()[c] = panel()[k + (size_t)c * nx] * pinv;
                tm.sync();
                // other panel columns: rank-1 update; column kk itself afterwards
                PT_FOR(e, 0, nx * nb) {
                    const int c = e / nx, i = e - c * nx;
                    if (c == kk) continue;
                    if (i == k) panel()[e] = rowk()[c];
                    else panel()[e] -= colk[i] * rowk()[c];
                }
                tm.sync();
                PT_FOR(i, 0, nx) colk[i] = (i == k) ? pinv : -colk[i] * pinv;
                tm.sync();
            }
            const long long t_p1 = now();
            cyc_panel += t_p1 - t_p0;
            // 3. the row swaps of this block step, applied in order to the other columns
            PT_FOR(j, 0, nx) {
                if (j >= k0 && j < k0 + nb) continue;
                double* col = A + (size_t)j * nx;
                for (int kk = 0; kk < nb; ++kk) {
                    const int lp = piv()[kk];
                    if (lp != k0 + kk) {
                        const double t_ = col[lp];
                        col[lp] = col[k0 + kk];
                        col[k0 + kk] = t_;
                    }
                }
            }
            tm.sync();
            // 4. rank-nb update of the other columns: A[:, j] = (rows of the block zeroed) + W * R[:, j]
            //    with W = panel (this thread's row in registers) and R = the old pivot rows of A
            for (int j0 = 0; j0 < nx; j0 += TEAM_CW) {
                const int cw = (nx - j0 < TEAM_CW) ? nx - j0 : TEAM_CW;
                PT_FOR(e, 0, TEAM_NB_MAX * ((cw + 1) & ~1)) {
                    const int jj = e / TEAM_NB_MAX, kk = e - jj * TEAM_NB_MAX;
                    rbuf()[e] = (kk < nb && jj < cw) ? A[(k0 + kk) + (size_t)(j0 + jj) * nx] : 0.0;
                }
                tm.sync();
                for (int i = tm.tid(); i < nx; i += tm.size()) {
                    double wr[TEAM_NB_MAX];
                    PB_UNROLL
                    for (int kk = 0; kk < TEAM_NB_MAX; ++kk) wr[kk] = (kk < nb) ? panel()[i + (size_t)kk * nx] : 0.0;
                    const bool in_block = (i >= k0 && i < k0 + nb);
                    // two columns per trip: two independent accumulation chains
                    for (int jj = 0; jj < cw; jj += 2) {
                        const int j = j0 + jj;
                        const bool do0 = !(j >= k0 && j < k0 + nb);
                        const bool do1 = (jj + 1 < cw) && !(j + 1 >= k0 && j + 1 < k0 + nb);
                        double x0 = (in_block || !do0) ? 0.0 : A[i + (size_t)j * nx];
                        double x1 = (in_block || !do1) ? 0.0 : A[i + (size_t)(j + 1) * nx];
#if defined(__CUDA_ARCH__)
                        const double2* r0 = reinterpret_cast<const double2*>(rbuf() + jj * TEAM_NB_MAX);
                        const double2* r1 = r0 + TEAM_NB_MAX / 2;  // next column (inside rbuf: cw is padded to even)
                        PB_UNROLL
                        for (int k2 = 0; k2 < TEAM_NB_MAX / 2; ++k2) {
                            const double2 a = r0[k2], bq = r1[k2];
                            x0 += wr[2 * k2] * a.x;
                            x1 += wr[2 * k2] * bq.x;
                            x0 += wr[2 * k2 + 1] * a.y;
                            x1 += wr[2 * k2 + 1] * bq.y;
                        }
#else
                        const double* r0 = rbuf() + jj * TEAM_NB_MAX;
                        const double* r1 = r0 + TEAM_NB_MAX;
                        for (int kk = 0; kk < TEAM_NB_MAX; ++kk) {
                            x0 += wr[kk] * r0[kk];
                            if (do1) x1 += wr[kk] * r1[kk];
                        }
#endif
                        if (do0) A[i + (size_t)j * nx] = x0;
                        if (do1) A[i + (size_t)(j + 1) * nx] = x1;
                    }
                }
                tm.sync();
            }
            // 5. the transformed panel replaces the block columns
            PT_FOR(e, 0, nx * nb) A[(size_t)k0 * nx + e] = panel()[e];
            tm.sync();
            cyc_update += now() - t_p1;
        }
        return 0;
    }

#if defined(__CUDA_ARCH__)
    /// The same blocked Gauss-Jordan inversion for nx <= blockDim.x: thread i owns matrix row i.
    /// Its row of the panel lives in REGISTERS from the panel factorisation through the rank-NB
    /// update (no panel in shared memory): per pivot one arg-max reduction and one exchange of the
    /// pivot row through shared memory; the row swaps of a block step are composed into one
    /// permutation and applied to the other columns with independent loads; the rank-NB update
    /// streams the matrix once per block step with the next 8 columns prefetched and the pivot-row
    /// tile double-buffered in shared memory.
    __device__ __noinline__ int invert_rows() {
        constexpr int NB = TEAM_NB_MAX;
        // members used in the loops are copied to locals: the object is addressable (out-of-line member
        // function), so a store through a generic pointer would force a reload of every member after it
        const int nx = this->dn;  // dimension of the dense matrix (the whole Newton matrix or the Schur block)
        double* const A = w.A;
        PB_ASSUME_GLOBAL(A);
        const int i = threadIdx.x;
        const bool rowok = i < nx;
        double* const sm = smem_base();
        double* const xbuf = sm + TEAM_SM_ROWK;
        double* const xbuf2 = sm + TEAM_SM_XBUF2;
        int* const pv = reinterpret_cast<int*>(sm + TEAM_SM_PIV);
        int* const tlist = reinterpret_cast<int*>(sm + TEAM_SM_TL);
        int* const tsrc = tlist + 2 * NB;
        int* const ntp = reinterpret_cast<int*>(sm + TEAM_SM_NT);
        int* const perm_s = reinterpret_cast<int*>(sm + TEAM_SM_PERM);  // row order, written back at the end
        int* const slot_of = reinterpret_cast<int*>(sm + TEAM_SM_SLOT);
        perm_s[i] = i;
        slot_of[i] = -1;
        __syncthreads();
        for (int k0 = 0; k0 < nx; k0 += NB) {
            const int nb = (nx - k0 < NB) ? nx - k0 : NB;
            const long long t_p0 = now();
            double prow[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) prow[c] = (rowok && c < nb) ? A[i + (size_t)(k0 + c) * nx] : 0.0;
            // ---- panel: Gauss-Jordan on the block columns, rows in registers.  The loop over the
            //      pivots is ROLLED (one copy of the body in the instruction cache: unrolled, its 32
            //      copies ran once per block step each and were fetched from L2 every time): the
            //      register array is rotated by one column per pivot, so the pivot column is always
            //      element 0 and every index is a compile-time constant.
#pragma unroll 1
            for (int kk = 0; kk < nb; ++kk) {
                const int k = k0 + kk;
                const double mag = (rowok && i >= k) ? fabs(prow[0]) : -1.0;
                int bi = i;
                const double best = tm.argmax_inline(mag, bi);
                if (!(best > 0.0) || !PB_FINITE(best)) return 1;
                const int lp = bi;
                if (i == lp) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) xbuf[c] = prow[c];
                }
                if (lp != k && i == k) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) xbuf2[c] = prow[c];
                }
                if (i == 0) {
                    pv[kk] = lp;
                    const int t_ = perm_s[k];
                    perm_s[k] = perm_s[lp];
                    perm_s[lp] = t_;
                }
                __syncthreads();
                if (lp != k && i == lp) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) prow[c] = xbuf2[c];
                }
                const double pinv = 1.0 / xbuf[0];
                double p0;
                if (i == k) {
#pragma unroll
                    for (int c = 1; c < NB; ++c) prow[c] = xbuf[c] * pinv;
                    p0 = pinv;
                } else {
                    const double m = prow[0] * pinv;
#pragma unroll
                    for (int c = 1; c < NB; ++c) prow[c] -= m * xbuf[c];
                    p0 = -m;
                }
#pragma unroll
                for (int c = 0; c < NB - 1; ++c) prow[c] = prow[c + 1];
                prow[NB - 1] = p0;
            }
#pragma unroll 1
            for (int r = nb; r < NB; ++r) {  // short last block: finish the turn of the register array
                const double p0 = prow[0];
#pragma unroll
                for (int c = 0; c < NB - 1; ++c) prow[c] = prow[c + 1];
                prow[NB - 1] = p0;
            }
            // ---- the row swaps of this block step as one permutation: position tlist[t] receives
            //      the row that was at tsrc[t] when the block step began
            if (i == 0) {
                int n = 0;
                for (int kk = 0; kk < nb; ++kk) {
                    const int k = k0 + kk, lp = pv[kk];
                    if (lp == k) continue;
                    int a = slot_of[k], b = slot_of[lp];  // list position of a row, -1 = not yet touched
                    if (a < 0) { a = n; tlist[n] = k; tsrc[n] = k; slot_of[k] = n; ++n; }
                    if (b < 0) { b = n; tlist[n] = lp; tsrc[n] = lp; slot_of[lp] = n; ++n; }
                    const int t_ = tsrc[a];
                    tsrc[a] = tsrc[b];
                    tsrc[b] = t_;
                }
                for (int t = 0; t < n; ++t) slot_of[tlist[t]] = -1;
                *ntp = n;
            }
            __syncthreads();
            const int nt = *ntp;
            if (nt > 0) {
                for (int j = i; j < nx; j += blockDim.x) {
                    if (j >= k0 && j < k0 + nb) continue;
                    double* col = A + (size_t)j * nx;
                    double v[2 * NB];
                    // all sources are read before the first write (the permutation may chain); eight
                    // independent loads are issued before their values are parked in local memory
                    for (int t0 = 0; t0 < nt; t0 += 8) {
                        double r8[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) r8[u] = (t0 + u < nt) ? col[tsrc[t0 + u]] : 0.0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[t0 + u] = r8[u];
                    }
                    for (int t = 0; t < nt; ++t)
                        if (tsrc[t] != tlist[t]) col[tlist[t]] = v[t];
                }
            }
            __syncthreads();
            const long long t_p1 = now();
            cyc_panel += t_p1 - t_p0;
            // ---- the transformed panel replaces the block columns (the update below does not touch
            //      them), and goes to the tensor-core fragment layout through shared memory: warp w
            //      owns rows 32w..32w+31 as four 8-row tiles, W fragments for all 8 k-steps in registers
            double* const xs = sm + TEAM_SM_XSTAGE;
            if (rowok) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < nb) A[i + (size_t)(k0 + c) * nx] = prow[c];
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) xs[c * 512 + i] = prow[c];
            // nx <= 256: only half of the warps own rows -- TWO warps then share a 32-row group and take
            // every other column group of the update (rows 256.. of the ring slots are free for the second)
            const bool split = nx <= 256;
            if (split) __syncthreads();
            else __syncwarp();
            const int lane = i & 31, lq = lane & 3, lr = lane >> 2;
            const int wrow0 = split ? ((i >> 5) & 7) * 32 : (i & ~31);  // first row of this warp's row group
            const int gstep = split ? 2 : 1, goff = split ? (i >> 8) : 0;
            const int rslot = split ? 256 * (i >> 8) : 0;                // row offset of this warp inside a ring slot
            const int rbase = wrow0 + lr;  // row of m-tile 0 this lane's fragments belong to
            double wfr[4][NB / 4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                for (int ks = 0; ks < NB / 4; ++ks) wfr[mt][ks] = xs[(4 * ks + lq) * 512 + rbase + 8 * mt];
            }
            if (split) __syncthreads();  // the ring below reuses this staging area
            else __syncwarp();
            // ---- rank-nb update of the other columns on the FP64 tensor cores:
            //      C(32 rows of the warp x 8 columns) = C' + W(32 x nb) * R(nb x 8) per column group
            const int nchunks = (nx + TEAM_CW - 1) / TEAM_CW;
            auto r_entry = [&](int ch, int e) -> double {  // entry e of the pivot-row tile of chunk ch
                const int jj = e / NB, kk = e - jj * NB;
                const int j = ch * TEAM_CW + jj;
                return (kk < nb && j < nx) ? A[(k0 + kk) + (size_t)j * nx] : 0.0;
            };
            auto r_pos = [&](int e) { return (e / NB) * TEAM_RS + (e % NB); };  // padded: conflict-free fragment reads
            {
                double* const b0 = sm + TEAM_SM_RBUF;
#pragma unroll
                for (int u = 0; u < TEAM_RBUF / 512; ++u) {
                    const int e = i + u * 512;
                    b0[r_pos(e)] = r_entry(0, e);
                }
            }
            __syncthreads();
            // this lane's fragment elements of the next TEAM_ST-1 column groups are always in flight
            // (cp.async into its own slots of the ring: no barrier needed to read them back)
            constexpr int GC = TEAM_GC, ST = TEAM_ST;
            static_assert(GC == 8, "one 8-column tensor-core tile per group");
            const int ngroups = (nx + GC - 1) / GC;
            auto col_ok = [&](int j) { return j < nx && !(j >= k0 && j < k0 + nb); };
            auto row_in_block = [&](int r) { return r >= k0 && r < k0 + nb; };
            // A ring slot holds the 8 columns of a group as [column][512 rows] (column stride TEAM_XCS).
            // For even nx the warp copies its 32 rows x 8 columns cooperatively in 16-byte pieces (two
            // consecutive rows of one column: half the asynchronous-copy operations of a per-element
            // copy, which ran at about one 8-byte element per clock per SM); a lane then reads its
            // fragment elements from what the warp copied, after __syncwarp.
            const bool wide = (nx & 1) == 0;
            constexpr int PF = 10;      // column groups the L2 prefetch runs ahead of the ring
            auto issue = [&](int g) {
                __syncwarp();  // the slot written below was read by the warp in the previous trip
                if (g < ngroups) {
                    double* slot = xs + (size_t)(((g / gstep) % ST) * GC) * TEAM_XCS + rslot;
                    if (wide) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int q = lane + 32 * u, col = q >> 4, r = wrow0 + 2 * (q & 15), j = g * GC + col;
                            const bool ok = r < nx && !row_in_block(r) && col_ok(j);
                            pb_cp_async16(slot + col * TEAM_XCS + r, ok ? A + r + (size_t)j * nx : A, ok ? 16 : 0);
                        }
                    } else {
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int r = rbase + 8 * mt, col = 2 * lq + e, j = g * GC + col;
                                const bool ok = r < nx && !row_in_block(r) && col_ok(j);
                                pb_cp_async8(slot + col * TEAM_XCS + r, ok ? A + r + (size_t)j * nx : A, ok ? 8 : 0);
                            }
                        }
                    }
                    // one lane per 64-byte row segment asks L2 for the same elements PF groups ahead
                    if (g + PF < ngroups && (lr == 0)) {
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int r = rbase + 8 * mt, j = (g + PF) * GC + 2 * lq + e;
                                if (r < nx && j < nx) pb_prefetch_l2(A + r + (size_t)j * nx);
                            }
                        }
                    }
                }
                pb_cp_async_commit();
            };
#pragma unroll
            for (int g = 0; g < ST - 1; ++g) issue(goff + g * gstep);
            for (int ch = 0; ch < nchunks; ++ch) {
                const double* const cur = sm + ((ch & 1) ? TEAM_SM_RBUF1 : TEAM_SM_RBUF);
                double* const nxt = sm + ((ch & 1) ? TEAM_SM_RBUF : TEAM_SM_RBUF1);
                if (ch + 1 < nchunks) {
                    // the pivot-row tile of the next chunk travels with the ring's next commit group
                    // (no registers): complete for this thread at its next wait, visible to the block
                    // after the barrier at the end of the chunk
#pragma unroll
                    for (int u = 0; u < TEAM_RBUF / 512; ++u) {
                        const int e = i + u * 512;
                        const int jj = e / NB, kk = e - jj * NB;
                        const int j = (ch + 1) * TEAM_CW + jj;
                        const bool ok = kk < nb && j < nx;
                        pb_cp_async8(nxt + r_pos(e), ok ? A + (k0 + kk) + (size_t)j * nx : A, ok ? 8 : 0);
                    }
                }
                for (int gl = goff; gl < TEAM_CW / GC; gl += gstep) {
                    const int g = ch * (TEAM_CW / GC) + gl;
                    if (g >= ngroups) break;
                    issue(g + (ST - 1) * gstep);
                    pb_cp_async_wait<ST - 1>();
                    __syncwarp();  // pieces copied by the other lanes of the warp
                    const double* src = xs + (size_t)(((g / gstep) % ST) * GC) * TEAM_XCS + (2 * lq) * TEAM_XCS + rslot + rbase;
                    double c[4][2];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        c[mt][0] = src[8 * mt];
                        c[mt][1] = src[TEAM_XCS + 8 * mt];
                    }
                    const double* rt = cur + (gl * GC + lr) * TEAM_RS + lq;  // B fragment: R[4ks + lq][column lr]
#pragma unroll
                    for (int ks = 0; ks < NB / 4; ++ks) {
                        const double bf = rt[4 * ks];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) pb_dmma(c[mt][0], c[mt][1], wfr[mt][ks], bf);
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int r = rbase + 8 * mt, j = g * GC + 2 * lq + e;
                            if (r < nx && col_ok(j)) A[r + (size_t)j * nx] = c[mt][e];
                        }
                    }
                }
                pb_cp_async_wait<ST - 1>();  // includes the tile copies issued at the top of this chunk
                __syncthreads();
            }
            pb_cp_async_wait<0>();
            __syncthreads();
            cyc_update += now() - t_p1;
        }
        if (rowok) w.perm[i] = perm_s[i];
        __syncthreads();
        return 0;
    }

    /// The solve GEMM of up to 8*NT right-hand sides for nx <= blockDim.x on the FP64 tensor cores.
    /// Warp w owns rows 32w..32w+31 and sweeps over K twice, once per half of 16 rows (two 8-row
    /// tiles x NT 8-column tiles accumulated in registers), so that every entry of the inverse is
    /// read from HBM/L2 exactly once for all right-hand sides; its fragments stream through the
    /// cp.async ring (four k-steps per group), the right-hand-side tile of the next 64 rows is
    /// fetched through the pivot order while the current one is used.
    template <int NT>
    __device__ __noinline__ void solve_pass_mma(int c0, int nc, double nsc) {
        constexpr int KC = 64, PRE = (KC * 8 * NT + 511) / 512;
        static_assert(8 * NT * TEAM_BS <= TEAM_RTILE, "tile does not fit");
        const int nx = this->dn;  // (locals: see invert_rows) dense dimension
        const int ldb = this->nx;  // leading dimension of the right-hand sides / solutions
        double* const scr = w.scr;
        const int i = threadIdx.x;
        const int lane = i & 31, lq = lane & 3, lr = lane >> 2;
        double* const sm = smem_base();
        const double* const A = w.A;
        const double* const b = w.b;
        const int* const perm = w.gidx;
        const int* const omap = lin_sparse ? SP.ord + SP.n1 : nullptr;  // solution row of block row r
        PB_ASSUME_GLOBAL(A); PB_ASSUME_GLOBAL(b); PB_ASSUME_GLOBAL(scr); PB_ASSUME_GLOBAL(perm);
        auto b_entry = [&](int ch, int e) -> double {
            const int cc = e / KC, kk = e - cc * KC;
            const int k = ch * KC + kk;
            return (e < KC * 8 * NT && k < nx && cc < nc) ? b[perm[k] + (size_t)(c0 + cc) * ldb] : 0.0;
        };
        auto b_pos = [&](int e) { return (e / KC) * TEAM_BS + (e % KC); };
        const int nchunks = (nx + KC - 1) / KC;
        constexpr int GK = 16, ST = TEAM_ST;  // 16 k's = four k-steps per ring group
        double* const xs = sm + TEAM_SM_XSTAGE;
        const int ngk = (nx + GK - 1) / GK;
        // nx <= 256: only half of the warps own rows -- two warps then share a 32-row group, one half
        // (16 rows) each, and the matrix is swept once instead of twice
        const bool split = nx <= 256;
        const int nsweeps = split ? 1 : 2;
        for (int sweep = 0; sweep < nsweeps; ++sweep) {
            const int half = split ? (i >> 8) : sweep;
            const int rbase = (split ? ((i >> 5) & 7) * 32 : (i & ~31)) + 16 * half + lr;
            double c[2][NT][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;
            __syncthreads();
            {
                double* const b0 = sm + TEAM_SM_RBUF;
#pragma unroll
                for (int u = 0; u < PRE; ++u) {
                    const int e = i + u * 512;
                    if (e < KC * 8 * NT) b0[b_pos(e)] = b_entry(0, e);
                }
            }
            __syncthreads();
            auto issue = [&](int g) {
                if (g < ngk) {
                    double* dst = xs + (size_t)((g % ST) * 8) * 512 + i;
#pragma unroll
                    for (int ks4 = 0; ks4 < 4; ++ks4) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const int r = rbase + 8 * mt, k = g * GK + 4 * ks4 + lq;
                            const bool ok = r < nx && k < nx;
                            pb_cp_async8(dst + (ks4 * 2 + mt) * 512, ok ? A + r + (size_t)k * nx : A, ok ? 8 : 0);
                        }
                    }
                }
                pb_cp_async_commit();
            };
#pragma unroll
            for (int g = 0; g < ST - 1; ++g) issue(g);
            for (int ch = 0; ch < nchunks; ++ch) {
                const double* const cur = sm + ((ch & 1) ? TEAM_SM_RBUF1 : TEAM_SM_RBUF);
                double* const nxt = sm + ((ch & 1) ? TEAM_SM_RBUF : TEAM_SM_RBUF1);
                double bpre[PRE];
                if (ch + 1 < nchunks) {
#pragma unroll
                    for (int u = 0; u < PRE; ++u) bpre[u] = b_entry(ch + 1, i + u * 512);
                }
                for (int gl = 0; gl < KC / GK; ++gl) {
                    const int g = ch * (KC / GK) + gl;
                    if (g >= ngk) break;
                    issue(g + ST - 1);
                    pb_cp_async_wait<ST - 1>();
                    const double* src = xs + (size_t)((g % ST) * 8) * 512 + i;
#pragma unroll
                    for (int ks4 = 0; ks4 < 4; ++ks4) {
                        const double a0 = src[(ks4 * 2) * 512], a1 = src[(ks4 * 2 + 1) * 512];
                        const double* bt = cur + lr * TEAM_BS + gl * GK + 4 * ks4 + lq;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            const double bf = bt[8 * nt * TEAM_BS];
                            pb_dmma(c[0][nt][0], c[0][nt][1], a0, bf);
                            pb_dmma(c[1][nt][0], c[1][nt][1], a1, bf);
                        }
                    }
                }
                if (ch + 1 < nchunks) {
#pragma unroll
                    for (int u = 0; u < PRE; ++u) {
                        const int e = i + u * 512;
                        if (e < KC * 8 * NT) nxt[b_pos(e)] = bpre[u];
                    }
                }
                __syncthreads();
            }
            pb_cp_async_wait<0>();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = rbase + 8 * mt, col = 8 * nt + 2 * lq + e;
                        if (r < nx && col < nc) scr[(omap ? omap[r] : r) + (size_t)(c0 + col) * ldb] = nsc * c[mt][nt][e];
                    }
                }
            }
        }
    }

    /// The solve GEMM for nx <= blockDim.x (one matrix row per thread): the right-hand-side tile of
    /// the next 64 rows is fetched (through the pivot order) while the current one is used, and the
    /// thread's next 8 matrix entries are in flight while the current 8 are multiplied.
    template <int TN>
    __device__ __noinline__ void solve_pass_rows(int c0, int nc, double nsc) {
        constexpr int KC = (TN == 1) ? 512 : 64;
        constexpr int PRE = (KC * TN + 511) / 512;
        static_assert(KC * TN <= TEAM_RBUF, "tile does not fit");
        const int nx = this->dn;  // (locals: see invert_rows) dense dimension
        const int ldb = this->nx;
        double* const scr = w.scr;
        const int i = threadIdx.x;
        const bool active = i < nx;
        double* const sm = smem_base();
        const double* const A = w.A;
        const double* const b = w.b;
        const int* const perm = w.gidx;
        const int* const omap = lin_sparse ? SP.ord + SP.n1 : nullptr;
        PB_ASSUME_GLOBAL(A); PB_ASSUME_GLOBAL(b); PB_ASSUME_GLOBAL(scr); PB_ASSUME_GLOBAL(perm);
        auto b_entry = [&](int ch, int e) -> double {
            const int kk = e / TN, cc = e - kk * TN;
            const int k = ch * KC + kk;
            return (k < nx && cc < nc && e < KC * TN) ? b[perm[k] + (size_t)(c0 + cc) * ldb] : 0.0;
        };
        double acc[TN];
#pragma unroll
        for (int cc = 0; cc < TN; ++cc) acc[cc] = 0.0;
        const int nchunks = (nx + KC - 1) / KC;
        __syncthreads();
        {
            double* const b0 = sm + TEAM_SM_RBUF;
#pragma unroll
            for (int u = 0; u < PRE; ++u) {
                const int e = i + u * 512;
                if (e < KC * TN) b0[e] = b_entry(0, e);
            }
        }
        __syncthreads();
        constexpr int GK = TEAM_GC, ST = TEAM_ST;
        double* const xs = sm + TEAM_SM_XSTAGE;
        const double* arow = A + i;
        const int ngk = (nx + GK - 1) / GK;  // groups of GK matrix columns (k indices)
        auto issue = [&](int g) {
            if (g < ngk) {
                double* dst = xs + (size_t)((g % ST) * GK) * 512 + i;
#pragma unroll
                for (int u = 0; u < GK; ++u) {
                    const int k = g * GK + u;
                    const bool ok = active && k < nx;
                    pb_cp_async8(dst + u * 512, ok ? arow + (size_t)k * nx : A, ok ? 8 : 0);
                }
            }
            pb_cp_async_commit();
        };
#pragma unroll
        for (int g = 0; g < ST - 1; ++g) issue(g);
        for (int ch = 0; ch < nchunks; ++ch) {
            const double* const cur = sm + ((ch & 1) ? TEAM_SM_RBUF1 : TEAM_SM_RBUF);
            double* const nxt = sm + ((ch & 1) ? TEAM_SM_RBUF : TEAM_SM_RBUF1);
            double bpre[PRE];
            if (ch + 1 < nchunks) {
#pragma unroll
                for (int u = 0; u < PRE; ++u) bpre[u] = b_entry(ch + 1, i + u * 512);
            }
            for (int gl = 0; gl < KC / GK; ++gl) {
                const int g = ch * (KC / GK) + gl;
                if (g >= ngk) break;
                issue(g + ST - 1);
                pb_cp_async_wait<ST - 1>();
                const double* src = xs + (size_t)((g % ST) * GK) * 512 + i;
#pragma unroll
                for (int u = 0; u < GK; ++u) {
                    const double a = src[u * 512];
                    if constexpr ((TN & 1) == 0) {
                        const double2* r2 = reinterpret_cast<const double2*>(cur + (gl * GK + u) * TN);
#pragma unroll
                        for (int c2 = 0; c2 < TN / 2; ++c2) {
                            const double2 rv = r2[c2];
                            acc[2 * c2] += a * rv.x;
                            acc[2 * c2 + 1] += a * rv.y;
                        }
                    } else {
                        const double* r = cur + (gl * GK + u) * TN;
#pragma unroll
                        for (int cc = 0; cc < TN; ++cc) acc[cc] += a * r[cc];
                    }
                }
            }
            if (ch + 1 < nchunks) {
#pragma unroll
                for (int u = 0; u < PRE; ++u) {
                    const int e = i + u * 512;
                    if (e < KC * TN) nxt[e] = bpre[u];
                }
            }
            __syncthreads();
        }
        pb_cp_async_wait<0>();
        if (active) {
#pragma unroll
            for (int cc = 0; cc < TN; ++cc)
                if (cc < nc) scr[(omap ? omap[i] : i) + (size_t)(c0 + cc) * ldb] = nsc * acc[cc];
        }
    }
#endif


    // ---------------------------------------------------------------- sparse path (sparse_lu.h)
    /// staging area of the sparse phases: shared memory from the broadcast tiles to the end of the
    /// block's dynamic shared memory (idle while no dense GEMM runs); a heap array in the host build
    PB_HD double* sp_stage() const {
#if defined(__CUDA_ARCH__)
        return smem_base() + TEAM_SM_RBUF;
#else
        return w.stage;
#endif
    }

    /// v - sum_e V[pl[e]] * V[pu[e]] over e in [e0, e1), products subtracted in ascending order; the loads of
    /// four products are issued together
    PB_HD static double pair_chain(double v, const double* V, const int* __restrict__ pl, const int* __restrict__ pu, int e0, int e1) {
        int e = e0;
        for (; e + 4 <= e1; e += 4) {
            const int l0 = pl[e], l1 = pl[e + 1], l2 = pl[e + 2], l3 = pl[e + 3];
            const int u0 = pu[e], u1 = pu[e + 1], u2 = pu[e + 2], u3 = pu[e + 3];
            const double a0 = V[l0], a1 = V[l1], a2 = V[l2], a3 = V[l3];
            const double b0 = V[u0], b1 = V[u1], b2 = V[u2], b3 = V[u3];
            v -= a0 * b0;
            v -= a1 * b1;
            v -= a2 * b2;
            v -= a3 * b3;
        }
        for (; e < e1; ++e) v -= V[pl[e]] * V[pu[e]];
        return v;
    }

    /// Numeric factorisation on the fixed pattern (the klu_refactor role): sparse part entry by entry
    /// over the dependency levels, then the Schur block assembled dense in w.A and inverted by the
    /// dense path.  KLU keeps the diagonal as pivot unless |diag| < 0.001 max|column| (klu_factor's
    /// tol; klu_refactor never re-pivots): the same bound is checked here on every multiplier, and
    /// a violation (or a zero / non-finite pivot) returns 2 -- the caller then factors this matrix
    /// by the dense path with partial pivoting instead.
    PB_HD int factor_sparse() {
        const int n1 = SP.n1, m = SP.m;
        // V in shared memory while the levels run (every entry is read by later levels)
        const bool v_staged = SP.nV <= w.stage_doubles;
        double* const V = v_staged ? sp_stage() : w.luV;
        const double* __restrict__ J = w.Jsaved;
        const double gam = gamma;
        bool bad = false;
        const long long t_f0 = now();
        tm.sync();
        for (int l = 0; l < SP.n_tlevels; ++l) {
            const int t1 = SP.tlevel_ptr[l + 1];
            // long pair lists (the late, dense-ish pivots): a warp per target, products spread over the lanes
            // (coalesced index loads), fixed shuffle tree; otherwise one thread per target
            const bool lng = SP.tlev_long[l] != 0;
            const int tstart = SP.tlevel_ptr[l] + (lng ? tm.warp() : tm.tid());
            const int tstep = lng ? tm.nwarps() : tm.size();
            for (int t = tstart; t < t1; t += tstep) {
                const int piv = SP.t_piv[t], src = SP.t_src[t];
                double v = (piv == -2) ? 1.0 : 0.0;
                if (src >= 0) v -= gam * J[src];
                if (lng) {
                    double acc = 0.0;
                    const int e1 = SP.t_pptr[t + 1];
                    PT_UNROLL4
                    for (int e = SP.t_pptr[t] + tm.lane(); e < e1; e += Team::WARP) acc += V[SP.p_l[e]] * V[SP.p_u[e]];
                    v -= tm.warp_sum(acc);
                } else {
                    v = pair_chain(v, V, SP.p_l, SP.p_u, SP.t_pptr[t], SP.t_pptr[t + 1]);
                }
                if (piv == -2) {
                    if (!(fabs(v) > 0.0) || !PB_FINITE(v)) bad = true;
                    v = 1.0 / v;
                } else if (piv >= 0) {
                    v *= V[piv];
                    if (!(fabs(v) <= 1000.0)) bad = true;  // |pivot| < 0.001 |entry|, or not finite
                }
                if (!lng || tm.lane() == 0) V[SP.t_dst[t]] = v;
            }
            tm.sync();
        }
        if (tm.any(bad)) return 2;
        // Schur block S = M22 - L21 U12, dense in w.A (leading dimension m)
        double* const A = w.A;
        for (size_t e = tm.tid(); e < (size_t)m * m; e += tm.size()) A[e] = 0.0;
        tm.sync();
        for (int t = tm.tid(); t < SP.n_starg; t += tm.size()) {
            const int src = SP.s_src[t];
            double v = SP.s_diag[t] ? 1.0 : 0.0;
            if (src >= 0) v -= gam * J[src];
            A[SP.s_dst[t]] = pair_chain(v, V, SP.p_l, SP.p_u, SP.s_pptr[t], SP.s_pptr[t + 1]);
        }
        if (v_staged) {
            PT_FOR(e, 0, SP.nV) w.luV[e] = V[e];
        }
        tm.sync();
        cyc_spfactor += now() - t_f0;
        dn = m;
        lin_sparse = true;
        if (m > 0) {
            const int r = invert_dense();
            if (r) return 2;
        }
        PT_FOR(k, 0, m) w.gidx[k] = SP.ord[n1 + w.perm[k]];
        tm.sync();
        (void)n1;
        return 0;
    }

    /// scr[:, c0 .. c0+nc) = nsc * M^-1 b[:, c0 .. c0+nc) through the sparse factors and the inverse
    /// of the Schur block:  forward substitution with L (sparse rows in the staging area, all columns
    /// of a batch side by side: one warp per row, lanes over the columns), Schur right-hand side,
    /// dense GEMM with S^-1, backward substitution with U.
    /// The factor values of the sparse part, copied to the END of the staging area for the duration of one
    /// substitution pass (every entry is read once per column batch, in a dependent chain per row: from
    /// shared memory that chain costs tens of cycles per entry instead of an L2 round trip).  zcap receives
    /// what is left for the column batches.  The caller synchronises before the first use.
    PB_HD const double* sp_stage_values(int& zcap) {
        zcap = w.stage_doubles;
        if (2 * SP.nV > w.stage_doubles) return w.luV;  // (too large: read through the caches)
        double* const Vs = sp_stage() + (w.stage_doubles - SP.nV);
        const double* __restrict__ src = w.luV;
        PT_FOR(e, 0, SP.nV) Vs[e] = src[e];
        zcap = w.stage_doubles - SP.nV;
        return Vs;
    }

    /// acc - sum_e V[v0 + e] * Z[col[e] * bc + cl] over e in [e0, e1), terms subtracted in ascending order; the
    /// index and value loads of four entries are issued together (they may come from global memory)
    PB_HD static double row_chain(double acc, const double* __restrict__ V, int v0, const unsigned short* __restrict__ col,
                                  const double* Z, int e0, int e1, int bc, int cl) {
        int e = e0;
        for (; e + 4 <= e1; e += 4) {
            const int c0 = col[e], c1 = col[e + 1], c2 = col[e + 2], c3 = col[e + 3];
            const double a0 = V[v0 + e], a1 = V[v0 + e + 1], a2 = V[v0 + e + 2], a3 = V[v0 + e + 3];
            const double z0 = Z[(size_t)c0 * bc + cl], z1 = Z[(size_t)c1 * bc + cl], z2 = Z[(size_t)c2 * bc + cl],
                         z3 = Z[(size_t)c3 * bc + cl];
            acc -= a0 * z0;
            acc -= a1 * z1;
            acc -= a2 * z2;
            acc -= a3 * z3;
        }
        for (; e < e1; ++e) acc -= V[v0 + e] * Z[(size_t)col[e] * bc + cl];
        return acc;
    }

    /// The rows [rows[0..nr)) (or n1 + 0..nr when rows == nullptr) of a level whose chains are LONG: a warp takes
    /// two columns of a row, the chain of each is split over 16 lanes (consecutive entries on consecutive
    /// lanes: coalesced factor loads) and reduced by a fixed shuffle tree.  scale != nullptr: result * scale[r].
    PB_HD void level_long_rows(double* Z, int bc, int nb, const double* __restrict__ V, int v0,
                               const int* __restrict__ rptr, const unsigned short* __restrict__ rcol,
                               const unsigned short* rows, int row0, int nr, const double* scale) {
        constexpr int P = Team::WARP >= 32 ? 16 : 1;  // lanes per chain
        constexpr int CPW = Team::WARP / P;           // columns per warp
        const int part = tm.lane() % P, which = tm.lane() / P;
        const int ncp = (nb + CPW - 1) / CPW;
        for (int task = tm.warp(); task < nr * ncp; task += tm.nwarps()) {
            const int rr = task / ncp, cp = task - rr * ncp;
            const int r = rows ? (int)rows[rr] : row0 + rr;
            const int cl = cp * CPW + which;
            const bool ok = cl < nb;
            double acc = 0.0;
            if (ok) {
                const int e1 = rptr[r + 1];
                PT_UNROLL4
                for (int e = rptr[r] + part; e < e1; e += P) acc += V[v0 + e] * Z[(size_t)rcol[e] * bc + cl];
            }
            acc = tm.sum16(acc);
            if (ok && part == 0) {
                const double v = Z[(size_t)r * bc + cl] - acc;
                Z[(size_t)r * bc + cl] = scale ? v * scale[r] : v;
            }
        }
    }

    /// forward levels + Schur rows on a staged batch Z[row][column] (stride bc) of nb columns
    PB_HD void sparse_forward_staged(double* Z, int bc, int nb, const double* V) {
        const int n1 = SP.n1, m = SP.m;
        // one thread per (row of the level, column): consecutive threads take consecutive columns of a row
        // (same factor entries: broadcast; consecutive words of Z: conflict-free)
        for (int l = 1; l < SP.n_flev; ++l) {  // (level 0: rows without L entries)
            const int r0 = SP.flev_ptr[l], nr = SP.flev_ptr[l + 1] - r0;
            if (SP.flev_long[l]) {
                level_long_rows(Z, bc, nb, V, n1, SP.Lr_ptr, SP.Lr_col, SP.flev_rows + r0, 0, nr, nullptr);
            } else {
                for (int t = tm.tid(); t < nr * nb; t += tm.size()) {
                    const int rr = t / nb, cl = t - rr * nb;
                    const int r = SP.flev_rows[r0 + rr];
                    Z[(size_t)r * bc + cl] = row_chain(Z[(size_t)r * bc + cl], V, n1, SP.Lr_col, Z, SP.Lr_ptr[r], SP.Lr_ptr[r + 1], bc, cl);
                }
            }
            tm.sync();
        }
        // Schur rows: gather from the finished y
        if (m > 0 && SP.flev_long[SP.n_flev]) {
            level_long_rows(Z, bc, nb, V, n1, SP.Lr_ptr, SP.Lr_col, nullptr, n1, m, nullptr);
        } else {
            for (int t = tm.tid(); t < m * nb; t += tm.size()) {
                const int rr = t / nb, cl = t - rr * nb;
                const int r = n1 + rr;
                Z[(size_t)r * bc + cl] = row_chain(Z[(size_t)r * bc + cl], V, n1, SP.Lr_col, Z, SP.Lr_ptr[r], SP.Lr_ptr[r + 1], bc, cl);
            }
        }
        tm.sync();
    }

    /// Forward substitution of the columns [cmin, cmax] of w.b in place: y = L11^-1 b1 in the sparse rows,
    /// b2' = b2 - L21 y in the Schur rows.  A batch of columns, ALL nx rows in permuted order, lives in the
    /// staging area as Z[row][column]: global memory is only touched by coalesced column sweeps (lanes along
    /// the original row index), the substitution runs in shared memory with the lanes along the columns.
    PB_HD void sparse_forward(int cmin, int cmax) {
        const int nx = this->nx;
        const int ncols = cmax - cmin + 1;
        const unsigned short* __restrict__ iord = SP.iord16;
        double* const Z = sp_stage();
        double* const b = w.b;
        tm.sync();
        int zcap;
        const double* const V = sp_stage_values(zcap);
        int bc_max = zcap / nx;
        if (bc_max < 1) bc_max = 1;
        const int nbatch = (ncols + bc_max - 1) / bc_max;
        const int bc = (ncols + nbatch - 1) / nbatch;
        for (int c0 = cmin; c0 <= cmax; c0 += bc) {
            const int nb = (cmax + 1 - c0 < bc) ? cmax + 1 - c0 : bc;
            tm.sync();
            PT_FOR(x, 0, nx * nb) {
                const int cl = x / nx, i = x - cl * nx;
                Z[(size_t)iord[i] * bc + cl] = b[i + (size_t)(c0 + cl) * nx];
            }
            tm.sync();
            sparse_forward_staged(Z, bc, nb, V);
            PT_FOR(x, 0, nx * nb) {
                const int cl = x / nx, i = x - cl * nx;
                b[i + (size_t)(c0 + cl) * nx] = Z[(size_t)iord[i] * bc + cl];
            }
        }
        tm.sync();
    }

    /// Backward substitution x1 = U11^-1 (y - U12 x2) of the columns [cmin, cmax] (y in the sparse rows of
    /// w.b, x2 in the Schur rows of w.scr).  update = false: scr = nsc * x.  update = true: the Newton
    /// update is fused -- acor += nsc * x and the WRMS norm of nsc * x (maximum over the columns) is
    /// returned; scr is not written.
    PB_HD double sparse_backward(int cmin, int cmax, double nsc, bool update) {
        const int n1 = SP.n1, nx = this->nx;
        const int ncols = cmax - cmin + 1;
        const unsigned short* __restrict__ iord = SP.iord16;
        double* const Z = sp_stage();
        const double* __restrict__ b = w.b;
        double* const scr = w.scr;
        double* const acor = w.acor;
        const double* __restrict__ ewt = w.ewt;
        tm.sync();
        int zcap;
        const double* const V = sp_stage_values(zcap);
        int bc_max = zcap / nx;
        if (bc_max < 1) bc_max = 1;
        const int nbatch = (ncols + bc_max - 1) / bc_max;
        const int bc = (ncols + nbatch - 1) / nbatch;
        const int nU0 = n1 + SP.nnzL;
        const bool azero = update && acor_zero_ && cmin >= 1;
        double loc = 0.0;
        for (int c0 = cmin; c0 <= cmax; c0 += bc) {
            const int nb = (cmax + 1 - c0 < bc) ? cmax + 1 - c0 : bc;
            tm.sync();
            {
                const double* __restrict__ xs2 = scr;
                PT_FOR(x, 0, nx * nb) {
                    const int cl = x / nx, i = x - cl * nx;
                    const int r = iord[i];
                    const size_t g = i + (size_t)(c0 + cl) * nx;
                    const double* __restrict__ src = (r < n1) ? b : xs2;
                    Z[(size_t)r * bc + cl] = src[g];
                }
            }
            tm.sync();
            for (int l = 0; l < SP.n_blev; ++l) {
                const int r0 = SP.blev_ptr[l], nr = SP.blev_ptr[l + 1] - r0;
                if (SP.blev_long[l]) {
                    level_long_rows(Z, bc, nb, V, nU0, SP.Ur_ptr, SP.Ur_col, SP.blev_rows + r0, 0, nr, V);
                } else {
                    for (int t = tm.tid(); t < nr * nb; t += tm.size()) {
                        const int rr = t / nb, cl = t - rr * nb;
                        const int r = SP.blev_rows[r0 + rr];
                        const double acc = row_chain(Z[(size_t)r * bc + cl], V, nU0, SP.Ur_col, Z, SP.Ur_ptr[r], SP.Ur_ptr[r + 1], bc, cl);
                        Z[(size_t)r * bc + cl] = acc * V[r];
                    }
                }
                tm.sync();
            }
            if (!update) {
                PT_FOR(x, 0, nx * nb) {
                    const int cl = x / nx, i = x - cl * nx;
                    scr[i + (size_t)(c0 + cl) * nx] = nsc * Z[(size_t)iord[i] * bc + cl];
                }
            } else {
                // flat and coalesced, a batch of independent loads ahead of the stores: acor += delta; the
                // squared weighted entries replace Z, so that the per-column sums can be formed in a FIXED order
                constexpr int U = 4;
                const int nel = nx * nb;
                PT_FOR_BATCH(xb, 0, nel, U) {
                    double a[U], ew[U], d[U];
                    int zi[U];
                    size_t g[U];
                    PB_UNROLL
                    for (int u = 0; u < U; ++u) {
                        const int x = xb + u * tm.size();
                        const bool ok = x < nel;
                        const int cl = ok ? x / nx : 0, i = ok ? x - cl * nx : 0;
                        g[u] = i + (size_t)(c0 + cl) * nx;
                        zi[u] = (int)iord[i] * bc + cl;
                        a[u] = (ok && !azero) ? acor[g[u]] : 0.0;  // (azero is uniform: the load is skipped for all)
                        ew[u] = ok ? ewt[g[u]] : 0.0;
                        d[u] = ok ? nsc * Z[zi[u]] : 0.0;
                    }
                    PB_UNROLL
                    for (int u = 0; u < U; ++u) {
                        const int x = xb + u * tm.size();
                        if (x < nel) {
                            acor[g[u]] = a[u] + d[u];
                            const double v = d[u] * ew[u];
                            Z[zi[u]] = v * v;
                        }
                    }
                }
                tm.sync();
                // one warp per column: lane-strided partial sums in ascending row order, then the shuffle tree
                for (int cl = tm.warp(); cl < nb; cl += tm.nwarps()) {
                    double sq = 0.0;
                    for (int i = tm.lane(); i < nx; i += Team::WARP) sq += Z[(size_t)iord[i] * bc + cl];
                    sq = tm.warp_sum(sq);
                    if (sq > loc) loc = sq;
                }
            }
        }
        tm.sync();
        if (!update) return 0.0;
        if (azero) acor_zero_ = false;  // acor now holds the first correction
        return sqrt(tm.max(loc) / (double)nx);
    }

    /// scr = nsc * (dense inverse) * b for the columns [cmin, cmax] (the whole matrix, or the Schur block
    /// with its row gather / scatter maps)
    PB_HD void solve_dense(int cmin, int cmax, double nsc) {
#if defined(__CUDA_ARCH__)
        const bool rows_path = dn <= tm.size();
#else
        const bool rows_path = false;
#endif
        if (cmin == cmax) {
#if defined(__CUDA_ARCH__)
            if (rows_path) solve_pass_rows<1>(cmin, 1, nsc);
            else
#endif
                solve_pass<1>(cmin, 1, nsc);
        } else {
#if defined(__CUDA_ARCH__)
            if (rows_path) {
                for (int c0 = cmin; c0 <= cmax; c0 += 56) {
                    const int nc = (cmax + 1 - c0 < 56) ? cmax + 1 - c0 : 56;
                    if (nc > 24) solve_pass_mma<7>(c0, nc, nsc);
                    else solve_pass_mma<3>(c0, nc, nsc);
                }
            } else
#endif
            {
                for (int c0 = cmin; c0 <= cmax; c0 += TEAM_TN) {
                    const int nc = (cmax + 1 - c0 < TEAM_TN) ? cmax + 1 - c0 : TEAM_TN;
                    solve_pass<TEAM_TN>(c0, nc, nsc);
                }
            }
        }
    }

    /// cvLsSetup + cvLsLinSys + the inversion; the Jacobian is evaluated at w.yj
    PB_HD int lsetup() {
        const double dgamma = fabs(gamma / gammap - 1.0);
        const bool jbad = (nst == 0) || (nst >= nstlj + cv::MSBJ) ||
                          ((convfail == cv::FAIL_BAD_J) && (dgamma < cv::DGMAX_JBAD)) || (convfail == cv::FAIL_OTHER);
        int ret = 0;
        tm.sync();
        if (jbad) {
            const bool ok = net::jac_values(T, tm, w.yj, w.p, w.Jsaved, true);
            if (tm.any(!ok)) ret = 1;  // AMICI's fJ returns checkFinite(J): recoverable
        }
        if (ret == 0) {
            if (SP.enabled && use_sparse) {
                ret = factor_sparse();
                if (ret == 2) {  // weak diagonal pivot: this matrix by the dense path with partial pivoting
                    nfallback++;
                    ret = invert_newton_matrix();
                }
            } else {
                ret = invert_newton_matrix();
            }
        }
        if (jbad) { nje++; nstlj = nst; }
        jcur = jbad;
        nsetups++;
        return ret;
    }

    /// scr[:, c0 .. c0+nc) = nsc * A * (P b)[:, c0 .. c0+nc): the solve GEMM.  Every thread owns
    /// rows of A (streamed from HBM/L2, coalesced across the warp); the right-hand-side tile is
    /// gathered through the pivot order into staging memory and broadcast.
    template <int TN>
    PB_HD void solve_pass(int c0, int nc, double nsc) {
        constexpr int KC = (TEAM_RBUF / TN) < 64 ? (TEAM_RBUF / TN) : ((TN == 1) ? TEAM_RBUF : 64);
        const int ldb = this->nx;
        const int nx = dn;  // (shadows the member: the dense dimension)
        const int* const omap = lin_sparse ? SP.ord + SP.n1 : nullptr;
        for (int row0 = 0; row0 < nx; row0 += tm.size()) {
            const int i = row0 + tm.tid();
            const bool active = i < nx;
            double acc[TN];
            PB_UNROLL
            for (int cc = 0; cc < TN; ++cc) acc[cc] = 0.0;
            for (int k0 = 0; k0 < nx; k0 += KC) {
                const int kc = (nx - k0 < KC) ? nx - k0 : KC;
                tm.sync();
                PT_FOR(e, 0, kc * TN) {
                    const int kk = e / TN, cc = e - kk * TN;
                    rbuf()[e] = (cc < nc) ? w.b[w.gidx[k0 + kk] + (size_t)(c0 + cc) * ldb] : 0.0;
                }
                tm.sync();
                if (active) {
                    const double* arow = w.A + i + (size_t)k0 * nx;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
                    for (int kk = 0; kk < kc; ++kk) {
                        const double a = arow[(size_t)kk * nx];
#if defined(__CUDA_ARCH__)
                        if constexpr ((TN & 1) == 0) {
                            const double2* r2 = reinterpret_cast<const double2*>(rbuf() + kk * TN);
                            PB_UNROLL
                            for (int c2 = 0; c2 < TN / 2; ++c2) {
                                const double2 rv = r2[c2];
                                acc[2 * c2] += a * rv.x;
                                acc[2 * c2 + 1] += a * rv.y;
                            }
                        } else
#endif
                        {
                            const double* r = rbuf() + kk * TN;
                            PB_UNROLL
                            for (int cc = 0; cc < TN; ++cc) acc[cc] += a * r[cc];
                        }
                    }
                }
            }
            if (active) {
                PB_UNROLL
                for (int cc = 0; cc < TN; ++cc)
                    if (cc < nc) w.scr[(omap ? omap[i] : i) + (size_t)(c0 + cc) * ldb] = nsc * acc[cc];
            }
        }
    }

    /// Newton update of the columns [cmin, cmax]: delta = -(I - gamma J)^-1 b with cvLsSolve's
    /// gamma-ratio scaling; acor += delta; returns the WRMS norm of delta.
    PB_HD double newton_update(int cmin, int cmax) {
        const double nsc = (gamrat != 1.0) ? -2.0 / (1.0 + gamrat) : -1.0;
        const long long t_s0 = now();
        if (lin_sparse) {
            // sparse factors + inverse of the Schur block (forward substitution already done by the fused
            // right-hand side of the sensitivity phase)
            if (acor_zero_ && cmin >= 1 && !fuse_newton) materialize_acor_zero();
            if (!b_forwarded_) sparse_forward(cmin, cmax);
            b_forwarded_ = false;
            const long long t_d0 = now();
            if (SP.m > 0) solve_dense(cmin, cmax, 1.0);
            tm.sync();
            const long long t_d1 = now();
            const double del = sparse_backward(cmin, cmax, nsc, fuse_newton);
            const long long t_e = now();
            cyc_sptri += (t_e - t_s0) - (t_d1 - t_d0);
            cyc_solve += t_e - t_s0;
            if (fuse_newton) return del;
        } else {
            if (acor_zero_ && cmin >= 1) materialize_acor_zero();
            solve_dense(cmin, cmax, nsc);
            tm.sync();
            cyc_solve += now() - t_s0;
        }
        {
            constexpr int U = 4;
            const int e1 = (cmax + 1) * nx;
            PT_FOR_BATCH(eb, cmin * nx, e1, U) {
                double a[U];
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    a[u] = (e < e1) ? w.acor[e] + w.scr[e] : 0.0;
                }
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    if (e < e1) w.acor[e] = a[u];
                }
            }
        }
        return wnorm1(cmin, cmax, w.scr);
    }

    // ---------------------------------------------------------------- BDF coefficients
    /// cvSetBDF + cvSetTqBDF + cvSet
    PB_HD void cvset() {
        double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
        l[0] = l[1] = xi_inv = xistar_inv = 1.0;
        for (int i = 2; i <= q; ++i) l[i] = 0.0;
        alpha0 = alpha0_hat = -1.0;
        hsum = h;
        if (q > 1) {
            for (int j = 2; j < q; ++j) {
                hsum += tau[j - 1];
                xi_inv = h / hsum;
                alpha0 -= 1.0 / (double)j;
                for (int i = j; i >= 1; --i) l[i] += l[i - 1] * xi_inv;
            }
            alpha0 -= 1.0 / (double)q;
            xistar_inv = -l[1] - alpha0;
            hsum += tau[q - 1];
            xi_inv = h / hsum;
            alpha0_hat = -l[1] - xi_inv;
            for (int i = q; i >= 1; --i) l[i] += l[i - 1] * xistar_inv;
        }
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        tq2 = fabs(A1 / (alpha0 * A2));
        tq5 = fabs(A2 * xistar_inv / (l[q] * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double C = xistar_inv / l[q];
                const double A3 = alpha0 + 1.0 / (double)q;
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = (1.0 - A4 + A3) / A3;
                tq1 = fabs(C * Cpinv);
            } else {
                tq1 = 1.0;
            }
            hsum += tau[q];
            xi_inv = h / hsum;
            const double A5 = alpha0 - 1.0 / (double)(q + 1);
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = (1.0 - A6 + A5) / A2;
            tq3 = fabs(Cppinv / (xi_inv * (q + 2) * A5));
        }
        tq4 = cv::CORTES / tq2;
        rl1 = 1.0 / l[1];
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? gamma / gammap : 1.0;
    }

    // ---------------------------------------------------------------- history array
    /// cvPredict (sgn = +1) / cvRestore (sgn = -1): Pascal-triangle sweep per element in registers
    PT_NOINLINE void shift_history(double sgn) {
        tm.sync();
        constexpr int U = 4;
        const int q = this->q;  // (locals: see invert_rows)
        const int e1 = nx * ncol;
        double* const zb = w.znbase;
        const size_t nxc = w.nxc;
        PT_FOR_BATCH(eb, 0, e1, U) {
            double z[U][cv::L_MAX];
            PB_UNROLL
            for (int u = 0; u < U; ++u) {
                const int e = eb + u * tm.size();
                PB_UNROLL
                for (int j = 0; j < cv::L_MAX; ++j) z[u][j] = (j <= q && e < e1) ? zb[j * nxc + e] : 0.0;
            }
            PB_UNROLL
            for (int u = 0; u < U; ++u) {
                PB_UNROLL
                for (int k = 1; k <= cv::Q_MAX; ++k) {
                    PB_UNROLL
                    for (int j = cv::Q_MAX; j >= k; --j)
                        if (j <= q && k <= q) z[u][j - 1] += sgn * z[u][j];
                }
            }
            PB_UNROLL
            for (int u = 0; u < U; ++u) {
                const int e = eb + u * tm.size();
                PB_UNROLL
                for (int j = 0; j < cv::Q_MAX; ++j)
                    if (j < q && e < e1) zb[j * nxc + e] = z[u][j];
            }
        }
        tm.sync();
    }
    /// cvRescale
    PT_NOINLINE void rescale() {
        tm.sync();
        constexpr int U = 4;
        const int q = this->q;  // (locals: see invert_rows)
        const double eta = this->eta;
        const int e1 = nx * ncol;
        double* const zb = w.znbase;
        const size_t nxc = w.nxc;
        PT_FOR_BATCH(eb, 0, e1, U) {
            double z[U][cv::L_MAX];
            PB_UNROLL
            for (int u = 0; u < U; ++u) {
                const int e = eb + u * tm.size();
                PB_UNROLL
                for (int j = 1; j < cv::L_MAX; ++j) z[u][j] = (j <= q && e < e1) ? zb[j * nxc + e] : 0.0;
            }
            PB_UNROLL
            for (int u = 0; u < U; ++u) {
                const int e = eb + u * tm.size();
                double factor = eta;
                PB_UNROLL
                for (int j = 1; j < cv::L_MAX; ++j) {
                    if (j <= q && e < e1) zb[j * nxc + e] = z[u][j] * factor;
                    factor *= eta;
                }
            }
        }
        h = hscale * eta;
        hscale = h;
        nscon = 0;
        tm.sync();
    }
    /// cvAdjustOrder -> cvIncreaseBDF / cvDecreaseBDF: the polynomial coefficients (scalars).  Returns
    /// false if the order change leaves the history untouched (q == 2, decrease).
    PB_HD bool adjust_order_coeffs(int deltaq, double* lc, double& A1, int& src, int& jmax) const {
        if ((q == 2) && (deltaq != 1)) return false;
        for (int i = 0; i <= cv::L_MAX; ++i) lc[i] = 0.0;
        lc[2] = 1.0;
        A1 = 0.0;
        if (deltaq == 1) {
            double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0;
            double hsum = hscale;
            for (int j = 1; j < q; ++j) {
                hsum += tau[j + 1];
                const double xi = hsum / hscale;
                prod *= xi;
                alpha0 -= 1.0 / (double)(j + 1);
                alpha1 += 1.0 / xi;
                for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xiold + lc[i - 1];
                xiold = xi;
            }
            A1 = (-alpha0 - alpha1) / prod;
            src = L;
            jmax = q;
        } else {
            double hsum = 0.0;
            for (int j = 1; j <= q - 2; ++j) {
                hsum += tau[j];
                const double xi = hsum / hscale;
                for (int i = j + 2; i >= 2; --i) lc[i] = lc[i] * xi + lc[i - 1];
            }
            for (int j = 2; j < q; ++j) lc[j] = -lc[j];
            src = q;
            jmax = q - 1;
        }
        return true;
    }
    PB_HD void adjust_order(int deltaq) {
        double lc[cv::L_MAX + 1], A1;
        int src, jmax;
        if (!adjust_order_coeffs(deltaq, lc, A1, src, jmax)) return;
        tm.sync();
        PT_FOR_ELEMS(e) {
            if (deltaq == 1) w.zn(src)[e] = A1 * w.zn(cv::Q_MAX)[e];
            const double zs = w.zn(src)[e];
            for (int j = 2; j <= jmax; ++j) w.zn(j)[e] += lc[j] * zs;
        }
        tm.sync();
    }

    // ---------------------------------------------------------------- nonlinear solve
    /// cvNls / cvStgrNls + SUNNonlinSolSolve_Newton + the convergence tests for the columns of
    /// `phase`; on success the corrected vector is zn[0] + acor.
    /// The residual of one Newton iteration: the columns of `phase` -- or, simultaneous corrector, the state column and
    /// then the sensitivity columns at the current state iterate with J and dxdotdp re-evaluated there (AMICI's
    /// fsxdot, model_ode.cpp:486-515; cvNlsResidualSensSim).  One call site of rhs() for both correctors.
    PB_HD int residual(int phase) {
        if constexpr (SS) {
            const bool sim = simultaneous_now(phase);
            int ret = 0;
            int ph_last = sim ? (int)PH_SENS : phase;
            PB_OPAQUE_INT(ph_last);
            for (int ph = phase; ph <= ph_last && ret == 0; ++ph) {
                if (sim && ph == PH_SENS) {
                    tm.sync();
                    PT_FOR(i, 0, nx) w.ysh[i] = w.yj[i];
                    refresh_Jcur();
                }
                ret = rhs(ph, false);
            }
            return ret;
        } else {
            return rhs(phase, false);
        }
    }
    /// true in the state phase of a step with the simultaneous corrector (steady-state instantiation only; the plain
    /// instantiation must not even declare a variable for it: its machine code is the measured one)
    PB_HD bool simultaneous_now(int phase) const {
        if constexpr (SS) return (phase == PH_STATE) && sensi && opt.simultaneous();
        else return false;
    }
    /// last column of the Newton iteration of `phase`
    PB_HD int nls_cmax(int phase) const {
        if constexpr (SS) return simultaneous_now(phase) ? ns : phase_cmax(phase);
        else return phase_cmax(phase);
    }
    /// norm of the accumulated correction for the error test, at convergence of iteration m_it
    PB_HD double nls_acnrm(int phase, int m_it, double del, int cmin, int cmax) {
        if constexpr (SS) {  // (simultaneous: the norm spans the columns that take part in the error test)
            const int ncmax = simultaneous_now(phase) ? errtest_cmax() : cmax;
            return (m_it == 0 && ncmax == cmax) ? del : wnorm1(cmin, ncmax, w.acor);
        } else {
            return (m_it == 0) ? del : wnorm1(cmin, cmax, w.acor);
        }
    }

    /// When simultaneous_now(phase) (steady-state instantiation only; SolverOptions::simultaneous()) the SIMULTANEOUS
    /// corrector: cvNls with do_sensi_sim (cvodes.c:6894-7007) over the wrapped vector of cvodes_nls_sim.c:374-560 --
    /// see Bdf::nls_simultaneous (bdf_core.cuh) for the scheme.  It runs through the SAME call sites as the staggered
    /// phases (the right-hand sides of both phases in turn, one setup, one update over the columns 0 .. ns, the joint
    /// norm), so the kernel does not carry a second copy of the solver; the fused sparse paths stay off.
    PB_HD int nls(int phase, int nflag) {
        const int cmin = phase_cmin(phase), cmax = nls_cmax(phase);
        bool setup_pending = false;
        if (phase == PH_STATE) {
            convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::NO_FAILURES : cv::FAIL_OTHER;
            setup_pending = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                            (nst >= nstlp + cv::MSBP) || (fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
        }
        tm.sync();
        // acor = 0 at the start of the iteration.  Sensitivity phase on the fused sparse path: the zeros are not
        // written (and not read back by the first right-hand side / update) -- acor_zero_ stands for them until
        // the first update stores acor = delta
        acor_zero_ = (phase == PH_SENS) && lazy_acor && fuse_newton && lin_sparse && !setup_pending &&
                     w.stage_doubles - (2 * SP.nV > w.stage_doubles ? 0 : SP.nV) >= 2 * nx;
        if (!acor_zero_) {
            PT_FOR(e, cmin * nx, (cmax + 1) * nx) w.acor[e] = 0.0;
        }
        bool jbad = false;
        bool top = true;
        int m_it = 0;
        for (;;) {
            fwd_ok_ = (phase == PH_SENS) && !setup_pending;
            int ret = residual(phase);
            fwd_ok_ = false;
            if (ret != 0 && top) return ret;
            if (ret == 0 && setup_pending) {
                if (jbad) convfail = cv::FAIL_BAD_J;
                if (phase == PH_SENS) {  // Jacobian at the converged state
                    tm.sync();
                    PT_FOR(i, 0, nx) w.yj[i] = w.ysh[i];
                }  // state phase: rhs() has just put zn[0]+acor into w.yj
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
                if (m_it > 0) cr = fmax(cv::CRDOWN * cr, del / delp);
                if (phase == PH_STATE) crate = cr;
                else crateS = cr;
                const double dcon = del * fmin(1.0, cr) / tq4;
                if (dcon <= 1.0) {
                    const double an = nls_acnrm(phase, m_it, del, cmin, cmax);
                    if (phase == PH_STATE) {
                        acnrm = an;
                        jcurNLS = false;
                    } else {
                        if (opt.errconS) acnrmS = an;
                        jcurNLSS = false;
                    }
                    jcur = false;
                    if constexpr (SS) {
                        if (simultaneous_now(phase)) publish_state();
                    }
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
            if ((ret > 0) && !((phase == PH_STATE) ? jcurNLS : jcurNLSS)) {
                setup_pending = true;
                jbad = true;
                top = true;
                m_it = 0;
                tm.sync();
                acor_zero_ = false;
                PT_FOR(e, cmin * nx, (cmax + 1) * nx) w.acor[e] = 0.0;
                continue;
            }
            return ret;
        }
    }

    // ---------------------------------------------------------------- step control
    /// cvCompleteStep, scalar part.  The element-wise part (zn[j] += l[j] * acor) is DEFERRED
    /// (pending_complete_): when another step of the same CVode call follows it is fused with that
    /// step's error weights, order change, rescaling and prediction into ONE pass over the history
    /// (advance_fused); otherwise complete_elements() runs it alone.  The norms that cvPrepareNextStep and
    /// cvBDFStab take of the completed history are formed as zn[j] + l[j] * acor on the fly.
    PB_HD void complete_step_scalars() {
        nst++;
        nscon++;
        qsum += q;
        for (int i = q; i >= 2; --i) tau[i] = tau[i - 1];
        if ((q == 1) && (nst > 1)) tau[2] = tau[1];
        tau[1] = h;
        qwait--;
        save_ = (qwait == 1) && (q != cv::Q_MAX);
        if (save_) saved_tq5 = tq5;
        pending_complete_ = true;
    }
    /// cvCompleteStep, element-wise part alone
    PT_NOINLINE void complete_elements() {
        const bool save = save_;
        tm.sync();
        {
            constexpr int U = 4;
            const int q = this->q;  // (locals: see invert_rows)
            const double* const acor = w.acor;
            const int e1 = nx * ncol;
            double* const zb = w.znbase;
            const size_t nxc = w.nxc;
            double lr[cv::L_MAX];  // l[] in registers (compile-time indices below)
            PB_UNROLL
            for (int j = 0; j < cv::L_MAX; ++j) lr[j] = l[j];
            PT_FOR_BATCH(eb, 0, e1, U) {
                double z[U][cv::L_MAX], a[U];
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    a[u] = (e < e1) ? acor[e] : 0.0;
                    PB_UNROLL
                    for (int j = 0; j < cv::L_MAX; ++j) z[u][j] = (j <= q && e < e1) ? zb[j * nxc + e] : 0.0;
                }
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    if (e < e1) {
                        PB_UNROLL
                        for (int j = 0; j < cv::L_MAX; ++j)
                            if (j <= q) zb[j * nxc + e] = z[u][j] + lr[j] * a[u];
                        if (save) zb[cv::Q_MAX * nxc + e] = a[u];
                    }
                }
            }
        }
        tm.sync();
        pending_complete_ = false;
    }

    /// The step that was just accepted is followed by another one of the same CVode call: ONE pass over
    /// the history does, per element and in CVODES' order of operations,
    ///   cvCompleteStep (zn[j] += l[j] acor) -> error weights of the new state (cvEwtSet; the CVode loop's
    ///   checks) -> cvAdjustOrder -> cvRescale -> cvPredict
    /// instead of five separate sweeps (7 + 2 + 2 + 10 + 11 array transfers at q = 5 become 13).  Scalars
    /// (step_begin, q, h, tn) are advanced here as attempt_predict would.  Returns 0 or a failure status.
    PT_NOINLINE int advance_fused() {
        const long long t_h0 = now();
        const int r = advance_fused_();
        cyc_hist += now() - t_h0;
        return r;
    }
    PT_NOINLINE int advance_fused_() {
        step_begin();
        // scalars of the order change / rescaling, from the state attempt_predict would see
        double lc[cv::L_MAX + 1], A1 = 0.0;
        int src = 0, jmax = 0;
        const int dq = dq_;
        const bool adj = (dq != 0) && adjust_order_coeffs(dq, lc, A1, src, jmax);
        const int q_old = q;
        const int q_new = q + dq;
        const bool resc = do_rescale_;
        const double eta_ = eta;
        const bool save = save_;
        bool ok = true;
        double s_all = 0.0;
        tm.sync();
        {
            constexpr int U = 4;
            const double* const acor = w.acor;
            const int e1 = nx * ncol;
            double* const zb = w.znbase;
            double* const ew = w.ewt;
            const size_t nxc = w.nxc;
            const double rtol = opt.rtol, atol = opt.atol;
            double lr[cv::L_MAX], lcr[cv::L_MAX + 1];
            PB_UNROLL
            for (int j = 0; j < cv::L_MAX; ++j) lr[j] = l[j];
            PB_UNROLL
            for (int j = 0; j <= cv::L_MAX; ++j) lcr[j] = adj ? lc[j] : 0.0;
            PT_FOR_BATCH(eb, 0, e1, U) {
                double z[U][cv::L_MAX], a[U];
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    a[u] = (e < e1) ? acor[e] : 0.0;
                    PB_UNROLL
                    for (int j = 0; j < cv::L_MAX; ++j) z[u][j] = (j <= q_old && e < e1) ? zb[j * nxc + e] : 0.0;
                    // order increase reads zn[Q_MAX] (saved correction); q_old < Q_MAX then, so slot Q_MAX is free
                    if (adj && dq == 1 && e < e1) z[u][cv::Q_MAX] = zb[cv::Q_MAX * nxc + e];
                }
                PB_UNROLL
                for (int u = 0; u < U; ++u) {
                    const int e = eb + u * tm.size();
                    if (e >= e1) continue;
                    // cvCompleteStep
                    const double zq = z[u][cv::Q_MAX];
                    PB_UNROLL
                    for (int j = 0; j < cv::L_MAX; ++j)
                        if (j <= q_old) z[u][j] = z[u][j] + lr[j] * a[u];
                    if (adj && dq == 1 && q_old < cv::Q_MAX) z[u][cv::Q_MAX] = zq;  // (not a history column yet)
                    if (save) zb[cv::Q_MAX * nxc + e] = a[u];
                    // cvEwtSet on the new state
                    const double wv = rtol * fabs(z[u][0]) + atol;
                    ok = ok && (wv > 0.0);
                    const double ewv = 1.0 / wv;
                    ew[e] = ewv;
                    const double zw = z[u][0] * ewv;
                    s_all += zw * zw;
                    // cvAdjustOrder
                    if (adj) {
                        if (dq == 1) {
                            const double zs = A1 * zq;
                            PB_UNROLL
                            for (int j = 0; j < cv::L_MAX; ++j) {
                                if (j == src) z[u][j] = zs;
                                if (j >= 2 && j <= jmax) z[u][j] += lcr[j] * zs;
                            }
                        } else {
                            double zs = 0.0;
                            PB_UNROLL
                            for (int j = 0; j < cv::L_MAX; ++j)
                                if (j == src) zs = z[u][j];
                            PB_UNROLL
                            for (int j = 0; j < cv::L_MAX; ++j)
                                if (j >= 2 && j <= jmax) z[u][j] += lcr[j] * zs;
                        }
                    }
                    // cvRescale
                    if (resc) {
                        double factor = eta_;
                        PB_UNROLL
                        for (int j = 1; j < cv::L_MAX; ++j) {
                            if (j <= q_new) z[u][j] = z[u][j] * factor;
                            factor *= eta_;
                        }
                    }
                    // cvPredict
                    PB_UNROLL
                    for (int k = 1; k <= cv::Q_MAX; ++k) {
                        PB_UNROLL
                        for (int j = cv::Q_MAX; j >= k; --j)
                            if (j <= q_new && k <= q_new) z[u][j - 1] += z[u][j];
                    }
                    PB_UNROLL
                    for (int j = 0; j < cv::L_MAX; ++j)
                        if (j <= q_new) zb[j * nxc + e] = z[u][j];
                }
            }
        }
        tm.sync();
        pending_complete_ = false;
        if (tm.any(!ok)) return SIM_ILL_INPUT;
        // CVode's "too much accuracy requested" check takes max over the columns of the per-column WRMS norm
        // of the new state; the sum over ALL columns formed above bounds it from above and is far below the
        // threshold (|y| ewt < 1/rtol) for any rtol above a few ulps
        s_all = tm.sum(s_all);
        if (cv::UROUND * sqrt(s_all / (double)nx) > 1.0) return SIM_TOO_MUCH_ACC;
        // scalars, as attempt_predict advances them
        if (dq != 0) {
            q += dq;
            L = q + 1;
            qwait = L;
            dq_ = 0;
        }
        if (resc) {
            h = hscale * eta_;
            hscale = h;
            nscon = 0;
            do_rescale_ = false;
        }
        tn += h;
        if (tstopset && ((tn - tstop) * h > 0.0)) tn = tstop;
        ahead_ = true;
        return 0;
    }

    /// cvPrepareNextStep with cvSetEta, cvComputeEtaqm1 / cvComputeEtaqp1 / cvChooseEta
    PB_HD void prepare_next_step(double dsm) {
        if (etamax == 1.0) {
            qwait = qwait > 2 ? qwait : 2;
            qprime = q;
            hprime = h;
            eta = 1.0;
            return;
        }
        etaq = 1.0 / (pow(cv::BIAS2 * dsm, 1.0 / (double)L) + cv::ADDON);
        qprime = q;
        eta = etaq;
        if (qwait == 0) {
            qwait = 2;
            const int ecmax = errtest_cmax();
            etaqm1 = 0.0;
            if (q > 1) {
                // (the completed zn[q] = zn[q] + l[q] * acor: cvCompleteStep's element-wise part is deferred)
                double ddn = pending_complete_ ? wnorm(0, ecmax, w.zn(q), 1.0, w.acor, l[q]) : wnorm1(0, ecmax, w.zn(q));
                ddn = ddn * tq1;
                etaqm1 = 1.0 / (pow(cv::BIAS1 * ddn, 1.0 / (double)q) + cv::ADDON);
            }
            etaqp1 = 0.0;
            if (q != cv::Q_MAX) {
                if (saved_tq5 != 0.0) {
                    const double hr = h / tau[2];
                    double hp = 1.0;
                    for (int i = 0; i < L; ++i) hp *= hr;  // SUNRpowerI
                    const double cquot = (tq5 / saved_tq5) * hp;
                    double dup = wnorm(0, ecmax, w.zn(cv::Q_MAX), -cquot, w.acor, 1.0);
                    dup = dup * tq3;
                    etaqp1 = 1.0 / (pow(cv::BIAS3 * dup, 1.0 / (double)(L + 1)) + cv::ADDON);
                }
            }
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
                tm.sync();
                PT_FOR(e, 0, (ecmax + 1) * nx) w.zn(cv::Q_MAX)[e] = w.acor[e];
                tm.sync();
            }
        }
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

    /// cvSLdet: every thread evaluates the (scalar) stability-limit detection redundantly
    PB_HD int sldet() const {
        const double rrcut = 0.98, vrrtol = 1.0e-4, vrrt2 = 5.0e-4, sqtol = 1.0e-3, rrtol = 1.0e-2;
        double rav[4], vrat[4], qc1[4], qc2[4], qc4[4], qc5[4], ssmax[4], smax[4], qkr[4], rrc[4], sqmx[4];
        double qjk[4][4], sigsq[4];
        const double (*ss)[4] = ssdat;
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
                const double rat = ss[i][k] / ss[i + 1][k];
                sumrat += rat;
                sumrsq += rat * rat;
            }
            rav[k] = 0.25 * sumrat;
            vrat[k] = fabs(0.25 * sumrsq - rav[k] * rav[k]);
            qc5[k] = ss[1][k] * ss[3][k] - ss[2][k] * ss[2][k];
            qc4[k] = ss[2][k] * ss[3][k] - ss[1][k] * ss[4][k];
            qc2[k] = ss[2][k] * ss[5][k] - ss[3][k] * ss[4][k];
            qc1[k] = ss[4][k] * ss[4][k] - ss[3][k] * ss[5][k];
        }
        int kflag = 0;
        double rr = 0.0;
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
            if (fabs(qc1[1]) < cv::TINY * ssmax[1]) return -4;
            double tem = qc1[2] / qc1[1];
            const double q22 = qc2[2] - tem * qc2[1];
            const double q42 = qc4[2] - tem * qc4[1], q52 = qc5[2] - tem * qc5[1];
            tem = qc1[3] / qc1[1];
            const double q23 = qc2[3] - tem * qc2[1];
            double q43 = qc4[3] - tem * qc4[1], q53 = qc5[3] - tem * qc5[1];
            if (fabs(q22) < cv::TINY * ssmax[2]) return -4;
            tem = q23 / q22;
            q43 -= tem * q42;
            q53 -= tem * q52;
            if (fabs(q43) < cv::TINY * ssmax[3]) return -4;
            rr = -q53 / q43;
            if (rr < cv::TINY || rr > 100.0) return -5;
            double sqmax = 0.0;
            for (int k = 1; k <= 3; ++k) {
                qkr[k] = qc5[k] + rr * (qc4[k] + rr * rr * (qc2[k] + rr * qc1[k]));
                sqmax = fmax(sqmax, fabs(qkr[k]) / ssmax[k]);
            }
            double sqmin = 0.0;
            if (sqmax < sqtol) {
                kflag = 2;
            } else {
                for (int it = 1; it <= 3; ++it) {
                    for (int k = 1; k <= 3; ++k) {
                        const double qp = qc4[k] + rr * rr * (3.0 * qc2[k] + rr * 4.0 * qc1[k]);
                        double drr = 0.0;
                        if (fabs(qp) > cv::TINY * ssmax[k]) drr = -qkr[k] / qp;
                        rrc[k] = rr + drr;
                    }
                    for (int k = 1; k <= 3; ++k) {
                        const double s = rrc[k];
                        double sqmaxk = 0.0;
                        for (int j = 1; j <= 3; ++j) {
                            qjk[j][k] = qc5[j] + s * (qc4[j] + s * s * (qc2[j] + s * qc1[j]));
                            const double saqj = fabs(qjk[j][k]) / ssmax[j];
                            if (saqj > sqmaxk) sqmaxk = saqj;
                        }
                        sqmx[k] = sqmaxk;
                    }
                    int kmin = 1;
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
                    }
                    for (int j = 1; j <= 3; ++j) qkr[j] = qjk[j][kmin];
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

    /// cvBDFStab
    PB_HD void bdf_stab() {
        if (q >= 3) {
            const double n0 = pending_complete_ ? wnorm(0, 0, w.zn(q), 1.0, w.acor, l[q]) : wnorm1(0, 0, w.zn(q));
            const double n1 = pending_complete_ ? wnorm(0, 0, w.zn(q - 1), 1.0, w.acor, l[q - 1]) : wnorm1(0, 0, w.zn(q - 1));
            int factorial = 1;
            for (int i = 1; i <= q - 1; ++i) factorial *= i;
            const double sq = factorial * q * (q + 1) * acnrm / fmax(tq5, cv::TINY);
            const double sqm1 = n0 * (factorial * q);
            const double sqm2 = n1 * factorial;
            for (int k = 1; k <= 3; ++k)
                for (int i = 5; i >= 2; --i) ssdat[i][k] = ssdat[i - 1][k];
            ssdat[1][1] = sqm2 * sqm2;
            ssdat[1][2] = sqm1 * sqm1;
            ssdat[1][3] = sq * sq;
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

    /// zn[1] = h * f(tn, zn[0]) for the state and (if any) the sensitivities; acor is scratch
    PB_HD int reload_zn1() {
        tm.sync();
        PT_FOR_ELEMS(e) w.acor[e] = 0.0;
        tm.sync();
        const int r = rhs_all_raw(sensi);
        if (r) return r;
        PT_FOR_ELEMS(e) w.zn(1)[e] = h * w.b[e];
        tm.sync();
        return 0;
    }

    PB_HD void step_begin() {
        ncf_ = nef_ = ncfS_ = nefS_ = 0;
        dsm_ = dsmS_ = 0.0;
        saved_t_ = tn;
        nflag_ = cv::FIRST_CALL;
        do_rescale_ = (nst > 0) && (hprime != h);
        dq_ = do_rescale_ ? (qprime - q) : 0;
        undo_ = restart_ = false;
    }

    /// order / step-size adjustment, (after a failed attempt: restore,) prediction, coefficients
    PB_HD int attempt_predict() {
        if (ahead_) {  // advance_fused has done order change, rescaling and prediction of this attempt
            ahead_ = false;
            cvset();
            fail_ = 0;
            ret_ = 0;
            return 0;
        }
        for (;;) {
            if (undo_) {
                tn = saved_t_;
            } else {
                if (restart_) {
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
                    const long long t_h0 = now();
                    rescale();
                    cyc_hist += now() - t_h0;
                    do_rescale_ = false;
                }
                tn += h;
                if (tstopset && ((tn - tstop) * h > 0.0)) tn = tstop;
            }
            {
                const long long t_h0 = now();
                shift_history(undo_ ? -1.0 : 1.0);
                cyc_hist += now() - t_h0;
            }
            if (!undo_) break;
            undo_ = false;
        }
        cvset();
        fail_ = 0;
        ret_ = 0;
        return 0;
    }

    PB_HD void attempt_correct(int phase) {
        const long long t_c0 = now();
        attempt_correct_(phase);
        if (phase == PH_STATE) cyc_state += now() - t_c0;
        else cyc_sens += now() - t_c0;
    }
    PB_HD void attempt_correct_(int phase) {
        // (steady-state instantiation: the simultaneous corrector does all columns in the state phase -- one Newton
        // iteration and ONE error test -- and nothing in the sensitivity phase)
        if (phase == PH_SENS) {
            if (fail_ || !sensi) return;
            if constexpr (SS) {
                if (opt.simultaneous()) return;
            }
            ncf_ = nef_ = 0;
            publish_state();
            PT_FOR(i, 0, nx) w.yj[i] = w.ysh[i];
            if (!eval_xdot()) { fail_ = 5; return; }
            refresh_Jcur();
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

    /// cvCompleteStep / cvPrepareNextStep / cvBDFStab on success, otherwise cvHandleNFlag and
    /// cvDoErrorTest.  Returns 0 (step completed), 1 (another attempt) or a negative failure code.
    PB_HD int attempt_finish() {
        const int fail = fail_;
        const int ret = ret_;
        if (fail == 0) {
            complete_step_scalars();
            if (!fuse_history) complete_elements();
            prepare_next_step(dsm_);
            if (opt.stldet) bdf_stab();
            etamax = (nst <= cv::SMALL_NST) ? cv::ETA_MAX_ES : cv::ETA_MAX_GS;
            return 0;
        }
        if (fail == 5) {
            nflag_ = cv::PREV_CONV_FAIL;
            return 1;
        }
        undo_ = true;
        if (fail == 1 || fail == 3) {
            if (fail == 1) { ncfn++; ncf_++; } else { ncfnS++; ncfS_++; }
            if (ret < 0) return ret;
            etamax = 1.0;
            if ((fabs(h) <= 0.0) || ((fail == 1 ? ncf_ : ncfS_) == cv::MXNCF)) {
                if (ret == cv::CONV_RECVR) return SIM_CONV_FAILURE;
                if (ret == cv::RHSFUNC_RECVR) return SIM_REPTD_RHSFUNC_ERR;
                return SIM_REPTD_SRHSFUNC_ERR;
            }
            eta = cv::ETA_CF;
            nflag_ = cv::PREV_CONV_FAIL;
            do_rescale_ = true;
        } else {
            int e;
            if (fail == 2) { e = ++nef_; netf++; } else { e = ++nefS_; netfS++; }
            nflag_ = cv::PREV_ERR_FAIL;
            if ((fabs(h) <= 0.0) || (e == cv::MXNEF)) return SIM_ERR_FAILURE;
            etamax = 1.0;
            if (e <= cv::MXNEF1) {
                eta = 1.0 / (pow(cv::BIAS2 * (fail == 2 ? dsm_ : dsmS_), 1.0 / (double)L) + cv::ADDON);
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
    /// cvHin with cvUpperBoundH0 and cvYddNorm; zn[1] holds y'(t0) unscaled
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
        double hub_inv = 0.0;
        tm.sync();
        PT_FOR(e, 0, (ecmax + 1) * nx) {
            const double den = cv::HUB_FACTOR * fabs(w.zn(0)[e]) + 1.0 / w.ewt[e];
            const double v = fabs(w.zn(1)[e]) / den;
            if (v > hub_inv) hub_inv = v;
        }
        hub_inv = tm.max(hub_inv);
        double hub = cv::HUB_FACTOR * tdist;
        if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
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
                const double hgs = hg * sign;
                tm.sync();
                PT_FOR_ELEMS(e) w.acor[e] = hgs * w.zn(1)[e];
                // AMICI's right-hand sides do not depend on t explicitly for network models
                retval = rhs_all_raw(withS);
                if (retval == 0) {
                    const double hinv = 1.0 / hgs;
                    yddnrm = wnorm(0, ecmax, w.b, hinv, w.zn(1), -hinv);
                    hgOK = true;
                    break;
                }
                hg *= 0.2;
            }
            tm.sync();
            PT_FOR_ELEMS(e) w.acor[e] = 0.0;
            tm.sync();
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

    /// Prologue of CVode(tout) in CV_NORMAL mode with tstop = tout, as AMICI drives it
    PB_HD int cvode_begin(double tout, bool onestep = false) {
        // Solver::run sets tstop = tout; Solver::step (CV_ONE_STEP, simulation to steady state:
        // deps/AMICI/src/solver.cpp:161-175) sets none
        const bool one = SS && onestep;
        if (!one) {
            tstop = tout;
            tstopset = true;
        }
        if (opt.maxsteps <= nst) return SIM_ERROR;
        mxstep_ = opt.maxsteps - nst;
        nstloc_ = 0;
        if (nst == 0) {
            if (!ewt_set()) return SIM_ILL_INPUT;
            h = 1.0;
            const int r = reload_zn1();
            if (r) return r == cv::RHSFUNC_RECVR ? SIM_FIRST_RHSFUNC_ERR : SIM_FIRST_SRHSFUNC_ERR;
            if (!one && (tstop - tn) * (tout - tn) <= 0.0) return SIM_ILL_INPUT;
            const int hr = hin(tout);
            if (hr != 0) return hr;
            if (!one && (tn + h - tstop) * h > 0.0) h = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
            hscale = h;
            hprime = h;
            tm.sync();
            PT_FOR_ELEMS(e) w.zn(1)[e] *= h;
            tm.sync();
        } else if (!one) {
            const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
            if (fabs(tn - tstop) <= troundoff) {
                tstopset = false;
                tn = tstop;
                return cv::TSTOP_RETURN;
            } else if ((tn + hprime - tstop) * h > 0.0) {
                hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                eta = hprime / h;
            }
        }
        return 0;
    }

    // ---------------------------------------------------------------- steady states (SS only)
    /// Convergence test of a simulation to steady state (SteadystateProblem::getWrms / getWrmsFSA,
    /// deps/AMICI/src/steadystateproblem.cpp:519-626): wrms(xdot(x)) < 1 with weights 1 / (rtol |x| + atol) and,
    /// with sensitivities, wrms(J sx_c + dxdotdp_c) < 1 with 1 / (rtol |sx_c| + atol) for the columns in
    /// order, stopping at the first above the threshold; x / sx = zn[0] (the solver's state at tn); the
    /// tolerances are 100 x the integration tolerances (AMICI's steady-state defaults, solver.h:1908-1923).
    /// Scratch: yj, ysh, vb, v, fout, Jcur, dfdp -- all rewritten by the next step attempt.
    PB_HD bool ss_converged() {
        const double rtol = 100.0 * opt.rtol, atol = 100.0 * opt.atol;
        const double* __restrict__ z0 = w.zn(0);
        tm.sync();
        PT_FOR(i, 0, nx) {
            w.yj[i] = z0[i];
            w.ysh[i] = z0[i];
        }
        eval_xdot();
        nfe--;  // (Model::fxdot called by the steady-state problem itself: not one of CVODES' evaluations)
        tm.sync();
        double s = 0.0;
        PT_FOR(i, 0, nx) {
            const double v = w.fout[i] * (1.0 / (rtol * fabs(z0[i]) + atol));
            s += v * v;
        }
        s = tm.sum(s);
        if (!(sqrt(s / (double)nx) < 1.0)) return false;
        if (!sensi) return true;
        refresh_Jcur();
        const double* __restrict__ jc = w.Jcur;
        const int* __restrict__ jcol = T.jell_col;
        const int jw = T.jell_w;
        bool above = false, last_ok = true;
        for (int c = 1 + tm.warp(); c <= ns; c += tm.nwarps()) {  // one warp per column, lanes over the rows
            const double* __restrict__ sx = z0 + (size_t)c * nx;
            const double* __restrict__ dcol = w.dfdp + (size_t)w.plist[c - 1] * nx;
            double sc = 0.0;
            for (int i = tm.lane(); i < nx; i += Team::WARP) {
                double F = dcol[i];
                for (int t = 0; t < jw; ++t) {
                    const int j = jcol[t * nx + i];
                    if (j >= 0) F += jc[t * nx + i] * sx[j];
                }
                const double v = F * (1.0 / (rtol * fabs(sx[i]) + atol));
                sc += v * v;
            }
            sc = tm.warp_sum(sc);
            const double wr = sqrt(sc / (double)nx);
            if (wr > 1.0) above = true;
            if (c == ns && !(wr < 1.0)) last_ok = false;
        }
        return !tm.any(above || !last_ok);
    }

    /// One internal step: the body of the CVode loop (cvodes.c:3366-3605) in CV_ONE_STEP mode; the step is
    /// completed alone (the convergence test needs the corrected history).  Returns 0 or a failure status.
    PB_HD int ss_step() {
        if (nst > 0 && !ewt_set()) return SIM_ILL_INPUT;
        if (nstloc_ >= mxstep_) return SIM_TOO_MUCH_WORK;
        const double nrm = wnorm1(0, errtest_cmax(), w.zn(0));
        if (cv::UROUND * nrm > 1.0) return SIM_TOO_MUCH_ACC;
        step_begin();
        for (;;) {
            int a = attempt_predict();
            if (a < 0) return a;
            attempt_correct(PH_STATE);
            attempt_correct(PH_SENS);
            a = attempt_finish();
            if (a < 0) return a;
            if (a == 0) break;
        }
        nstloc_++;
        if (pending_complete_) {
            const long long t_h0 = now();
            complete_elements();
            cyc_hist += now() - t_h0;
        }
        return 0;
    }

    /// SteadystateProblem::runSteadystateSimulation (steadystateproblem.cpp:711-794): test, then one step
    /// towards max(t, 1) * 10, until converged.  Any failure ends as AMICI_ERROR (:198-255, 397-425).
    /// Out of line: ONE copy of the step machinery beside the main loop's, shared by pre- and post-equilibration.
    PT_NOINLINE int to_steady_state() {
        for (;;) {
            if (ss_converged()) return SIM_SUCCESS;
            if (nst >= opt.maxsteps) return SIM_ERROR;
            if (cvode_begin(fmax(tn, 1.0) * 10.0, true) < 0) return SIM_ERROR;
            if (ss_step() < 0) return SIM_ERROR;
        }
    }

    /// Pre-equilibration (amici.cpp:95-108; workSteadyStateProblem with it = -1): w.k holds the constants of
    /// the PRE-EQUILIBRATION condition; from x0 / sx0 to steady state, which stays in zn[0].
    PB_HD int preequilibrate(const ConditionIO& io, double rtol, double atol) {
        begin(io, rtol, atol);
        return finish(to_steady_state());
    }

    // ---------------------------------------------------------------- likelihood
    /// One output time: llh and sllh contributions (Model::addObservableObjective /
    /// addObservableObjectiveSensitivity); the state at the output time is zn[0] of column 0.
    /// Observables are linear in the state, sigmas are constants (dydp = dsigmaydp = 0).
    PB_HD bool data_point(const double* my, const double* sd, double* y_out, double* sigmay_out) {
        tm.sync();
        const int ny = T.ny;
        double dJydy[TEAM_MAXNY];
        bool ok = true;
        for (int iy = 0; iy < ny; ++iy) {
            const double y = net::observable(T, iy, w.zn(0));
            double sig = T.sigma0[iy];
            if (!isnan(sd[iy])) sig = sd[iy];
            if (tm.tid() == 0) {
                if (y_out) y_out[iy] = y;
                if (sigmay_out) sigmay_out[iy] = sig;
            }
            dJydy[iy] = 0.0;
            if (isnan(my[iy])) continue;
            if (!(sig > 0.0) || !PB_FINITE(sig)) { ok = false; continue; }
            const double s2 = sig * sig;
            const double r = my[iy] - y;
            llh_ -= log(s2 * 3.141592653589793 * 2.0) * 0.5 + 1.0 / s2 * (r * r) * 0.5;
            dJydy[iy] = 1.0 / s2 * (my[iy] * 2.0 - y * 2.0) * -0.5;
        }
        if (!ok) return false;
        if (!sensi) return true;
        PT_FOR(c, 1, ns + 1) {
            const double* sx = w.zn(0) + (size_t)c * nx;
            double a = 0.0;
            for (int iy = 0; iy < ny; ++iy) {
                if (dJydy[iy] == 0.0) continue;
                double d = 0.0;
                for (int e = T.o_ptr[iy]; e < T.o_ptr[iy + 1]; ++e) d += T.o_w[e] * sx[T.o_idx[e]];
                a += dJydy[iy] * d;
            }
            w.sllh[c] -= a;
        }
        return true;
    }

    // ---------------------------------------------------------------- one simulation
    /// w.p / w.k / w.plist must be set (and synchronised) by the caller
    ///
    /// keep_state (SS only; after preequilibrate(), w.k now the simulation condition's constants): the main
    /// simulation starts from the state and sensitivities in zn[0] (ForwardProblem, forwardproblem.cpp:59-63);
    /// the states of the list reinit[0..n_reinit) whose initial value depends on the constants are reset to it,
    /// their sensitivities to sx0 = 0 (Model::reinitialize, model.cpp:277-291,1849-1923).
    PB_HD void begin(const ConditionIO& io, double rtol, double atol, bool keep_state = false,
                     const int* reinit = nullptr, int n_reinit = 0) {
        io_ = io;
        opt.rtol = rtol;
        opt.atol = atol;
        nx = T.nx;
        ns = io.ns;
        ncol = ns + 1;
        sensi = ns > 0;
        tm.sync();
        if (SS && keep_state) {
            for (int e = tm.tid(); e < n_reinit; e += tm.size()) {
                const int i = reinit[e];
                const int ki = T.x0_kidx[i];
                if (ki < 0) continue;
                w.zn(0)[i] = T.x0_base[i] * w.k[ki];
                for (int c = 1; c < ncol; ++c) w.zn(0)[(size_t)c * nx + i] = 0.0;
            }
        } else {
            net::x0(T, tm, w.k, w.zn(0));
            PT_FOR(e, nx, nx * ncol) w.zn(0)[e] = 0.0;  // sx0 = 0: initial values do not depend on p
        }
        PT_FOR_ELEMS(e) {
            for (int j = 1; j < cv::L_MAX; ++j) w.zn(j)[e] = 0.0;
            w.acor[e] = 0.0;
            w.ewt[e] = 1.0;
            w.b[e] = 0.0;
        }
        PT_FOR(c, 0, ncol) w.sllh[c] = 0.0;
        for (size_t e = tm.tid(); e < (size_t)nx * T.np; e += tm.size()) w.dfdp[e] = 0.0;
        for (int j = 0; j < cv::L_MAX; ++j) l[j] = 0.0;
        for (int i = 0; i <= cv::L_MAX; ++i) tau[i] = 0.0;
        for (int i = 0; i < 6; ++i)
            for (int k = 0; k < 4; ++k) ssdat[i][k] = 0.0;
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
        save_ = pending_complete_ = ahead_ = false;
        nfallback = 0;
        dn = nx;
        lin_sparse = false;
        llh_ = 0.0;
        it_ = 0;
        sim_status_ = SIM_SUCCESS;
        fail_ = ret_ = 0;
        cyc_panel = cyc_update = cyc_solve = cyc_rhs = cyc_spfactor = cyc_sptri = cyc_hist = cyc_state = cyc_sens = 0;
        cyc_total = now();
        tm.sync();
    }

    /// ForwardProblem::workForwardProblem: output times in order, CVode(tout) with tstop = tout
    PB_HD int simulate(const ConditionIO& io, double rtol, double atol, bool keep_state = false,
                       const int* reinit = nullptr, int n_reinit = 0) {
        begin(io, rtol, atol, keep_state, reinit, n_reinit);
        const int ny = T.ny;
        for (it_ = 0; it_ < io_.nt; ++it_) {
            tout_ = io_.t[it_];
            if (!PB_FINITE(tout_)) {
                // t = inf: post-equilibration (amici.cpp:109-114; ForwardProblem stops at the first infinite output
                // time, forwardproblem.cpp:149-150).  Without steady-state support an error, never a silently
                // dropped point.
                if (!SS) return finish(SIM_ERROR);
                if (to_steady_state() != SIM_SUCCESS) return finish(SIM_ERROR);
                // ReturnData::processPostEquilibration (rdata.cpp:234-258): every t = inf point at the steady state
                for (; it_ < io_.nt; ++it_) {
                    double* yo = io_.y_out ? io_.y_out + (size_t)it_ * ny : nullptr;
                    double* so = io_.sigmay_out ? io_.sigmay_out + (size_t)it_ * ny : nullptr;
#if defined(PB200_HOST_TRACE)
                    if (io_.trace) { io_.trace[2 * it_] = nst; io_.trace[2 * it_ + 1] = q; }
#endif
                    if (!data_point(io_.my + (size_t)it_ * ny, io_.sd + (size_t)it_ * ny, yo, so)) return finish(SIM_ERROR);
                }
                return finish(SIM_SUCCESS);
            }
            bool integrate = (tout_ > opt.t0) && (tn < tout_);
            if (integrate) {  // is_next_t_too_close
                const double tdiff = tout_ - tn;
                if (fabs(tdiff) < 2.0 * cv::UROUND * fmax(fabs(tn), fabs(tout_))) integrate = false;
            }
            if (integrate) {
                int r = cvode_begin(tout_);
                if (r < 0) return finish(r);
                while (r != cv::TSTOP_RETURN) {
                    // top of the CVode step loop (after a fused advance its checks are already done)
                    if (!ahead_) {
                        if (nst > 0 && !ewt_set()) return finish(SIM_ILL_INPUT);
                        if (nstloc_ >= mxstep_) return finish(SIM_TOO_MUCH_WORK);
                        const double nrm = wnorm1(0, errtest_cmax(), w.zn(0));
                        if (cv::UROUND * nrm > 1.0) return finish(SIM_TOO_MUCH_ACC);
                        step_begin();
                    }
                    for (;;) {
                        int a = attempt_predict();
                        if (a < 0) return finish(a);
                        attempt_correct(PH_STATE);
                        attempt_correct(PH_SENS);
                        a = attempt_finish();
                        if (a < 0) return finish(a);
                        if (a == 0) break;
                    }
                    nstloc_++;
                    const double troundoff = cv::FUZZ_FACTOR * cv::UROUND * (fabs(tn) + fabs(h));
                    if (fabs(tn - tstop) <= troundoff) {
                        tn = tstop;
                        tstopset = false;
                        break;
                    }
                    if ((tn + hprime - tstop) * h > 0.0) {
                        hprime = (tstop - tn) * (1.0 - 4.0 * cv::UROUND);
                        eta = hprime / h;
                    }
                    if ((tn - tout_) * h >= 0.0) break;
                    // another step of this CVode call follows: complete this one and set up the next in one pass
                    if (pending_complete_) {
                        if (nstloc_ >= mxstep_) return finish(SIM_TOO_MUCH_WORK);
                        const int af = advance_fused();
                        if (af < 0) return finish(af);
                    }
                }
                if (pending_complete_) {  // the output time is reached: the corrected history
                    const long long t_h0 = now();
                    complete_elements();
                    cyc_hist += now() - t_h0;
                }
            }
#if defined(PB200_HOST_TRACE)
            if (io_.trace) { io_.trace[2 * it_] = nst; io_.trace[2 * it_ + 1] = q; }
#endif
            double* yo = io_.y_out ? io_.y_out + (size_t)it_ * ny : nullptr;
            double* so = io_.sigmay_out ? io_.sigmay_out + (size_t)it_ * ny : nullptr;
            if (!data_point(io_.my + (size_t)it_ * ny, io_.sd + (size_t)it_ * ny, yo, so)) return finish(SIM_ERROR);
        }
        return finish(SIM_SUCCESS);
    }

    PB_HD int finish(int status) {
        sim_status_ = status;
        if (status != SIM_SUCCESS) llh_ = NAN;
        tm.sync();
        return status;
    }
    PB_HD int status() const { return sim_status_; }
    PB_HD double llh() const { return llh_; }

    PB_HD void export_stats(int64_t* s, int trials) const {
        s[STAT_NST] = nst; s[STAT_NFE] = nfe; s[STAT_NSETUPS] = nsetups; s[STAT_NJE] = nje;
        s[STAT_NNI] = nni; s[STAT_NNIS] = nniS; s[STAT_NFSE] = nfSe; s[STAT_NETF] = netf;
        s[STAT_NETFS] = netfS; s[STAT_NCFN] = ncfn; s[STAT_NCFNS] = ncfnS; s[STAT_NOR] = nor;
        s[STAT_TRIALS] = trials; s[STAT_QSUM] = qsum;
        s[14] = ((cyc_panel >> 10) << 32) | ((cyc_update >> 10) & 0xffffffffll);
        s[15] = ((cyc_solve >> 10) << 32) | ((cyc_rhs >> 10) & 0xffffffffll);
        s[STAT_NFALLBACK] = nfallback;
        s[STAT_CYC_SPARSE] = ((cyc_spfactor >> 10) << 32) | ((cyc_sptri >> 10) & 0xffffffffll);
        s[18] = ((cyc_hist >> 10) << 32) | ((cyc_state >> 10) & 0xffffffffll);
        s[19] = ((cyc_sens >> 10) << 32) | (((now() - cyc_total) >> 10) & 0xffffffffll);
    }
#undef PT_FOR
#undef PT_FOR_ELEMS
};

}  // namespace pb200
#endif
