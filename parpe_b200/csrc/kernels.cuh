// kernels.cuh -- the CUDA kernels of the hot path and the per-model launch table.
//
//   simulate_kernel<M,G,CPL>  persistent, one lane group per condition, dynamic
//                             work queue (replaces AmiciSimulationRunner +
//                             LoadBalancerMaster/Worker + messageHandler +
//                             runAndLogSimulation for one GPU: SURVEY.md rows
//                             A2-A4, A8-A14)
//   reduce_kernel             fixed-order reduction of the per-condition results
//                             into (nllh, #errors, gradient) -- aggregateLikelihood
//                             (row A6)
#ifndef PB200_KERNELS_CUH
#define PB200_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <string>
#include <type_traits>

#include "bdf_core.cuh"

namespace pb200 {

/// Device-resident problem tables + per-evaluation buffers handed to the kernels.
struct KernelArgs {
    // --- resident tables (uploaded once at pb200_create) ---
    int n_conditions, n_opt, np, nk, ny;
    const int64_t* t_offsets;      // [C+1]
    const double* t;               // [sum nt]
    const double* my;              // [sum nt * ny]
    const double* sd;              // [sum nt * ny]
    const double* k;               // [C*nk]
    const int* mapping;            // [C*np]
    const int* pscale_sim;         // [C*np]
    const int* pscale_opt;         // [n_opt]
    const double* ov_scaled;       // [C*np] parameter overrides (simulation scale) or null
    const double* ov_unscaled;     // [C*np] same, unscaled on the host at create time
    const int* plist;              // [C*np] sensitivity parameter list per condition
    const int* nplist;             // [C]
    const int* inv_j;              // [C*np] optimisation indices a condition maps to, ascending (padding: INT_MAX)
    const int* inv_i;              // [C*np] the model parameter of each entry of inv_j (ascending within equal j)
    // steady states (pb200_set_preequilibration / t = inf output times): null / 0 without them
    const double* k_preeq;         // [C*nk] constants of the pre-equilibration condition
    const int* preeq_flag;         // [C] 1 = the condition is pre-equilibrated
    const unsigned* reinit_mask;   // [C] states re-initialised after pre-equilibration (bit i = state i; nx <= 32)
    const int64_t* reinit_off;     // [C+1] the same as a ragged list of state indices (large-model kernel)
    const int* reinit_idx;
    int steady;                    // 1 = some condition has a pre-equilibration or a t = inf output time:
                                   //     the steady-state instantiation of the kernel is launched
    // --- per evaluation ---
    const double* theta_tab;       // [n_opt*3*2]: for optimisation parameter j and simulation
                                   // scale s: (value in simulation scale, unscaled value)
    const int* idx;                // [n] condition index of every position of this partition
    const int* order;              // [n] processing order of positions (cost-descending) or null
    int n;                         // positions in this partition
    int want_grad;
    int* counter;                  // work-queue head
    // --- per-position outputs ---
    double* llh;                   // [n]
    double* sllh;                  // [n*np] plist-compacted, simulation scale (A7 result package)
    double* gcontrib;              // [n*np] -chain-ruled gradient contribution per model parameter
    int* status;                   // [n]
    int64_t* stats;                // [n*NSTATS]
    double* y_out;                 // layout of my[] (by condition) or null
    double* sigmay_out;
    // --- options ---
    int sync_period;               // CTA barrier every this many step attempts (>= 1)
    int stage_sync;                // CTA barrier (instead of a warp barrier) between the stages of an attempt
    int ctas_per_sm_cap;           // experiments (pb200_set_option): 0 = occupancy limit
    int max_groups;                // lane groups of a CTA that pull work (launch geometry: few conditions are
                                   // spread over all SMs instead of filling the first CTAs)
    int linsol;                    // large-model path: 0 automatic, 1 dense inverse, 2 sparse LU + dense Schur block
    int fuse_off;                  // large-model path, experiments: bit 0 no fused history pass, bit 1 no fused Newton
                                   // iteration, bit 2 acor zeroed eagerly (default 0: everything on)
    SolverOptions opt;
    int num_trials;
    double relaxation_factor;
};

constexpr double PB_LN10 = 2.302585092994046;
// register cap: shared memory holds ~24 conditions per SM for a Boehm-sized model, so few warps
// are resident and each may use up to 255 registers
#define PB_MIN_BLOCKS_FOR(T, G) ((T) <= 32 ? 8 : ((T) <= 64 ? 4 : ((T) <= 128 ? 2 : 1)))

/// One lane group per condition, persistent CTAs, in-kernel work queue.  The group's working
/// set (GroupShared<M>: Nordsieck history of all columns, Jacobian, inverse Newton matrix, ...)
/// lives in dynamic shared memory; the integrator's scalars live in registers.
///
/// Every trip of the main loop advances each group by ONE step attempt, in four stages
/// (Bdf::stage_pre / stage_nls(state) / stage_nls(sensitivities) / stage_post) with a warp
/// barrier between them, and -- with CTA_SYNC -- a CTA barrier per attempt: the warps of the CTA
/// then walk through the same instructions at about the same time.  The step loop is far
/// larger than the SM's instruction caches and is streamed from L2 every attempt
/// (profiles/README.md); a line fetched for one warp is reused by the others only while they
/// are close behind.
///
/// SS = true is the instantiation with steady-state support: output times
/// t = inf (post-equilibration) and a pre-equilibration phase per condition
/// (amici.cpp:95-114: preeq -> forward problem -> posteq); it is launched only for problems that
/// have either, the SS = false kernel is the code without any of it.
template <class M, int G, int THREADS, bool CTA_SYNC, bool SS = false>
__global__ void __launch_bounds__(THREADS, PB_MIN_BLOCKS_FOR(THREADS, G)) simulate_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) unsigned char pb_smem_raw[];
    LaneGroup<G> grp;
    GroupShared<M>* sh = reinterpret_cast<GroupShared<M>*>(pb_smem_raw) + (threadIdx.x / G);
    PB_ASSUME_SHARED(sh);
    using BdfT = Bdf<M, LaneGroup<G>, SS>;
    BdfT bdf;
    bool preeq_phase = false;  // SS: the pre-equilibration of the current condition is running
    bdf.sh = sh;
    bdf.opt = a.opt;
    const int lane = grp.lane();

    // Launch geometry: with fewer conditions than lane groups on the GPU only the first a.max_groups
    // groups of every CTA pull work, numbered round-robin over the warps (so that the busy groups are
    // spread over the SM's four schedulers): the conditions cover all SMs instead of filling the first CTAs.
    bool active = false;
    bool done;
    {
        constexpr int GPW = 32 / G;  // groups per warp
        const int g = threadIdx.x / G, wv = g / GPW, slot = g % GPW;
        done = (slot * (THREADS / 32) + wv) >= a.max_groups;
    }
    int since_sync = 0;
    int r = 0, trial = 1;
    size_t cp = 0;
    ConditionIO io;
    for (;;) {
        __syncwarp();  // the groups of a warp start every step attempt together
        if (!active && !done) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(a.counter, 1);
            pos = grp.bcast0(pos);
            if (pos >= a.n) {
                done = true;
            } else {
                r = a.order ? a.order[pos] : pos;
                const int c = a.idx[r];
                cp = (size_t)c * M::NP;
                // ---- A5: optimisation -> simulation parameters (integer gather; the values come
                //      from the host-computed table so they are bit-identical to the reference's
                //      getScaledParameter(getUnscaledParameter(.)) chain) ----
                grp.sync();
                for (int i = lane; i < M::NP; i += G) {
                    const int j = a.mapping[cp + i];
                    double pu;
                    if (j >= 0) pu = a.theta_tab[((size_t)j * 3 + a.pscale_sim[cp + i]) * 2 + 1];
                    else if (a.ov_unscaled) pu = a.ov_unscaled[cp + i];
                    else pu = NAN;
                    sh->p[i] = pu;
                    sh->plist[i] = a.plist[cp + i];
                }
                if constexpr (SS) preeq_phase = a.preeq_flag && a.preeq_flag[c] != 0;
                {
                    // pre-equilibration runs with the constants of the pre-equilibration condition
                    // (ConditionContext, deps/AMICI/src/edata.cpp:511-524)
                    const double* ksrc = (SS && preeq_phase) ? a.k_preeq : a.k;
                    for (int i = lane; i < M::NK; i += G) sh->k[i] = ksrc[(size_t)c * M::NK + i];
                }
                grp.sync();
                const int64_t t0off = a.t_offsets[c];
                io.nt = (int)(a.t_offsets[c + 1] - t0off);
                io.t = a.t + t0off;
                io.my = a.my + t0off * M::NY;
                io.sd = a.sd + t0off * M::NY;
                io.ns = a.want_grad ? a.nplist[c] : 0;
                io.y_out = a.y_out ? a.y_out + t0off * M::NY : nullptr;
                io.sigmay_out = a.sigmay_out ? a.sigmay_out + t0off * M::NY : nullptr;
                trial = 1;
                if constexpr (SS) bdf.begin(io, a.opt.rtol, a.opt.atol, preeq_phase ? BdfT::SS_PREEQ : BdfT::SS_OFF);
                else bdf.begin(io, a.opt.rtol, a.opt.atol);
                active = true;
            }
        }
        // one step attempt in four stages; the groups of the warp re-converge between them, so
        // whatever a group did in one stage (Jacobian setup, a second Newton iteration, an
        // order change ...) costs the others that stage only
        if (active) bdf.stage_pre();
        for (int ph_ = 0; ph_ <= 1; ++ph_) {
            int ph = ph_;
            PB_OPAQUE_INT(ph);
            if (CTA_SYNC && a.stage_sync) __syncthreads();
            else __syncwarp();
            if (active) bdf.stage_nls(ph);
        }
        if (CTA_SYNC && a.stage_sync) __syncthreads();
        else __syncwarp();
        if (active && bdf.stage_post()) {
            const int status = bdf.status();
            bool resumed = false;
            if constexpr (SS) {
                if (preeq_phase && status == SIM_SUCCESS) {
                    // ---- pre-equilibrated: the main simulation starts from the steady state with the
                    //      simulation condition's constants (ForwardProblem, forwardproblem.cpp:59-63,
                    //      119-120), at the tolerances of the current trial ----
                    const int c = a.idx[r];
                    double relax = 1.0;
                    for (int e = 1; e < trial; ++e) relax *= a.relaxation_factor;
                    grp.sync();
                    for (int i = lane; i < M::NK; i += G) sh->k[i] = a.k[(size_t)c * M::NK + i];
                    grp.sync();
                    preeq_phase = false;
                    bdf.begin(io, a.opt.rtol * relax, a.opt.atol * relax, BdfT::SS_POSTEQ,
                              a.reinit_mask ? a.reinit_mask[c] : 0u);
                    resumed = true;
                }
            }
            if (resumed) {
            } else if (status != SIM_SUCCESS && trial < a.num_trials) {
                // ---- A8: runAndLogSimulation retry (multiConditionProblem.cpp:369-465): trial t
                //      runs at tolerances * relaxation^(t-1) relative to the template ----
                double relax = 1.0;
                for (int e = 0; e < trial; ++e) relax *= a.relaxation_factor;
                ++trial;
                if constexpr (SS) {
                    const int c = a.idx[r];
                    preeq_phase = a.preeq_flag && a.preeq_flag[c] != 0;
                    if (preeq_phase) {  // the whole simulation is repeated, pre-equilibration included
                        grp.sync();
                        for (int i = lane; i < M::NK; i += G) sh->k[i] = a.k_preeq[(size_t)c * M::NK + i];
                        grp.sync();
                    }
                    bdf.begin(io, a.opt.rtol * relax, a.opt.atol * relax, preeq_phase ? BdfT::SS_PREEQ : BdfT::SS_OFF);
                } else {
                    bdf.begin(io, a.opt.rtol * relax, a.opt.atol * relax);
                }
            } else {
                // ---- outputs: A7 result package + A6 chain rule to the optimiser scale ----
                if (lane == 0) {
                    a.llh[r] = bdf.llh();
                    a.status[r] = status;
                    bdf.export_stats(a.stats + (size_t)r * NSTATS, trial);
                }
                if (a.want_grad) {
                    for (int i = lane; i < M::NP; i += G) {
                        a.gcontrib[(size_t)r * M::NP + i] = 0.0;
                        a.sllh[(size_t)r * M::NP + i] = 0.0;
                    }
                    grp.sync();
                    for (int col = lane; col <= io.ns; col += G) {
                        if (col < 1) continue;
                        const int ip = sh->plist[col - 1];
                        const int ssim = a.pscale_sim[cp + ip];
                        const int jopt = a.mapping[cp + ip];
                        const double pu = sh->p[ip];
                        // ReturnData::applyChainRuleFactorToSimulationResults (rdata.cpp:672-686, 706-708)
                        double s = sh->sllh[col];
                        if (ssim == 2) s *= pu * PB_LN10;
                        else if (ssim == 1) s *= pu;
                        a.sllh[(size_t)r * M::NP + (col - 1)] = s;
                        // applyChainRule (multiConditionDataProvider.cpp:762-796), coefficient -1
                        const int sopt = a.pscale_opt[jopt];
                        double gr = s;
                        if (ssim != sopt) {
                            if (ssim == 2) gr /= pu * PB_LN10;
                            else if (ssim == 1) gr /= pu;
                            if (sopt == 2) gr *= pu * PB_LN10;
                            else if (sopt == 1) gr *= pu;
                        }
                        a.gcontrib[(size_t)r * M::NP + ip] = -1.0 * gr;
                    }
                }
                grp.sync();
                active = false;
            }
        }
        if constexpr (CTA_SYNC) {
            if (++since_sync >= a.sync_period) {
                since_sync = 0;
                if (__syncthreads_and(done ? 1 : 0)) break;
            }
        } else {
            if (done) break;
        }
    }
}

/// Deterministic reduction: block 0 -> nllh and error count; block 1+j -> gradient entry j.
/// out[0] = sum_r -llh[r]; out[1] = #conditions with status != 0; out[2+j] = gradient.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) reduce_kernel(const KernelArgs a, double* __restrict__ out) {
    __shared__ double red[THREADS];
    __shared__ double red2[THREADS];
    const int tid = threadIdx.x;
    double acc = 0.0, acc2 = 0.0;
    if (blockIdx.x == 0) {
        for (int r = tid; r < a.n; r += THREADS) {
            acc -= a.llh[r];
            acc2 += (a.status[r] != 0) ? 1.0 : 0.0;
        }
    } else {
        // gradient entry j: the contributions of every position whose condition maps a model
        // parameter to j, found by binary search in the condition's sorted inverse list (O(log np) per
        // position instead of np; summation order as before: positions by stride, model parameters ascending)
        const int j = blockIdx.x - 1;
        if (a.want_grad) {
            for (int r = tid; r < a.n; r += THREADS) {
                const int* __restrict__ lj = a.inv_j + (size_t)a.idx[r] * a.np;
                const int* __restrict__ li = a.inv_i + (size_t)a.idx[r] * a.np;
                const size_t rp = (size_t)r * a.np;
                int lo = 0, hi = a.np;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (lj[mid] < j) lo = mid + 1;
                    else hi = mid;
                }
                for (; lo < a.np && lj[lo] == j; ++lo) acc += a.gcontrib[rp + li[lo]];
            }
        }
    }
    red[tid] = acc;
    red2[tid] = acc2;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            red[tid] += red[tid + s];
            red2[tid] += red2[tid + s];
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (blockIdx.x == 0) {
            out[0] = red[0];
            out[1] = red2[0];
        } else {
            out[1 + blockIdx.x] = red[0];
        }
    }
}

inline void launch_reduce(const KernelArgs& a, double* d_out, cudaStream_t stream) {
    reduce_kernel<256><<<1 + a.n_opt, 256, 0, stream>>>(a, d_out);
}

// ---- per-model launch table ------------------------------------------------------
struct ModelOps {
    const char* name;
    int nx, np, nk, ny;
    int default_lanes;  // G used for gradient evaluations unless overridden
    int flops_xdot, flops_J, flops_dxdotdp, nnzJ;
    // launches simulate_kernel; lanes: requested G (0 = default); returns cudaError_t as int
    int (*launch)(const KernelArgs& a, int lanes, int variant, int sm_count, cudaStream_t stream, int* lanes_used,
                  int* grid_used, int* block_used, void* state);
    // optional per-handle device state (large-model path: device copy of the model tables + the
    // teams' HBM workspace); null for the small-model kernels
    void* (*create_state)(int sm_count, std::string* err);
    void (*destroy_state)(void* state);
    // large-model path: description of the linear solver chosen at create (null otherwise)
    int (*linsol_info)(void* state, int64_t* out, int n);
    // 1 = the model's kernels handle pre-equilibration and t = inf output times (KernelArgs::steady)
    int supports_steady;
    // 1 = ... and Newton's method for steady states (SolverOptions::newton_maxsteps > 0)
    int supports_newton;
};

void register_model(const ModelOps& ops);
const ModelOps* find_model(const char* name);
int model_count();
const ModelOps* model_at(int i);

template <class M, int G, int THREADS, bool CTA_SYNC, bool SS = false>
int launch_one(const KernelArgs& a, int sm_count, cudaStream_t stream, int* grid_used, int* block_used) {
    auto kern = simulate_kernel<M, G, THREADS, CTA_SYNC, SS>;
    const int groups_per_block = THREADS / G;
    const size_t smem = sizeof(GroupShared<M>) * groups_per_block;
    // the attribute is per device (context): set before every launch (cheap), not cached per process
    if (smem > 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    if (a.ctas_per_sm_cap > 0 && per_sm > a.ctas_per_sm_cap) per_sm = a.ctas_per_sm_cap;
    int grid = sm_count * per_sm;  // persistent: one resident wave, work queue inside
    if (grid > a.n) grid = a.n;
    if (grid < 1) grid = 1;
    // n < (lane groups of a resident wave): spread the conditions over all SMs -- ceil(n / grid) groups
    // per CTA pull work (the time of a step attempt is per SM and nearly independent of the number of
    // resident conditions: DESIGN.md 3.1)
    KernelArgs b = a;
    b.max_groups = (a.n + grid - 1) / grid;
    if (b.max_groups > groups_per_block) b.max_groups = groups_per_block;
    kern<<<grid, THREADS, smem, stream>>>(b);
    if (grid_used) *grid_used = grid;
    if (block_used) *block_used = THREADS;
    return (int)cudaGetLastError();
}

constexpr int pb_pow2_at_least(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

/// Instantiations offered per model: lane groups of GN = pow2 >= NX lanes (one state index per
/// lane, the default) or GN/2 (two per lane, twice the conditions per warp); `variant` selects
/// the CTA shape: 0 = two warps with a barrier per step attempt (default), 1 = one free-running
/// warp per CTA.  Cost-only evaluations use the same kernels with the state column alone.
/// threads of the "one CTA per SM" shape: as many lane groups as fit the SM's shared memory,
/// whole warps, at most 1024 threads
template <class M, int G>
constexpr int pb_big_cta_threads() {
    int groups = (int)((227u * 1024u) / sizeof(GroupShared<M>));
    if (groups > 1024 / G) groups = 1024 / G;
    const int per_warp = 32 / G;
    groups -= groups % per_warp;
    if (groups < per_warp) groups = per_warp;
    return groups * G;
}

/// CTA shapes: variant 2 (default) = ONE CTA per SM holding as many lane groups as fit the SM's
/// shared memory, all warps meeting at a barrier per step attempt (the kernel streams its
/// instructions from L2 -- profiles/README.md -- and warps that walk through the same code
/// together share the fetched lines); variant 0 = two-warp CTAs with the same barrier.
template <class M, int G>
int launch_shape(const KernelArgs& a, int variant, int sm_count, cudaStream_t stream, int* grid_used, int* block_used) {
    if (variant == 0) return launch_one<M, G, (G < 32 ? 64 : 128), true>(a, sm_count, stream, grid_used, block_used);
    return launch_one<M, G, pb_big_cta_threads<M, G>(), true>(a, sm_count, stream, grid_used, block_used);
}

/// Lane-group widths offered per model: GN = pow2 >= NX lanes (one state index per lane, the
/// default), GN/2 (two state indices per lane, twice the conditions per warp) and 2*GN (two
/// column groups).  Cost-only evaluations use the same kernels with the state column alone.
template <class M>
int launch_model(const KernelArgs& a, int lanes, int variant, int sm_count, cudaStream_t stream, int* lanes_used,
                 int* grid_used, int* block_used, void* /*state*/) {
    constexpr int GN = pb_pow2_at_least(M::NX);
    constexpr int GH = GN >= 4 ? GN / 2 : GN;
    constexpr int G2 = GN * 2 <= 32 ? GN * 2 : GN;
    // default: one state index per lane; two per lane when that fills every lane (nx a power of two
    // >= 8): half the warps walk the instruction stream for the same conditions (measured on the
    // Boehm shape: +4.5 %; four per lane: -40 %)
    int G = ((M::NX & (M::NX - 1)) == 0 && M::NX >= 8) ? GH : GN;
    if (a.steady) {
        // problems with a pre-equilibration or t = inf output times: the steady-state instantiation, offered
        // in the default shape only (default lane-group width, one CTA per SM)
        constexpr int GS = ((M::NX & (M::NX - 1)) == 0 && M::NX >= 8) ? GH : GN;
        if (lanes_used) *lanes_used = GS;
        return launch_one<M, GS, pb_big_cta_threads<M, GS>(), true, true>(a, sm_count, stream, grid_used, block_used);
    }
    if (lanes == GN || lanes == GH || lanes == G2) G = lanes;
    if (lanes_used) *lanes_used = G;
    if (G == GH && GH != GN) return launch_shape<M, GH>(a, variant, sm_count, stream, grid_used, block_used);
    if (G == G2 && G2 != GN) return launch_shape<M, G2>(a, variant, sm_count, stream, grid_used, block_used);
    return launch_shape<M, GN>(a, variant, sm_count, stream, grid_used, block_used);
}

template <class M>
ModelOps make_ops() {
    ModelOps o;
    o.name = M::NAME;
    o.nx = M::NX; o.np = M::NP; o.nk = M::NK; o.ny = M::NY;
    o.default_lanes = ((M::NX & (M::NX - 1)) == 0 && M::NX >= 8) ? pb_pow2_at_least(M::NX) / 2 : pb_pow2_at_least(M::NX);
    o.flops_xdot = M::FLOPS_XDOT; o.flops_J = M::FLOPS_J; o.flops_dxdotdp = M::FLOPS_DXDOTDP; o.nnzJ = M::NNZ;
    o.launch = &launch_model<M>;
    o.create_state = nullptr;
    o.destroy_state = nullptr;
    o.linsol_info = nullptr;
    o.supports_steady = 1;
    o.supports_newton = 1;
    return o;
}

#define PB200_REGISTER_MODEL(MODEL)                                  \
    namespace {                                                      \
    struct Reg_##MODEL {                                             \
        Reg_##MODEL() { pb200::register_model(pb200::make_ops<pb200::models::MODEL>()); } \
    } reg_instance_##MODEL;                                          \
    }

}  // namespace pb200
#endif
