// team_kernels.cuh -- the large-model path: one thread block (team) per experimental
// condition, persistent CTAs with an in-kernel work queue (the same scheduler as
// simulate_kernel: SURVEY.md rows A2-A4, A8-A14, for BASELINE configs 3-4).  The model is a
// table-driven mass-action network (net_model.cuh) resident in HBM; the integrator is
// bdf_team.cuh; the per-team working set lives in an HBM workspace owned by the handle.
#ifndef PB200_TEAM_KERNELS_CUH
#define PB200_TEAM_KERNELS_CUH

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bdf_team.cuh"
#include "kernels.cuh"

namespace pb200 {

constexpr int TEAM_THREADS = 512;

/// SS = true: the instantiation with steady-state support (pre-equilibration, t = inf output times), launched
/// only for problems that have either (KernelArgs::steady); SS = false is the code without any of it.
template <int TEAM_THREADS_, bool SS = false>
__global__ void __launch_bounds__(TEAM_THREADS_, 1)
simulate_team_kernel(const KernelArgs a, const NetTables T, const SparseLU SP, double* __restrict__ workspace,
                     size_t ws_stride, int smem_doubles, int sp_tab_off, int use_sparse) {
    extern __shared__ __align__(16) double pb_team_smem[];
    __shared__ int s_pos;
    CtaTeam tm;
    BdfTeam<CtaTeam, SS> bdf(tm);
    bdf.T = T;
    bdf.SP = SP;
    bdf.use_sparse = use_sparse != 0;
    bdf.fuse_history = !(a.fuse_off & 1);
    bdf.fuse_newton = !(a.fuse_off & 2);
    bdf.lazy_acor = !(a.fuse_off & 4);
    bdf.opt = a.opt;
    bdf.w.carve(workspace + (size_t)blockIdx.x * ws_stride, T, SP.enabled ? SP.nV : 0);
    bdf.w.NB = 0;  // panel width of the generic dense path: chosen per inversion from the panel capacity
    // dynamic shared memory: [fixed buffers | tiles | ring / panel / sparse staging ...  | sparse solve tables]
    const int stage_end = sp_tab_off >= 0 ? sp_tab_off : smem_doubles;
    bdf.w.stage_doubles = stage_end - TEAM_SM_RBUF;
    bdf.w.panel_doubles = stage_end - TEAM_SM_PANEL;
    if (sp_tab_off >= 0) {
        // the row pointers and 16-bit indices of the sparse triangular solves live in shared memory for the
        // lifetime of the (persistent) block: every solve walks them, entry after entry
        int* ip = reinterpret_cast<int*>(pb_team_smem + sp_tab_off);
        const int n_lp = SP.nx + 1, n_up = SP.n1 + 1, n_fp = SP.n_flev + 1, n_bp = SP.n_blev + 1;
        int* const s_lp = ip; int* const s_up = s_lp + n_lp; int* const s_fp = s_up + n_up; int* const s_bp = s_fp + n_fp;
        unsigned short* const s_lc = reinterpret_cast<unsigned short*>(s_bp + n_bp);
        unsigned short* const s_uc = s_lc + SP.nnzL;
        unsigned short* const s_fr = s_uc + SP.nnzU;
        unsigned short* const s_br = s_fr + SP.n1;
        unsigned short* const s_io = s_br + SP.n1;
        for (int e = threadIdx.x; e < SP.nx; e += TEAM_THREADS_) s_io[e] = SP.iord16[e];
        bdf.SP.iord16 = s_io;
        for (int e = threadIdx.x; e < n_lp; e += TEAM_THREADS_) s_lp[e] = SP.Lr_ptr[e];
        for (int e = threadIdx.x; e < n_up; e += TEAM_THREADS_) s_up[e] = SP.Ur_ptr[e];
        for (int e = threadIdx.x; e < n_fp; e += TEAM_THREADS_) s_fp[e] = SP.flev_ptr[e];
        for (int e = threadIdx.x; e < n_bp; e += TEAM_THREADS_) s_bp[e] = SP.blev_ptr[e];
        for (int e = threadIdx.x; e < SP.nnzL; e += TEAM_THREADS_) s_lc[e] = SP.Lr_col[e];
        for (int e = threadIdx.x; e < SP.nnzU; e += TEAM_THREADS_) s_uc[e] = SP.Ur_col[e];
        for (int e = threadIdx.x; e < SP.n1; e += TEAM_THREADS_) { s_fr[e] = SP.flev_rows[e]; s_br[e] = SP.blev_rows[e]; }
        bdf.SP.Lr_ptr = s_lp; bdf.SP.Ur_ptr = s_up; bdf.SP.flev_ptr = s_fp; bdf.SP.blev_ptr = s_bp;
        bdf.SP.Lr_col = s_lc; bdf.SP.Ur_col = s_uc; bdf.SP.flev_rows = s_fr; bdf.SP.blev_rows = s_br;
        __syncthreads();
    }
    const int tid = threadIdx.x;
    const int np = T.np;
    for (;;) {
        if (tid == 0) s_pos = atomicAdd(a.counter, 1);
        __syncthreads();
        const int pos = s_pos;
        __syncthreads();
        if (pos >= a.n) break;
        const int r = a.order ? a.order[pos] : pos;
        const int c = a.idx[r];
        const size_t cp = (size_t)c * np;
        // A5: optimisation -> simulation parameters (integer gather from the host-computed table)
        for (int i = tid; i < np; i += TEAM_THREADS) {
            const int j = a.mapping[cp + i];
            double pu;
            if (j >= 0) pu = a.theta_tab[((size_t)j * 3 + a.pscale_sim[cp + i]) * 2 + 1];
            else if (a.ov_unscaled) pu = a.ov_unscaled[cp + i];
            else pu = NAN;
            bdf.w.p[i] = pu;
            bdf.w.plist[i] = a.plist[cp + i];
        }
        for (int i = tid; i < T.nk; i += TEAM_THREADS) bdf.w.k[i] = a.k[(size_t)c * T.nk + i];
        __syncthreads();
        ConditionIO io;
        const int64_t t0off = a.t_offsets[c];
        io.nt = (int)(a.t_offsets[c + 1] - t0off);
        io.t = a.t + t0off;
        io.my = a.my + t0off * T.ny;
        io.sd = a.sd + t0off * T.ny;
        io.ns = a.want_grad ? a.nplist[c] : 0;
        io.y_out = a.y_out ? a.y_out + t0off * T.ny : nullptr;
        io.sigmay_out = a.sigmay_out ? a.sigmay_out + t0off * T.ny : nullptr;
        // A8: runAndLogSimulation retries at relaxed tolerances
        int trial = 1, status;
        double relax = 1.0;
        for (;;) {
            // pre-equilibration with the pre-equilibration condition's constants (ConditionContext,
            // deps/AMICI/src/edata.cpp:511-524), then the main simulation from the steady state with the
            // simulation condition's; a retry repeats both (multiConditionProblem.cpp:369-465)
            const bool pre = SS && a.preeq_flag && a.preeq_flag[c] != 0;
            const int* reinit = nullptr;
            int nre = 0;
            status = SIM_SUCCESS;
            if (pre) {
                __syncthreads();
                for (int i = tid; i < T.nk; i += TEAM_THREADS) bdf.w.k[i] = a.k_preeq[(size_t)c * T.nk + i];
                __syncthreads();
                status = bdf.preequilibrate(io, a.opt.rtol * relax, a.opt.atol * relax);
                __syncthreads();
                for (int i = tid; i < T.nk; i += TEAM_THREADS) bdf.w.k[i] = a.k[(size_t)c * T.nk + i];
                __syncthreads();
                if (a.reinit_off) {
                    reinit = a.reinit_idx + a.reinit_off[c];
                    nre = (int)(a.reinit_off[c + 1] - a.reinit_off[c]);
                }
            }
            if (status == SIM_SUCCESS) status = bdf.simulate(io, a.opt.rtol * relax, a.opt.atol * relax, pre, reinit, nre);
            if (status == SIM_SUCCESS || trial >= a.num_trials) break;
            relax *= a.relaxation_factor;
            ++trial;
        }
        // outputs: A7 result package + A6 chain rule to the optimiser scale
        if (tid == 0) {
            a.llh[r] = bdf.llh();
            a.status[r] = status;
            bdf.export_stats(a.stats + (size_t)r * NSTATS, trial);
        }
        if (a.want_grad) {
            for (int i = tid; i < np; i += TEAM_THREADS) {
                a.gcontrib[(size_t)r * np + i] = 0.0;
                a.sllh[(size_t)r * np + i] = 0.0;
            }
            __syncthreads();
            for (int col = 1 + tid; col <= io.ns; col += TEAM_THREADS) {
                const int ip = bdf.w.plist[col - 1];
                const int ssim = a.pscale_sim[cp + ip];
                const int jopt = a.mapping[cp + ip];
                const double pu = bdf.w.p[ip];
                double s = bdf.w.sllh[col];  // ReturnData::applyChainRuleFactorToSimulationResults
                if (ssim == 2) s *= pu * PB_LN10;
                else if (ssim == 1) s *= pu;
                a.sllh[(size_t)r * np + (col - 1)] = s;
                const int sopt = a.pscale_opt[jopt];  // applyChainRule, coefficient -1
                double gr = s;
                if (ssim != sopt) {
                    if (ssim == 2) gr /= pu * PB_LN10;
                    else if (ssim == 1) gr /= pu;
                    if (sopt == 2) gr *= pu * PB_LN10;
                    else if (sopt == 1) gr *= pu;
                }
                a.gcontrib[(size_t)r * np + ip] = -1.0 * gr;
            }
        }
        __syncthreads();
    }
}

/// Device copy of a network's tables + the teams' HBM workspace (owned by the handle).
struct TeamState {
    NetTables dev{};
    SparseLU sp{};          // device view of the symbolic factorisation
    SparseLUHost sp_host;   // host copy (counts for the roofline report)
    bool sparse_default = false;
    std::vector<void*> allocs;
    double* workspace = nullptr;
    size_t ws_stride = 0;
    int grid_cap = 0;
    size_t smem = 0;      // dynamic shared memory of a block (bytes)
    int sp_tab_off = -1;  // where the sparse solve tables start in it (doubles), -1: they stay in global memory
};

template <class V>
static cudaError_t team_upload(TeamState* st, const V* src, size_t n, const V** dst) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (n > 0 ? n : 1) * sizeof(V));
    if (e != cudaSuccess) return e;
    st->allocs.push_back(p);
    if (n > 0) e = cudaMemcpy(p, src, n * sizeof(V), cudaMemcpyHostToDevice);
    *dst = static_cast<const V*>(p);
    return e;
}

inline void team_destroy_state(void* s) {
    auto* st = static_cast<TeamState*>(s);
    if (!st) return;
    for (void* p : st->allocs) cudaFree(p);
    if (st->workspace) cudaFree(st->workspace);
    delete st;
}

inline void* team_create_state(const NetTables& H, int sm_count, std::string* err) {
    auto* st = new TeamState;
    auto fail = [&](cudaError_t e) -> void* {
        if (err) *err = cudaGetErrorString(e);
        team_destroy_state(st);
        return nullptr;
    };
    if (H.ny > TEAM_MAXNY || H.jell_w > TEAM_JW_MAX) {
        if (err) *err = "too many observables / too wide a Jacobian row for the team kernel";
        delete st;
        return nullptr;
    }
    st->dev = H;
    cudaError_t e;
#define UP(field, n) if ((e = team_upload(st, H.field, (size_t)(n), &st->dev.field)) != cudaSuccess) return fail(e)
    UP(r_a, H.nr); UP(r_b, H.nr); UP(kappa, H.nr); UP(pidx, H.nr);
    UP(s_ptr, H.nx + 1); UP(s_reac, H.n_s); UP(s_coef, H.n_s);
    UP(j_ptr, H.nx + 1); UP(j_col, H.nnz); UP(jt_ptr, H.nnz + 1); UP(jt_reac, H.n_jt); UP(jt_other, H.n_jt); UP(jt_coef, H.n_jt);
    UP(jell_col, (size_t)H.jell_w * H.nx); UP(jell_pos, H.nnz); UP(csc_ptr, H.nx + 1); UP(csc_row, H.nnz); UP(csc_from_csr, H.nnz);
    UP(x0_base, H.nx); UP(x0_kidx, H.nx); UP(o_ptr, H.ny + 1); UP(o_idx, H.n_o); UP(o_w, H.n_o); UP(sigma0, H.ny);
#undef UP
    // symbolic factorisation of the Newton matrix's pattern, once per model (the klu_analyze role)
    {
        st->sp_host = SparseLUHost::build_for_team(H.nx, H.j_ptr, H.j_col, H.jell_pos, TEAM_THREADS);
        const SparseLUHost& S = st->sp_host;
        st->sp = S.view();
#define UPV(field, vec) if ((e = team_upload(st, S.vec.data(), S.vec.size(), &st->sp.field)) != cudaSuccess) return fail(e)
        UPV(ord, ord); UPV(iord, iord);
        UPV(tlevel_ptr, tlevel_ptr); UPV(t_dst, t_dst); UPV(t_src, t_src); UPV(t_piv, t_piv); UPV(t_pptr, t_pptr);
        UPV(p_l, p_l); UPV(p_u, p_u);
        UPV(s_dst, s_dst); UPV(s_src, s_src); UPV(s_diag, s_diag); UPV(s_pptr, s_pptr);
        UPV(Lr_ptr, Lr_ptr); UPV(Lr_col, Lr_col16); UPV(Ur_ptr, Ur_ptr); UPV(Ur_col, Ur_col16);
        UPV(flev_ptr, flev_ptr); UPV(flev_rows, flev_rows16); UPV(blev_ptr, blev_ptr); UPV(blev_rows, blev_rows16);
        UPV(iord16, iord16); UPV(flev_long, flev_long); UPV(blev_long, blev_long); UPV(tlev_long, tlev_long);
#undef UPV
        // automatic choice (PB200_OPT_LINSOL = 0): sparse where it removes a real part of the dense work --
        // the dense block is at most 3/4 of the matrix -- dense for small or nearly full patterns
        st->sparse_default = H.nx >= 128 && 4 * S.m <= 3 * H.nx;
    }
    // Dynamic shared memory: all a block may have (one block per SM); the sparse solve tables sit at its
    // end when what is left still holds the rows path's ring (dense dimension <= 512) or a panel of at
    // least 8 columns of the generic path.
    {
        const size_t total = (227u * 1024u - 1024u) / sizeof(double);  // (static shared memory + margin)
        const size_t tab = ((size_t)st->sp.solve_table_bytes() + 15) / 16 * 2;
        // the dense path may be asked to take the whole matrix (fallback / PB200_OPT_LINSOL = 1): rows path (ring)
        // up to 512 rows, otherwise the generic path with a panel of at least 8 columns; the Schur block of the
        // sparse solver takes the rows path whenever it has <= 512 rows
        size_t need = H.nx <= TEAM_THREADS ? (size_t)TEAM_SM_ROWS_END : (size_t)TEAM_SM_PANEL + (size_t)H.nx * 8;
        if (st->sp.m <= TEAM_THREADS && need < (size_t)TEAM_SM_ROWS_END) need = (size_t)TEAM_SM_ROWS_END;
        st->smem = total * sizeof(double);
        if (total < need) {
            if (err) *err = "model too large for the team kernel's shared-memory staging";
            team_destroy_state(st);
            return nullptr;
        }
        st->sp_tab_off = (total >= need + tab) ? (int)(total - tab) : -1;
    }
    st->ws_stride = TeamWork::hbm_doubles(H, st->sp.nV);
    st->grid_cap = sm_count;
    e = cudaMalloc((void**)&st->workspace, st->ws_stride * sizeof(double) * (size_t)st->grid_cap);
    if (e != cudaSuccess) return fail(e);
    return st;
}

int team_launch_kernel_plain(const KernelArgs& a, const TeamState& st, int grid, int use_sparse, cudaStream_t stream);
int team_launch_kernel_steady(const KernelArgs& a, const TeamState& st, int grid, int use_sparse, cudaStream_t stream);

/// body of the two functions above (SS = false / true)
template <bool SS>
int team_launch_kernel(const KernelArgs& a, const TeamState& st, int grid, int use_sparse, cudaStream_t stream) {
    auto kern = simulate_team_kernel<TEAM_THREADS, SS>;
    // (the attribute is per device and per function: set at every launch, to this model's size)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st.smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, TEAM_THREADS, st.smem, stream>>>(a, st.dev, st.sp, st.workspace, st.ws_stride,
                                                  (int)(st.smem / sizeof(double)), st.sp_tab_off, use_sparse);
    return (int)cudaGetLastError();
}

inline int team_launch(const KernelArgs& a, void* state, cudaStream_t stream, int* lanes_used, int* grid_used,
                       int* block_used) {
    auto* st = static_cast<TeamState*>(state);
    if (!st) return (int)cudaErrorInvalidValue;
    int grid = st->grid_cap;  // persistent: one team per SM, work queue inside
    if (grid > a.n) grid = a.n;
    if (grid < 1) grid = 1;
    const int use_sparse = a.linsol == 2 ? 1 : (a.linsol == 1 ? 0 : (st->sparse_default ? 1 : 0));
    // the kernel is generic over the network tables: it is instantiated ONCE per variant, in its own translation
    // unit (team_kernel_plain.cu / team_kernel_steady.cu), not once per model
    const int e = a.steady ? team_launch_kernel_steady(a, *st, grid, use_sparse, stream)
                           : team_launch_kernel_plain(a, *st, grid, use_sparse, stream);
    if (e != 0) return e;
    if (lanes_used) *lanes_used = TEAM_THREADS;
    if (grid_used) *grid_used = grid;
    if (block_used) *block_used = TEAM_THREADS;
    return 0;
}

/// description of the linear solver for reports (pb200_linsol_info): see parpe_b200/problem.py for the names
inline int team_linsol_info(void* state, int64_t* v, int n) {
    auto* st = static_cast<TeamState*>(state);
    if (!st) return 0;
    const SparseLUHost& S = st->sp_host;
    const int64_t vals[14] = {st->sparse_default ? 2 : 1, S.nx, S.n1, S.m, S.nnzJ, S.nnzL, S.nnzU,
                              S.n_pairs_sparse + S.n_pairs_schur, (int64_t)S.flev_ptr.size() - 1,
                              (int64_t)S.blev_ptr.size() - 1, S.factor_flops(), S.solve_flops(),
                              8 * ((int64_t)S.nV + (int64_t)S.m * S.m), 8 * ((int64_t)S.nV + (int64_t)S.m * S.m)};
    int k = 0;
    for (; k < 14 && k < n; ++k) v[k] = vals[k];
    return k;
}

#define PB200_REGISTER_NET(NAME)                                                                              \
    namespace {                                                                                               \
    void* create_##NAME(int sm_count, std::string* err) {                                                     \
        return pb200::team_create_state(pb200::nets::NAME::tables(), sm_count, err);                          \
    }                                                                                                         \
    int launch_##NAME(const pb200::KernelArgs& a, int, int, int, cudaStream_t stream, int* lanes_used,        \
                      int* grid_used, int* block_used, void* state) {                                         \
        return pb200::team_launch(a, state, stream, lanes_used, grid_used, block_used);                       \
    }                                                                                                         \
    struct RegNet_##NAME {                                                                                    \
        RegNet_##NAME() {                                                                                     \
            const pb200::NetTables T = pb200::nets::NAME::tables();                                           \
            pb200::ModelOps o{};                                                                              \
            o.name = T.name;                                                                                  \
            o.nx = T.nx; o.np = T.np; o.nk = T.nk; o.ny = T.ny;                                               \
            o.default_lanes = pb200::TEAM_THREADS;                                                            \
            o.flops_xdot = T.flops_xdot; o.flops_J = T.flops_J; o.flops_dxdotdp = T.flops_dxdotdp;            \
            o.nnzJ = T.nnz;                                                                                   \
            o.launch = &launch_##NAME;                                                                        \
            o.create_state = &create_##NAME;                                                                  \
            o.destroy_state = &pb200::team_destroy_state;                                                     \
            o.linsol_info = &pb200::team_linsol_info;                                                         \
            o.supports_steady = 1;                                                                            \
            pb200::register_model(o);                                                                         \
        }                                                                                                     \
    } reg_net_instance_##NAME;                                                                                \
    }

}  // namespace pb200
#endif
