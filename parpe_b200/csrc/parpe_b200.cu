// parpe_b200.cu -- C ABI (include/parpe_b200.h) over the CUDA kernels.
//
// Host side of the per-GPU condition-batch scheduler that replaces
// AmiciSimulationRunner + LoadBalancerMaster/Worker:
// tables are resident in HBM, one evaluation = upload of the parameter table
// (48*n_opt bytes) and, if it changed, the index list; one persistent kernel with an
// in-kernel work queue; one reduction kernel; download of 2+n_opt doubles.
// There is no CPU fallback: without a CUDA device every compute entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/parpe_b200.h"
#include "h5wmin.h"
#include "hier_kernels.cuh"
#include "kernels.cuh"
#include "nccl_dyn.h"

namespace pb200 {

static std::vector<ModelOps>& registry() {
    static std::vector<ModelOps> r;
    return r;
}
void register_model(const ModelOps& ops) { registry().push_back(ops); }
const ModelOps* find_model(const char* name) {
    for (auto& m : registry())
        if (!std::strcmp(m.name, name)) return &m;
    return nullptr;
}
int model_count() { return (int)registry().size(); }
const ModelOps* model_at(int i) { return (i >= 0 && i < model_count()) ? &registry()[i] : nullptr; }

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_message(int code, const std::string& msg) { return fail(code, msg); }
#define PB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(PB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// amici::getUnscaledParameter / getScaledParameter (deps/AMICI/src/misc.cpp:26-63), evaluated
// on the host with the same libm calls as the reference so the values are bit-identical
static double unscale(double v, int s) {
    switch (s) {
    case PB200_SCALE_LOG10: return std::pow(10, v);
    case PB200_SCALE_LN: return std::exp(v);
    default: return v;
    }
}
static double scale(double v, int s) {
    switch (s) {
    case PB200_SCALE_LOG10: return std::log10(v);
    case PB200_SCALE_LN: return std::log(v);
    default: return v;
    }
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const T* src, size_t count) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMallocHost((void**)&p, count * sizeof(T));
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    ~PinnedBuf() { release(); }
};

}  // namespace pb200

using namespace pb200;

struct pb200_handle {
    const ModelOps* ops = nullptr;
    int device = 0, sm_count = 0;
    int C = 0, n_opt = 0, np = 0, nk = 0, ny = 0;
    int64_t nt_total = 0;
    int lanes = 0;
    int rank = 0, world = 1;
    std::vector<int> pscale_opt;
    std::vector<int64_t> t_offsets;
    // host copies for the result log (simulation parameters per condition: A5 on the host)
    std::vector<int> mapping_host, pscale_sim_host;
    std::vector<double> overrides_host;
    double last_seconds = 0.0;
    // resident tables
    DevBuf<int64_t> d_t_offsets;
    DevBuf<double> d_t, d_my, d_sd, d_k, d_ov_scaled, d_ov_unscaled;
    DevBuf<int> d_mapping, d_pscale_sim, d_pscale_opt, d_plist, d_nplist, d_inv_j, d_inv_i;
    // per-evaluation buffers
    int cap = 0;
    DevBuf<double> d_llh, d_sllh, d_gcontrib, d_theta_tab, d_out, d_yout, d_sout;
    DevBuf<int> d_status, d_idx, d_counter, d_order;
    // cost-aware processing order: positions sorted by the previous evaluation's step
    // counts, valid while the index list is unchanged
    std::vector<int> order_host;
    std::vector<int64_t> stats_host;
    bool order_valid = false;
    int reorder = 0;  // 0 off, 1 = descending previous nst
    DevBuf<int64_t> d_stats;
    PinnedBuf<double> h_theta_tab, h_out;
    PinnedBuf<int> h_idx;
    std::vector<int> idx_cache;  // index list currently resident in d_idx
    int last_n = 0;
    int last_want_grad = 0;
    int last_lanes = 0, last_grid = 0, last_block = 0;
    KernelArgs args{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_staged = nullptr;
    bool staged_pending = false;
    // hierarchical
    HierDevice* hier = nullptr;
    // collectives of partitioned evaluations: a host hook (pb200_set_allreduce: the caller's transport,
    // e.g. MPI_Allreduce) or an in-library NCCL communicator (pb200_nccl_init)
    pb200_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    NcclApi::Comm nccl = nullptr;
    DevBuf<double> d_coll;  // staging buffer of the small NCCL reductions of the hierarchical passes
    // tuning options (pb200_set_option; never read from the environment)
    int opt_variant = 2, opt_ctas_per_sm = 0, opt_sync_period = 1, opt_stage_sync = 0, opt_partial_sums = 0;
    int opt_linsol = 0, opt_order_refresh = 16, opt_fuse_off = 0;
    bool steady_sens_newton = false;  // pb200_problem_desc::steady_sens_newton
    bool simultaneous = false;        // pb200_problem_desc::ism == 1: needs the full-feature kernel instantiation
    int evals_since_order = 0;
    int64_t last_h2d = 0, last_d2h = 0;
    // large-model path: device tables + workspace (ModelOps::create_state)
    void* model_state = nullptr;
    // steady states: t = inf output times in the table / pre-equilibration tables (pb200_set_preequilibration)
    bool has_inf = false, has_preeq = false;
    DevBuf<double> d_k_preeq;
    DevBuf<int> d_preeq_flag;
    DevBuf<unsigned> d_reinit_mask;
    DevBuf<int64_t> d_reinit_off;
    DevBuf<int> d_reinit_idx;
};

static int ensure_capacity(pb200_handle* h, int n) {
    if (n <= h->cap) return 0;
    int cap = n;
    PB_CUDA(h->d_llh.alloc(cap));
    PB_CUDA(h->d_sllh.alloc((size_t)cap * h->np));
    PB_CUDA(h->d_gcontrib.alloc((size_t)cap * h->np));
    PB_CUDA(h->d_status.alloc(cap));
    PB_CUDA(h->d_stats.alloc((size_t)cap * NSTATS));
    PB_CUDA(h->d_idx.alloc(cap));
    PB_CUDA(h->d_order.alloc(cap));
    PB_CUDA(h->h_idx.alloc(cap));
    h->order_valid = false;
    h->idx_cache.clear();
    h->cap = cap;
    return 0;
}

// stage theta table + index list and enqueue simulate + reduce on `stream`
static int enqueue_evaluation(pb200_handle* h, const double* theta, const int* idx, int n_idx, int want_grad,
                              double* d_out, double* d_yout, double* d_sout, cudaStream_t stream) {
    PB_CUDA(cudaSetDevice(h->device));
    int lo, hi;
    pb200_partition_bounds(n_idx, h->rank, h->world, &lo, &hi);
    const int n = hi - lo;
    // the WHOLE list is validated on every rank (not only this rank's share): with a collective inside the
    // evaluation a rank that bailed out alone would leave the others waiting in the all-reduce
    for (int r = 0; r < n_idx; ++r)
        if (idx[r] < 0 || idx[r] >= h->C) return fail(PB200_ERR_ARG, "data index out of range");
    if (int rc = ensure_capacity(h, n > 0 ? n : 1)) return rc;
    if (h->staged_pending) {  // pinned staging buffers may still be in flight
        PB_CUDA(cudaEventSynchronize(h->ev_staged));
        h->staged_pending = false;
    }
    // A5: per optimisation parameter and simulation scale: value in simulation scale and its
    // unscaled value, exactly as mapAndSetOptimizationToSimulationVariables + AMICI compute them
    for (int j = 0; j < h->n_opt; ++j) {
        const double lin = unscale(theta[j], h->pscale_opt[j]);
        for (int s = 0; s < 3; ++s) {
            const double sc = scale(lin, s);
            h->h_theta_tab.p[((size_t)j * 3 + s) * 2 + 0] = sc;
            h->h_theta_tab.p[((size_t)j * 3 + s) * 2 + 1] = unscale(sc, s);
        }
    }
    PB_CUDA(cudaMemcpyAsync(h->d_theta_tab.p, h->h_theta_tab.p, sizeof(double) * 6 * h->n_opt,
                            cudaMemcpyHostToDevice, stream));
    const bool same_idx = (int)h->idx_cache.size() == n && (n == 0 || !std::memcmp(h->idx_cache.data(), idx + lo, sizeof(int) * n));
    if (!same_idx) {
        h->order_valid = false;
        std::memcpy(h->h_idx.p, idx + lo, sizeof(int) * n);
        PB_CUDA(cudaMemcpyAsync(h->d_idx.p, h->h_idx.p, sizeof(int) * n, cudaMemcpyHostToDevice, stream));
        h->idx_cache.assign(idx + lo, idx + hi);
    }
    PB_CUDA(cudaMemsetAsync(h->d_counter.p, 0, sizeof(int), stream));
    PB_CUDA(cudaEventRecord(h->ev_staged, stream));
    h->staged_pending = true;

    KernelArgs a = h->args;
    a.theta_tab = h->d_theta_tab.p;
    a.idx = h->d_idx.p;
    // (any permutation is valid; the ranking by step counts holds with and without sensitivities)
    a.order = (h->reorder && h->order_valid) ? h->d_order.p : nullptr;
    a.n = n;
    a.want_grad = want_grad;
    a.counter = h->d_counter.p;
    a.llh = h->d_llh.p;
    a.sllh = h->d_sllh.p;
    a.gcontrib = h->d_gcontrib.p;
    a.status = h->d_status.p;
    a.stats = h->d_stats.p;
    a.y_out = d_yout;
    a.sigmay_out = d_sout;
    a.sync_period = h->opt_sync_period > 0 ? h->opt_sync_period : 1;
    a.stage_sync = h->opt_stage_sync;
    a.ctas_per_sm_cap = h->opt_ctas_per_sm;
    a.linsol = h->opt_linsol;
    a.fuse_off = h->opt_fuse_off;
    PB_CUDA(cudaEventRecord(h->ev_start, stream));
    if (n > 0) {
        int e = h->ops->launch(a, h->lanes, h->opt_variant, h->sm_count, stream, &h->last_lanes, &h->last_grid,
                               &h->last_block, h->model_state);
        if (e != 0) return fail(PB200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)e));
    }
    launch_reduce(a, d_out, stream);
    PB_CUDA(cudaGetLastError());
    // A6 over all ranks: the ONE all-reduce of (nllh, #errors, gradient) that replaces the master's
    // summation of the workers' results (multiConditionProblem.cpp:738-748,775-824)
    if (h->nccl) {
        NcclApi& nc = NcclApi::get();
        const int rc = nc.AllReduce(d_out, d_out, (size_t)(2 + h->n_opt), NcclApi::Float64, NcclApi::Sum, h->nccl, stream);
        if (rc != NcclApi::Success) return fail(PB200_ERR_CUDA, std::string("ncclAllReduce: ") + nc.GetErrorString(rc));
    }
    PB_CUDA(cudaEventRecord(h->ev_stop, stream));
    h->last_h2d = (int64_t)sizeof(double) * 6 * h->n_opt + (same_idx ? 0 : (int64_t)sizeof(int) * n);
    h->last_d2h = 0;
    h->last_n = n;
    h->last_want_grad = want_grad;
    return 0;
}

extern "C" {

int pb200_abi_version(void) { return PB200_ABI_VERSION; }

int pb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pb200_model_count(void) { return model_count(); }
const char* pb200_model_name(int i) {
    const ModelOps* m = model_at(i);
    return m ? m->name : nullptr;
}
int pb200_model_info(const char* model_name, int* nx, int* np, int* nk, int* ny) {
    const ModelOps* m = model_name ? find_model(model_name) : nullptr;
    if (!m) return fail(PB200_ERR_MODEL, "unknown model");
    if (nx) *nx = m->nx;
    if (np) *np = m->np;
    if (nk) *nk = m->nk;
    if (ny) *ny = m->ny;
    return 0;
}
const char* pb200_last_error(void) { return g_err.c_str(); }

void pb200_partition_bounds(int n, int rank, int world, int* lo, int* hi) {
    if (world < 1) world = 1;
    *lo = (int)(((int64_t)rank * n) / world);
    *hi = (int)(((int64_t)(rank + 1) * n) / world);
}

int pb200_create(const pb200_problem_desc* d, int device, pb200_handle** out) {
    if (!d || !out) return fail(PB200_ERR_ARG, "null argument");
    *out = nullptr;
    if (d->abi_version != PB200_ABI_VERSION) return fail(PB200_ERR_ARG, "ABI version mismatch");
    const ModelOps* ops = d->model_name ? find_model(d->model_name) : nullptr;
    if (!ops) return fail(PB200_ERR_MODEL, std::string("unknown model ") + (d->model_name ? d->model_name : "(null)"));
    if (ops->np != d->np || ops->nk != d->nk || ops->ny != d->ny)
        return fail(PB200_ERR_MODEL, "np/nk/ny do not match the model");
    if (d->n_conditions < 0 || d->n_opt < 0 || !d->t_offsets) return fail(PB200_ERR_ARG, "bad dimensions");
    if (pb200_device_count() <= 0) return fail(PB200_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    PB_CUDA(cudaSetDevice(device));
    const int C = d->n_conditions;
    bool has_inf = false;
    for (int c = 0; c < C; ++c) {
        if (d->t_offsets[c + 1] < d->t_offsets[c]) return fail(PB200_ERR_ARG, "t_offsets not monotone");
        // Output times: non-decreasing (amici::ExpData requires sorted time points), finite or +inf.
        // t = +inf asks for post-equilibration (steadystateproblem.cpp): the data point is evaluated at the
        // steady state the simulation is carried on to.  Models whose kernels do not handle it refuse the
        // table rather than silently dropping the point from the objective.
        for (int64_t it = d->t_offsets[c]; it < d->t_offsets[c + 1]; ++it) {
            if (std::isnan(d->t[it]) || d->t[it] == -std::numeric_limits<double>::infinity())
                return fail(PB200_ERR_ARG, "output time is NaN or -inf");
            if (std::isinf(d->t[it])) {
                if (!ops->supports_steady)
                    return fail(PB200_ERR_ARG, "non-finite output time (t = inf / post-equilibration is not supported for this model)");
                has_inf = true;
            }
            if (it > d->t_offsets[c] && d->t[it] < d->t[it - 1])
                return fail(PB200_ERR_ARG, "output times of a condition must be non-decreasing");
        }
        for (int i = 0; i < d->np; ++i) {
            const int j = d->mapping[(size_t)c * d->np + i];
            if (j >= d->n_opt) return fail(PB200_ERR_ARG, "mapping entry out of range");
            const int s = d->pscale_sim[(size_t)c * d->np + i];
            if (s < 0 || s > 2) return fail(PB200_ERR_ARG, "invalid simulation parameter scale");
        }
    }
    for (int j = 0; j < d->n_opt; ++j)
        if (d->pscale_opt[j] < 0 || d->pscale_opt[j] > 2) return fail(PB200_ERR_ARG, "invalid optimisation parameter scale");

    auto* h = new pb200_handle;
    std::unique_ptr<pb200_handle, void (*)(pb200_handle*)> guard(h, pb200_destroy);
    h->ops = ops;
    h->device = device;
    cudaDeviceProp prop;
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->C = C;
    h->n_opt = d->n_opt;
    h->np = d->np;
    h->nk = d->nk;
    h->ny = d->ny;
    h->lanes = d->lanes_per_condition;
    h->reorder = 1;
    h->has_inf = has_inf;
    h->pscale_opt.assign(d->pscale_opt, d->pscale_opt + d->n_opt);
    h->t_offsets.assign(d->t_offsets, d->t_offsets + C + 1);
    h->nt_total = C > 0 ? d->t_offsets[C] : 0;
    h->mapping_host.assign(d->mapping, d->mapping + (size_t)C * d->np);
    h->pscale_sim_host.assign(d->pscale_sim, d->pscale_sim + (size_t)C * d->np);
    if (d->overrides) h->overrides_host.assign(d->overrides, d->overrides + (size_t)C * d->np);

    // A5 getSensitivityParameterList (multiConditionDataProvider.cpp:323-339), once per condition
    std::vector<int> plist((size_t)C * d->np, 0), nplist(C, 0);
    for (int c = 0; c < C; ++c) {
        int k = 0;
        for (int i = 0; i < d->np; ++i)
            if (d->mapping[(size_t)c * d->np + i] >= 0) plist[(size_t)c * d->np + k++] = i;
        nplist[c] = k;
    }
    PB_CUDA(h->d_t_offsets.upload(d->t_offsets, C + 1));
    PB_CUDA(h->d_t.upload(d->t, h->nt_total));
    PB_CUDA(h->d_my.upload(d->y, h->nt_total * d->ny));
    PB_CUDA(h->d_sd.upload(d->sigma, h->nt_total * d->ny));
    PB_CUDA(h->d_k.upload(d->k, (size_t)C * d->nk));
    PB_CUDA(h->d_mapping.upload(d->mapping, (size_t)C * d->np));
    PB_CUDA(h->d_pscale_sim.upload(d->pscale_sim, (size_t)C * d->np));
    PB_CUDA(h->d_pscale_opt.upload(d->pscale_opt, d->n_opt));
    PB_CUDA(h->d_plist.upload(plist.data(), plist.size()));
    PB_CUDA(h->d_nplist.upload(nplist.data(), nplist.size()));
    {
        // per condition: (optimisation index, model parameter) pairs sorted by optimisation index -- the
        // gradient reduction looks a gradient entry up by binary search instead of scanning the mapping
        std::vector<int> inv_j((size_t)C * d->np, std::numeric_limits<int>::max()), inv_i((size_t)C * d->np, 0);
        std::vector<std::pair<int, int>> tmp;
        for (int c = 0; c < C; ++c) {
            tmp.clear();
            for (int i = 0; i < d->np; ++i) {
                const int j = d->mapping[(size_t)c * d->np + i];
                if (j >= 0) tmp.emplace_back(j, i);
            }
            std::sort(tmp.begin(), tmp.end());
            for (size_t e = 0; e < tmp.size(); ++e) {
                inv_j[(size_t)c * d->np + e] = tmp[e].first;
                inv_i[(size_t)c * d->np + e] = tmp[e].second;
            }
        }
        PB_CUDA(h->d_inv_j.upload(inv_j.data(), inv_j.size()));
        PB_CUDA(h->d_inv_i.upload(inv_i.data(), inv_i.size()));
    }
    if (d->overrides) {
        std::vector<double> ovu((size_t)C * d->np);
        for (size_t i = 0; i < ovu.size(); ++i) ovu[i] = unscale(d->overrides[i], d->pscale_sim[i]);
        PB_CUDA(h->d_ov_scaled.upload(d->overrides, ovu.size()));
        PB_CUDA(h->d_ov_unscaled.upload(ovu.data(), ovu.size()));
    }
    PB_CUDA(h->d_theta_tab.alloc((size_t)6 * (d->n_opt > 0 ? d->n_opt : 1)));
    PB_CUDA(h->h_theta_tab.alloc((size_t)6 * (d->n_opt > 0 ? d->n_opt : 1)));
    PB_CUDA(h->d_out.alloc(2 + d->n_opt));
    PB_CUDA(h->h_out.alloc(2 + d->n_opt));
    PB_CUDA(h->d_counter.alloc(1));
    PB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreate(&h->ev_start));
    PB_CUDA(cudaEventCreate(&h->ev_stop));
    PB_CUDA(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));

    KernelArgs& a = h->args;
    a.n_conditions = C;
    a.n_opt = d->n_opt;
    a.np = d->np;
    a.nk = d->nk;
    a.ny = d->ny;
    a.t_offsets = h->d_t_offsets.p;
    a.t = h->d_t.p;
    a.my = h->d_my.p;
    a.sd = h->d_sd.p;
    a.k = h->d_k.p;
    a.mapping = h->d_mapping.p;
    a.pscale_sim = h->d_pscale_sim.p;
    a.pscale_opt = h->d_pscale_opt.p;
    a.ov_scaled = h->d_ov_scaled.p;
    a.ov_unscaled = h->d_ov_unscaled.p;
    a.plist = h->d_plist.p;
    a.nplist = h->d_nplist.p;
    a.inv_j = h->d_inv_j.p;
    a.inv_i = h->d_inv_i.p;
    a.k_preeq = nullptr;
    a.preeq_flag = nullptr;
    a.reinit_mask = nullptr;
    a.reinit_off = nullptr;
    a.reinit_idx = nullptr;
    a.steady = has_inf ? 1 : 0;
    a.opt.t0 = d->t0;
    a.opt.rtol = d->rtol;
    a.opt.atol = d->atol;
    a.opt.maxsteps = d->maxsteps;
    a.opt.stldet = d->stldet;
    a.opt.errconS = d->sensi_errcon;
    a.opt.newton_maxsteps = d->newton_maxsteps > 0 ? d->newton_maxsteps : 0;
    if (d->newton_only != 0) {
        // steadyStateComputationMode newtonOnly: with newton_maxsteps = 0 the reference runs neither Newton's method
        // nor the simulation (steadystateproblem.cpp:133-146) and every simulation fails -- a configuration error
        if (d->newton_only != 1) return fail(PB200_ERR_ARG, "newton_only must be 0 or 1");
        if (a.opt.newton_maxsteps <= 0) return fail(PB200_ERR_ARG, "newton_only needs newton_maxsteps > 0");
        a.opt.newton_maxsteps = -a.opt.newton_maxsteps;  // (SolverOptions: negative = Newton's method only)
    }
    if (d->linsol < 0 || d->linsol > 2) return fail(PB200_ERR_ARG, "linsol: 0 automatic, 1 dense, 2 sparse");
    if (d->steady_sens_newton != 0) {
        if (d->steady_sens_newton != 1) return fail(PB200_ERR_ARG, "steady_sens_newton must be 0 or 1");
        if (!a.opt.errconS) return fail(PB200_ERR_ARG, "steady_sens_newton needs sensi_errcon = 1 (AMICI forces it)");
        // (the mode matters only where a steady state is computed: refused for a model without Newton's linear
        // solve once the problem has t = inf output times -- here -- or gets a pre-equilibration -- there)
        h->steady_sens_newton = true;
        if (ops->supports_newton) a.opt.errconS = 3;  // (SolverOptions::errconS bit 1)
        else if (has_inf)
            return fail(PB200_ERR_ARG, "steady_sens_newton with t = inf output times: the linear solve for the sensitivities "
                                       "of a steady state is not supported for this model");
    }
    h->opt_linsol = d->linsol;  // (PB200_OPT_LINSOL can still change it)
    if (d->ism != 0) {
        // InternalSensitivityMethod: 0 / 2 staggered (the path's corrector), 1 simultaneous -- run by the
        // full-feature (steady-state) instantiation of either kernel family; 3 staggered1 is not built
        if (d->ism != 1 && d->ism != 2) return fail(PB200_ERR_ARG, "ism: 0 or 2 staggered, 1 simultaneous");
        if (d->ism == 1) {
            if (!ops->supports_steady)
                return fail(PB200_ERR_ARG, "ism = 1: the simultaneous corrector lives in the model's full-feature kernel "
                                           "instantiation, which this model was built without");
            if (!a.opt.errconS) return fail(PB200_ERR_ARG, "ism = 1 needs sensi_errcon = 1 (AMICI forces it)");
            a.opt.errconS |= 4;  // (SolverOptions::errconS bit 2)
            h->simultaneous = true;
            a.steady = 1;
        }
    }
    if (has_inf && a.opt.newton_maxsteps != 0 && !ops->supports_newton)
        return fail(PB200_ERR_ARG, "newton_maxsteps > 0 with t = inf output times: Newton's method for steady states is not "
                                   "supported for this model (only steady state by simulation, newton_maxsteps = 0)");
    a.num_trials = d->num_trials > 0 ? d->num_trials : 1;
    a.relaxation_factor = d->relaxation_factor;
    if (int rc = ensure_capacity(h, C > 0 ? C : 1)) return rc;
    if (ops->create_state) {
        std::string err;
        h->model_state = ops->create_state(h->sm_count, &err);
        if (!h->model_state) return fail(PB200_ERR_CUDA, "model state: " + err);
    }
    guard.release();
    *out = h;
    return 0;
}

void pb200_destroy(pb200_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    hier_destroy(h->hier);
    if (h->nccl) NcclApi::get().CommDestroy(h->nccl);
    if (h->model_state && h->ops && h->ops->destroy_state) h->ops->destroy_state(h->model_state);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    if (h->ev_staged) cudaEventDestroy(h->ev_staged);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int pb200_set_preequilibration(pb200_handle* h, const int32_t* has_preeq, const double* k_preeq,
                               const int64_t* reinit_offsets, const int32_t* reinit_idx) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    PB_CUDA(cudaSetDevice(h->device));
    if (h->stream) PB_CUDA(cudaStreamSynchronize(h->stream));
    KernelArgs& a = h->args;
    if (!has_preeq) {  // switch pre-equilibration off
        h->has_preeq = false;
        a.k_preeq = nullptr;
        a.preeq_flag = nullptr;
        a.reinit_mask = nullptr;
        a.reinit_off = nullptr;
        a.reinit_idx = nullptr;
        a.steady = (h->has_inf || h->simultaneous) ? 1 : 0;
        return 0;
    }
    if (!h->ops->supports_steady) return fail(PB200_ERR_ARG, "pre-equilibration is not supported for this model");
    if (h->steady_sens_newton && !h->ops->supports_newton)
        return fail(PB200_ERR_ARG, "steady_sens_newton with pre-equilibration: the linear solve for the sensitivities of a "
                                   "steady state is not supported for this model");
    if (a.opt.newton_maxsteps != 0 && !h->ops->supports_newton)
        return fail(PB200_ERR_ARG, "newton_maxsteps > 0 with pre-equilibration: Newton's method for steady states is not "
                                   "supported for this model (only steady state by simulation, newton_maxsteps = 0)");
    if (!k_preeq && h->nk > 0) return fail(PB200_ERR_ARG, "k_preeq is null");
    std::vector<int> flag(h->C);
    std::vector<unsigned> mask(h->C, 0u);
    std::vector<int64_t> roff(h->C + 1, 0);
    std::vector<int> ridx;
    bool any = false;
    for (int c = 0; c < h->C; ++c) {
        flag[c] = has_preeq[c] != 0;
        any = any || flag[c];
        roff[c + 1] = roff[c];
        if (!reinit_offsets) continue;
        if (reinit_offsets[c + 1] < reinit_offsets[c]) return fail(PB200_ERR_ARG, "reinit_offsets not monotone");
        for (int64_t e = reinit_offsets[c]; e < reinit_offsets[c + 1]; ++e) {
            const int i = reinit_idx[e];
            if (i < 0 || i >= h->ops->nx) return fail(PB200_ERR_ARG, "re-initialisation state index out of range");
            if (i < 32) mask[c] |= 1u << i;  // (the small-model kernels -- nx <= 32 -- take the list as a bit mask)
            ridx.push_back(i);
        }
        roff[c + 1] = (int64_t)ridx.size();
    }
    if (ridx.empty()) ridx.push_back(0);
    PB_CUDA(h->d_k_preeq.upload(k_preeq, (size_t)h->C * h->nk));
    PB_CUDA(h->d_preeq_flag.upload(flag.data(), flag.size()));
    PB_CUDA(h->d_reinit_mask.upload(mask.data(), mask.size()));
    PB_CUDA(h->d_reinit_off.upload(roff.data(), roff.size()));
    PB_CUDA(h->d_reinit_idx.upload(ridx.data(), ridx.size()));
    h->has_preeq = any;
    a.k_preeq = h->d_k_preeq.p;
    a.preeq_flag = h->d_preeq_flag.p;
    a.reinit_mask = h->d_reinit_mask.p;
    a.reinit_off = h->d_reinit_off.p;
    a.reinit_idx = h->d_reinit_idx.p;
    a.steady = (h->has_inf || any || h->simultaneous) ? 1 : 0;
    return 0;
}

int pb200_set_partition(pb200_handle* h, int rank, int world) {
    if (!h || world < 1 || rank < 0 || rank >= world) return fail(PB200_ERR_ARG, "bad partition");
    h->rank = rank;
    h->world = world;
    h->idx_cache.clear();
    return 0;
}

int pb200_set_allreduce(pb200_handle* h, pb200_allreduce_fn fn, void* user) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    h->allreduce = fn;
    h->allreduce_user = user;
    return 0;
}

int pb200_set_option(pb200_handle* h, int key, int value) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    switch (key) {
    case PB200_OPT_REORDER: h->reorder = value != 0; h->order_valid = false; break;
    case PB200_OPT_ORDER_REFRESH: h->opt_order_refresh = value > 0 ? value : 1; break;
    case PB200_OPT_KERNEL_VARIANT: h->opt_variant = value; break;
    case PB200_OPT_CTAS_PER_SM: h->opt_ctas_per_sm = value; break;
    case PB200_OPT_SYNC_PERIOD: h->opt_sync_period = value; break;
    case PB200_OPT_STAGE_SYNC: h->opt_stage_sync = value; break;
    case PB200_OPT_PARTIAL_SUMS: h->opt_partial_sums = value != 0; break;
    case PB200_OPT_LINSOL:
        if (value < 0 || value > 2) return fail(PB200_ERR_ARG, "linear solver: 0 automatic, 1 dense, 2 sparse");
        h->opt_linsol = value;
        break;
    case PB200_OPT_FUSE_OFF: h->opt_fuse_off = value; break;
    default: return fail(PB200_ERR_ARG, "unknown option");
    }
    return 0;
}

int pb200_nccl_unique_id(char* id128) {
    if (!id128) return fail(PB200_ERR_ARG, "null argument");
    NcclApi& nc = NcclApi::get();
    if (!nc.ok()) return fail(PB200_ERR_CUDA, nc.error);
    NcclApi::UniqueId id;
    const int rc = nc.GetUniqueId(&id);
    if (rc != NcclApi::Success) return fail(PB200_ERR_CUDA, std::string("ncclGetUniqueId: ") + nc.GetErrorString(rc));
    std::memcpy(id128, id.internal, 128);
    return 0;
}

int pb200_nccl_init(pb200_handle* h, const char* id128, int rank, int world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(PB200_ERR_ARG, "bad argument");
    NcclApi& nc = NcclApi::get();
    if (!nc.ok()) return fail(PB200_ERR_CUDA, nc.error);
    PB_CUDA(cudaSetDevice(h->device));
    if (h->nccl) {
        nc.CommDestroy(h->nccl);
        h->nccl = nullptr;
    }
    NcclApi::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    const int rc = nc.CommInitRank(&h->nccl, world, id, rank);
    if (rc != NcclApi::Success) {
        h->nccl = nullptr;
        return fail(PB200_ERR_CUDA, std::string("ncclCommInitRank: ") + nc.GetErrorString(rc));
    }
    PB_CUDA(h->d_coll.alloc(256));
    return pb200_set_partition(h, rank, world);
}

/// all-reduce of a small HOST buffer over the ranks of a partitioned evaluation: NCCL through a device
/// staging buffer if a communicator is set, otherwise the caller's hook
static int collective_host(pb200_handle* h, double* buf, int n, int op) {
    if (h->nccl) {
        NcclApi& nc = NcclApi::get();
        if ((size_t)n > h->d_coll.n) PB_CUDA(h->d_coll.alloc((size_t)n));
        PB_CUDA(cudaMemcpyAsync(h->d_coll.p, buf, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        const int rc = nc.AllReduce(h->d_coll.p, h->d_coll.p, (size_t)n, NcclApi::Float64,
                                    op == PB200_REDUCE_MAX ? NcclApi::Max : NcclApi::Sum, h->nccl, h->stream);
        if (rc != NcclApi::Success) return fail(PB200_ERR_CUDA, std::string("ncclAllReduce: ") + nc.GetErrorString(rc));
        PB_CUDA(cudaMemcpyAsync(buf, h->d_coll.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
        PB_CUDA(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (h->allreduce) {
        h->allreduce(h->allreduce_user, buf, n, op);
        return 0;
    }
    return fail(PB200_ERR_ARG, "no collective set");
}

int pb200_evaluate(pb200_handle* h, const double* theta, int n_opt, const int* idx, int n_idx, double* fval,
                   double* grad, double* seconds) {
    if (!h || !theta || !fval || (n_idx > 0 && !idx)) return fail(PB200_ERR_ARG, "null argument");
    if (n_opt != h->n_opt) return fail(PB200_ERR_ARG, "parameter vector size does not match numParameters");
    // A partitioned handle returns the GLOBAL fval / gradient / status (the reference master sums all
    // workers' results before evaluate returns: multiConditionProblem.cpp:738-748,775-824); rank-local
    // partial sums only on explicit request.
    if (h->world > 1 && !h->nccl && !h->allreduce && !h->opt_partial_sums)
        return fail(PB200_ERR_ARG, "partitioned evaluate needs a collective: pb200_nccl_init or pb200_set_allreduce "
                                   "(or PB200_OPT_PARTIAL_SUMS for rank-local sums)");
    const int want_grad = grad != nullptr;
    if (int rc = enqueue_evaluation(h, theta, idx, n_idx, want_grad, h->d_out.p, nullptr, nullptr, h->stream)) return rc;
    PB_CUDA(cudaMemcpyAsync(h->h_out.p, h->d_out.p, sizeof(double) * (2 + h->n_opt), cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
    h->last_d2h += (int64_t)sizeof(double) * (2 + h->n_opt);
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop);
        h->last_seconds = ms * 1e-3;
        if (seconds) *seconds = h->last_seconds;
    }
    if (h->world > 1 && !h->nccl && h->allreduce)  // host transport; NCCL has reduced d_out on the stream already
        h->allreduce(h->allreduce_user, h->h_out.p, 2 + h->n_opt, PB200_REDUCE_SUM);
    // the processing order is refreshed when the index list changed and every opt_order_refresh
    // evaluations after that (the ranking of the conditions by cost changes slowly with theta)
    const bool refresh = !h->order_valid || ++h->evals_since_order >= h->opt_order_refresh;
    if (h->reorder && h->last_n > 1 && refresh) {
        h->evals_since_order = 0;
        h->last_d2h += (int64_t)sizeof(int64_t) * h->last_n;
        h->last_h2d += (int64_t)sizeof(int) * h->last_n;
        // processing order of the next evaluation of the same index list: costliest conditions first
        // (they also sit next to conditions of similar cost in a warp / CTA)
        const int n = h->last_n;
        // only the step counts travel (one int64 out of the 16 of each record)
        h->stats_host.resize((size_t)n);
        PB_CUDA(cudaMemcpy2DAsync(h->stats_host.data(), sizeof(int64_t), h->d_stats.p + STAT_NST,
                                  sizeof(int64_t) * NSTATS, sizeof(int64_t), (size_t)n, cudaMemcpyDeviceToHost,
                                  h->stream));
        PB_CUDA(cudaStreamSynchronize(h->stream));
        h->order_host.resize(n);
        for (int r = 0; r < n; ++r) h->order_host[r] = r;
        const int64_t* st = h->stats_host.data();
        std::stable_sort(h->order_host.begin(), h->order_host.end(), [st](int x, int y) { return st[x] > st[y]; });
        PB_CUDA(cudaMemcpyAsync(h->d_order.p, h->order_host.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
        PB_CUDA(cudaStreamSynchronize(h->stream));
        h->order_valid = true;
    }
    // A1: zero-filled outputs, then sums; any error or non-finite -> +inf, failure
    *fval = h->h_out.p[0];
    if (grad)
        for (int j = 0; j < h->n_opt; ++j) grad[j] = h->h_out.p[2 + j];
    const double errors = h->h_out.p[1];
    if (errors != 0.0 || !std::isfinite(*fval)) {
        *fval = std::numeric_limits<double>::infinity();
        return PB200_EVAL_FAILURE;
    }
    return PB200_EVAL_SUCCESS;
}

int pb200_evaluate_device(pb200_handle* h, const double* theta, int n_opt, const int* idx, int n_idx, int want_grad,
                          double* d_out, void* stream) {
    if (!h || !theta || !d_out || (n_idx > 0 && !idx)) return fail(PB200_ERR_ARG, "null argument");
    if (n_opt != h->n_opt) return fail(PB200_ERR_ARG, "parameter vector size does not match numParameters");
    return enqueue_evaluation(h, theta, idx, n_idx, want_grad, d_out, nullptr, nullptr, (cudaStream_t)stream);
}

int pb200_linsol_info(pb200_handle* h, int64_t* out, int n) {
    if (!h || !out) return fail(PB200_ERR_ARG, "null argument");
    if (!h->ops->linsol_info) return 0;
    return h->ops->linsol_info(h->model_state, out, n);
}

int pb200_last_transfer_bytes(pb200_handle* h, int64_t* h2d, int64_t* d2h) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    if (h2d) *h2d = h->last_h2d;
    if (d2h) *d2h = h->last_d2h;
    return 0;
}

int pb200_last_launch(pb200_handle* h, int* lanes, int* grid, int* block, double* seconds) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    if (lanes) *lanes = h->last_lanes;
    if (grid) *grid = h->last_grid;
    if (block) *block = h->last_block;
    if (seconds) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop) != cudaSuccess) {
            cudaGetLastError();
            ms = -1e3f;
        }
        *seconds = ms * 1e-3;
    }
    return 0;
}

int pb200_get_condition_results(pb200_handle* h, double* llh, double* sllh, int* status, int64_t* stats) {
    if (!h) return fail(PB200_ERR_ARG, "null handle");
    PB_CUDA(cudaSetDevice(h->device));
    PB_CUDA(cudaDeviceSynchronize());
    const int n = h->last_n;
    if (llh) PB_CUDA(cudaMemcpy(llh, h->d_llh.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (sllh) {
        if (h->last_want_grad) PB_CUDA(cudaMemcpy(sllh, h->d_sllh.p, sizeof(double) * (size_t)n * h->np, cudaMemcpyDeviceToHost));
        else std::memset(sllh, 0, sizeof(double) * (size_t)n * h->np);
    }
    if (status) PB_CUDA(cudaMemcpy(status, h->d_status.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (stats) PB_CUDA(cudaMemcpy(stats, h->d_stats.p, sizeof(int64_t) * (size_t)n * NSTATS, cudaMemcpyDeviceToHost));
    return n;
}

// getModelOutputsAndSigmas (multiConditionProblem.cpp:506-563): all conditions, no sensitivities
static int run_outputs(pb200_handle* h, const double* theta, double* d_out) {
    const size_t ny_total = (size_t)h->nt_total * h->ny;
    if (h->d_yout.n < ny_total || !h->d_yout.p) {
        PB_CUDA(h->d_yout.alloc(ny_total > 0 ? ny_total : 1));
        PB_CUDA(h->d_sout.alloc(ny_total > 0 ? ny_total : 1));
    }
    // conditions this partition does not simulate stay NaN
    PB_CUDA(cudaMemsetAsync(h->d_yout.p, 0xFF, sizeof(double) * ny_total, h->stream));
    PB_CUDA(cudaMemsetAsync(h->d_sout.p, 0xFF, sizeof(double) * ny_total, h->stream));
    std::vector<int> all(h->C);
    for (int c = 0; c < h->C; ++c) all[c] = c;
    return enqueue_evaluation(h, theta, all.data(), h->C, 0, d_out, h->d_yout.p, h->d_sout.p, h->stream);
}

int pb200_get_model_outputs(pb200_handle* h, const double* theta, int n_opt, double* y_out, double* sigmay_out) {
    if (!h || !theta) return fail(PB200_ERR_ARG, "null argument");
    if (n_opt != h->n_opt) return fail(PB200_ERR_ARG, "parameter vector size does not match numParameters");
    if (int rc = run_outputs(h, theta, h->d_out.p)) return rc;
    const size_t ny_total = (size_t)h->nt_total * h->ny;
    PB_CUDA(cudaMemcpyAsync(h->h_out.p, h->d_out.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
    if (y_out) PB_CUDA(cudaMemcpyAsync(y_out, h->d_yout.p, sizeof(double) * ny_total, cudaMemcpyDeviceToHost, h->stream));
    if (sigmay_out) PB_CUDA(cudaMemcpyAsync(sigmay_out, h->d_sout.p, sizeof(double) * ny_total, cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
    return h->h_out.p[1] == 0.0 ? PB200_EVAL_SUCCESS : PB200_EVAL_FAILURE;
}

// ---- per-simulation result logging ---------------------------------
// Replaces: logSimulation / saveSimulation (multiConditionProblem.cpp:254-325) and the standalone simulator's
// result datasets (standaloneSimulator.cpp:57-270; doc/hdf5.md:175-252).  Records are collected in host memory
// and written once, as a classic-format HDF5 file (h5wmin.h).
}  // extern "C"

struct pb200_result_log {
    struct Group {
        int np = 0;
        std::vector<int32_t> job_id, status, trials;
        std::vector<std::string> label;
        std::vector<double> llh, seconds;
        std::vector<double> sllh, params;   // [record][np]
        std::vector<int64_t> stats;         // [record][14]
    };
    std::vector<std::pair<std::string, Group>> groups;  // in order of first use
    int next_job = 0;
    Group& get(const std::string& name, int np) {
        for (auto& g : groups)
            if (g.first == name) return g.second;
        groups.emplace_back(name, Group());
        groups.back().second.np = np;
        return groups.back().second;
    }
};

extern "C" {

int pb200_result_log_create(pb200_result_log** out) {
    if (!out) return fail(PB200_ERR_ARG, "null argument");
    *out = new pb200_result_log;
    return 0;
}
void pb200_result_log_destroy(pb200_result_log* log) { delete log; }

int pb200_result_log_append_records(pb200_result_log* log, const char* group, int n, int np, const int32_t* condition_idx,
                                    int optimization_idx, int iteration_idx, const double* llh, const double* sllh,
                                    const double* parameters, const double* seconds, const int32_t* status,
                                    const int64_t* stats) {
    if (!log || !group || n < 0 || np < 0 || (n > 0 && (!condition_idx || !llh || !status)))
        return fail(PB200_ERR_ARG, "null argument");
    pb200_result_log::Group& g = log->get(group, np);
    if (g.np != np) return fail(PB200_ERR_ARG, "result log: records of one group must have the same number of model parameters");
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int r = 0; r < n; ++r) {
        g.job_id.push_back(log->next_job++);
        // simulation label: optimisation index 'o', iteration index 'i', condition index 'c' (doc/hdf5.md:208-210)
        g.label.push_back("o" + std::to_string(optimization_idx) + "i" + std::to_string(iteration_idx) + "c" +
                          std::to_string(condition_idx[r]));
        g.llh.push_back(llh[r]);
        g.status.push_back(status[r]);
        g.seconds.push_back(seconds ? seconds[r] : nan);
        for (int i = 0; i < np; ++i) {
            g.sllh.push_back(sllh ? sllh[(size_t)r * np + i] : nan);
            g.params.push_back(parameters ? parameters[(size_t)r * np + i] : nan);
        }
        g.trials.push_back(stats ? (int32_t)stats[(size_t)r * PB200_NSTATS + PB200_STAT_TRIALS] : 0);
        for (int k = 0; k < 14; ++k) g.stats.push_back(stats ? stats[(size_t)r * PB200_NSTATS + k] : 0);
    }
    return n;
}

int pb200_result_log_append(pb200_result_log* log, pb200_handle* h, const char* group, int optimization_idx,
                            int iteration_idx) {
    if (!log || !h || !group) return fail(PB200_ERR_ARG, "null argument");
    const int n = h->last_n, np = h->np;
    std::vector<double> llh(n), sllh((size_t)n * np), params((size_t)n * np), seconds(n);
    std::vector<int> status(n);
    std::vector<int64_t> stats((size_t)n * NSTATS);
    const int rc = pb200_get_condition_results(h, llh.data(), sllh.data(), status.data(), stats.data());
    if (rc < 0) return rc;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<int32_t> cidx(n);
    for (int r = 0; r < n; ++r) {
        const int c = h->idx_cache[r];
        cidx[r] = c;
        // the condition's model parameters in simulation scale (mapAndSetOptimizationToSimulationVariables,
        // multiConditionDataProvider.cpp:126-169) from the parameter table of the last evaluation
        int k = 0;
        for (int i = 0; i < np; ++i) {
            const int j = h->mapping_host[(size_t)c * np + i];
            double v = nan;
            if (j >= 0) v = h->h_theta_tab.p[((size_t)j * 3 + h->pscale_sim_host[(size_t)c * np + i]) * 2 + 0];
            else if (!h->overrides_host.empty()) v = h->overrides_host[(size_t)c * np + i];
            params[(size_t)r * np + i] = v;
            if (j >= 0) ++k;
        }
        // sllh is plist-compacted: entries beyond nplist are padding -> NaN, as saveSimulation pads
        for (int i = (h->last_want_grad ? k : 0); i < np; ++i) sllh[(size_t)r * np + i] = nan;
        // conditions run side by side on the GPU: the evaluation's device time shared equally
        seconds[r] = n > 0 ? h->last_seconds / n : 0.0;
    }
    return pb200_result_log_append_records(log, group, n, np, cidx.data(), optimization_idx, iteration_idx, llh.data(),
                                           sllh.data(), params.data(), seconds.data(), status.data(), stats.data());
}

int pb200_result_log_num_records(pb200_result_log* log, const char* group) {
    if (!log || !group) return fail(PB200_ERR_ARG, "null argument");
    for (auto& g : log->groups)
        if (g.first == group) return (int)g.second.llh.size();
    return 0;
}

int pb200_result_log_write(pb200_result_log* log, const char* path) {
    if (!log || !path) return fail(PB200_ERR_ARG, "null argument");
    try {
        H5Writer w;
        for (auto& kv : log->groups) {
            const std::string& gp = kv.first;
            const pb200_result_log::Group& g = kv.second;
            const uint64_t n = g.llh.size(), np = (uint64_t)g.np;
            // one COLUMN per simulation, as hdf5CreateOrExtendAndWriteTo*2DArray lays the records out
            auto transpose = [&](const std::vector<double>& v) {
                std::vector<double> t(v.size());
                for (uint64_t r = 0; r < n; ++r)
                    for (uint64_t i = 0; i < np; ++i) t[i * n + r] = v[r * np + i];
                return t;
            };
            w.doubles(gp + "/simulationLogLikelihood", g.llh, {1, n});
            w.ints(gp + "/jobId", g.job_id, {1, n});
            w.doubles(gp + "/simulationLogLikelihoodGradient", transpose(g.sllh), {np, n});
            w.doubles(gp + "/simulationParameters", transpose(g.params), {np, n});
            w.doubles(gp + "/simulationWallTimeInSec", g.seconds, {1, n});
            w.ints(gp + "/simulationStatus", g.status, {1, n});
            w.strings(gp + "/simulationLabel", g.label);
            // beyond the reference's datasets: the tolerance-relaxation retry count and CVODES' counters
            w.ints(gp + "/simulationTrials", g.trials, {1, n});
            std::vector<int64_t> st(g.stats.size());
            for (uint64_t r = 0; r < n; ++r)
                for (uint64_t k = 0; k < 14; ++k) st[k * n + r] = g.stats[r * 14 + k];
            w.int64s(gp + "/simulationIntegratorStatistics", st, {14, n});
        }
        w.save(path);
    } catch (std::exception const& e) {
        return fail(PB200_ERR_ARG, e.what());
    }
    return 0;
}

int pb200_write_simulation_results(pb200_handle* h, const double* theta, int n_opt, const char* path, const char* group) {
    if (!h || !theta || !path || !group) return fail(PB200_ERR_ARG, "null argument");
    if (n_opt != h->n_opt) return fail(PB200_ERR_ARG, "parameter vector size does not match numParameters");
    if (h->world > 1) return fail(PB200_ERR_ARG, "simulation result files are written from an unpartitioned handle");
    const size_t ny_total = (size_t)h->nt_total * h->ny;
    std::vector<double> y(ny_total > 0 ? ny_total : 1), s(ny_total > 0 ? ny_total : 1), t(h->nt_total > 0 ? h->nt_total : 1),
        my(ny_total > 0 ? ny_total : 1), llh(h->C > 0 ? h->C : 1);
    const int rc = pb200_get_model_outputs(h, theta, n_opt, y.data(), s.data());
    if (rc < 0) return rc;
    if (pb200_get_condition_results(h, llh.data(), nullptr, nullptr, nullptr) < 0) return PB200_ERR_CUDA;
    PB_CUDA(cudaMemcpy(t.data(), h->d_t.p, sizeof(double) * h->nt_total, cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(my.data(), h->d_my.p, sizeof(double) * ny_total, cudaMemcpyDeviceToHost));
    try {
        H5Writer w;
        const std::string g = group;
        llh.resize(h->C);
        w.doubles(g + "/llh", llh, {(uint64_t)h->C});
        w.doubles(g + "/parameters", std::vector<double>(theta, theta + n_opt), {(uint64_t)n_opt});
        for (int c = 0; c < h->C; ++c) {
            const int64_t lo = h->t_offsets[c], hi = h->t_offsets[c + 1];
            const uint64_t nt = (uint64_t)(hi - lo), ny = (uint64_t)h->ny;
            const std::string id = std::to_string(c);
            w.doubles(g + "/t/" + id, std::vector<double>(t.begin() + lo, t.begin() + hi), {nt});
            w.doubles(g + "/yMes/" + id, std::vector<double>(my.begin() + lo * h->ny, my.begin() + hi * h->ny), {nt, ny});
            w.doubles(g + "/ySim/" + id, std::vector<double>(y.begin() + lo * h->ny, y.begin() + hi * h->ny), {nt, ny});
        }
        w.save(path);
    } catch (std::exception const& e) {
        return fail(PB200_ERR_ARG, e.what());
    }
    return rc;
}

// ---- hierarchical optimisation (A15) -----------------------------------------------
int pb200_hier_attach(pb200_handle* h, const pb200_hier_desc* hd) {
    if (!h || !hd) return fail(PB200_ERR_ARG, "null argument");
    PB_CUDA(cudaSetDevice(h->device));
    hier_destroy(h->hier);
    h->hier = nullptr;
    std::string err;
    HierDevice* hv = hier_create(*hd, h->C, h->n_opt, h->ny, h->t_offsets.data(), h->pscale_opt.data(), &err);
    if (!hv) return fail(PB200_ERR_ARG, err);
    h->hier = hv;
    return 0;
}

int pb200_hier_num_outer(pb200_handle* h) {
    if (!h || !h->hier) return fail(PB200_ERR_ARG, "no hierarchical description attached");
    return h->n_opt - hier_num_inner(h->hier);
}

int pb200_hier_evaluate(pb200_handle* h, const double* theta_outer, int n_outer, double* fval, double* grad_outer,
                        double* full_parameters, double* seconds) {
    if (!h || !h->hier || !theta_outer || !fval) return fail(PB200_ERR_ARG, "null argument");
    const int n_inner = hier_num_inner(h->hier);
    // HierarchicalOptimizationWrapper::evaluate throws on size mismatch (hierarchicalOptimization.cpp:155-160)
    if (n_outer != h->n_opt - n_inner) return fail(PB200_ERR_ARG, "reduced parameter vector size does not match numParameters");
    if (h->world > 1 && !h->allreduce && !h->nccl)
        return fail(PB200_ERR_ARG, "partitioned hierarchical evaluation needs pb200_nccl_init or pb200_set_allreduce");
    PB_CUDA(cudaSetDevice(h->device));
    std::function<void(double*, int, int)> reduce;
    int coll_rc = 0;
    if (h->world > 1) reduce = [h, &coll_rc](double* buf, int n, int op) { if (int rc = collective_host(h, buf, n, op)) coll_rc = rc; };
    int c_lo = 0, c_hi = h->C;
    pb200_partition_bounds(h->C, h->rank, h->world, &c_lo, &c_hi);
    double t_total = 0.0;
    // pass 1: defaults for the inner parameters (scaling 1, offset 0, sigma 1 in linear scale)
    std::vector<double> full(h->n_opt);
    hier_splice_defaults(h->hier, theta_outer, full.data());
    if (int rc = run_outputs(h, full.data(), h->d_out.p)) return rc;
    PB_CUDA(cudaMemcpyAsync(h->h_out.p, h->d_out.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
    PB_CUDA(cudaStreamSynchronize(h->stream));
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop);
        t_total += ms * 1e-3;
    }
    double nerr1 = h->h_out.p[1];
    if (reduce && !h->nccl) reduce(&nerr1, 1, PB200_REDUCE_SUM);  // (NCCL: reduced on the stream with d_out)
    if (nerr1 != 0.0) {  // "Function evaluation failed." -> functionEvaluationFailure
        *fval = std::numeric_limits<double>::infinity();
        return PB200_EVAL_FAILURE;
    }
    // analytic inner parameters on the device from the resident outputs
    std::vector<double> inner_scaled;  // optimiser scale, order: scalings, offsets, sigmas
    double nllh_analytic = 0.0;
    std::string err;
    if (!hier_compute(h->hier, h->d_my.p, h->d_yout.p, h->d_sout.p, h->stream, grad_outer == nullptr, &inner_scaled,
                      &nllh_analytic, &err, reduce, c_lo, c_hi))
        return fail(PB200_ERR_CUDA, err);
    if (coll_rc) return coll_rc;
    hier_splice(h->hier, theta_outer, inner_scaled.data(), full.data());
    if (full_parameters) std::memcpy(full_parameters, full.data(), sizeof(double) * h->n_opt);
    int status;
    if (grad_outer) {
        // pass 2: full evaluation with sensitivities, gradient filtered to the outer parameters
        std::vector<int> all(h->C);
        for (int c = 0; c < h->C; ++c) all[c] = c;
        std::vector<double> full_grad(h->n_opt);
        double sec2 = 0.0;
        // (a partitioned handle all-reduces (nllh, #errors, gradient) inside pb200_evaluate)
        status = pb200_evaluate(h, full.data(), h->n_opt, all.data(), h->C, fval, full_grad.data(), &sec2);
        if (status < 0) return status;
        t_total += sec2;
        if (status == PB200_EVAL_SUCCESS) hier_filter(h->hier, full_grad.data(), grad_outer);
    } else {
        *fval = nllh_analytic;
        status = std::isfinite(*fval) ? PB200_EVAL_SUCCESS : PB200_EVAL_FAILURE;
    }
    if (seconds) *seconds = t_total;
    return status;
}

}  // extern "C"
