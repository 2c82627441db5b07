// hier_kernels.cuh -- analytic inner problem of parPE's hierarchical optimisation on the
// device.  Replaces the host loops of
// src/parpeamici/hierarchicalOptimization.cpp: computeAnalyticalScalings :702-794,
// computeAnalyticalOffsets :796-851, computeAnalyticalSigmas :853-930,
// applyOptimalScaling/Offset :932-987, fillInAnalyticalSigmas :403-441,
// computeNegLogLikelihood :1032-1046,1095-1134, spliceParameters :989-1029,
// fillFilteredParams :649-669.  The model outputs of pass 1 never leave HBM: the sums run
// as one block per inner parameter over its (condition, observable) pairs with a fixed
// thread->element assignment and a shared-memory tree, so results are run-to-run
// reproducible.  Only 2-3 doubles per inner parameter travel to the host, which applies
// the scalar rules (zero denominators, sigma floor, parameter scale) with the same libm
// calls as the reference.
#ifndef PB200_HIER_KERNELS_CUH
#define PB200_HIER_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include <functional>

#include "../../include/parpe_b200.h"

namespace pb200 {

struct HierKind {
    int n = 0;
    std::vector<int> idx;        // optimisation-parameter indices
    std::vector<int64_t> ptr;    // [n+1] CSR over pairs
    std::vector<int> cond, obs;  // pairs: conditions ascending per parameter, observables in row order
    int64_t* d_ptr = nullptr;
    int *d_cond = nullptr, *d_obs = nullptr;
    double* d_sums = nullptr;    // [n * HIER_SLICES * 3]: per inner parameter and slice of its pair list
    double* d_vals = nullptr;    // [n] linear-scale values to apply
};

struct HierDevice {
    HierKind scaling, offset, sigma;
    int C = 0, n_opt = 0, ny = 0;
    std::vector<int> pscale_opt;
    std::vector<int> inner_sorted;  // all analytical parameter indices, ascending
    int64_t* d_t_offsets = nullptr;
    std::vector<int64_t> t_offsets;  // host copy
    int64_t nt_total = 0;
    double* d_partial = nullptr;    // nllh partial sums
    int n_partial = 0;
};

enum { HIER_SCALING = 0, HIER_OFFSET = 1, HIER_SIGMA = 2 };
/// every inner parameter's (condition, observable) pair list is summed by this many blocks (a fixed split:
/// the partial sums are combined on the host in slice order, so results do not depend on the launch)
constexpr int HIER_SLICES = 32;

template <int KIND, int THREADS>
__global__ void __launch_bounds__(THREADS)
hier_sums_kernel(const int64_t* __restrict__ ptr, const int* __restrict__ cond, const int* __restrict__ obs,
                 const int64_t* __restrict__ t_offsets, int ny, const double* __restrict__ mes,
                 const double* __restrict__ sim, double* __restrict__ sums, int c_lo, int c_hi) {
    __shared__ double r0[THREADS], r1[THREADS], r2[THREADS];
    const int p = blockIdx.x, slice = blockIdx.y, tid = threadIdx.x;
    double num = 0.0, den = 0.0, mx = 0.0;
    const int64_t len = ptr[p + 1] - ptr[p];
    const int64_t chunk = (len + HIER_SLICES - 1) / HIER_SLICES;
    const int64_t e_lo = ptr[p] + slice * chunk;
    const int64_t e_hi = (e_lo + chunk < ptr[p + 1]) ? e_lo + chunk : ptr[p + 1];
    for (int64_t e = e_lo + tid; e < e_hi; e += THREADS) {
        const int c = cond[e], o = obs[e];
        if (c < c_lo || c >= c_hi) continue;  // a condition another rank simulates (ownership by partition bounds)
        const int64_t lo = t_offsets[c], hi = t_offsets[c + 1];
        for (int64_t it = lo; it < hi; ++it) {
            const double m = mes[it * ny + o];
            if (isnan(m)) continue;
            double s = sim[it * ny + o];
            if (KIND == HIER_SCALING) {
                if (s < 0 && s > -1e-18) s = 0.0;
                num += s * m;
                den += s * s;
            } else if (KIND == HIER_OFFSET) {
                num += m - s;
                den += 1.0;
            } else {
                num += (m - s) * (m - s);
                den += 1.0;
                mx = fmax(mx, fabs(m));
            }
        }
    }
    r0[tid] = num;
    r1[tid] = den;
    r2[tid] = mx;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            r0[tid] += r0[tid + s];
            r1[tid] += r1[tid + s];
            r2[tid] = fmax(r2[tid], r2[tid + s]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* o = sums + ((size_t)p * HIER_SLICES + slice) * 3;
        o[0] = r0[0];
        o[1] = r1[0];
        o[2] = r2[0];
    }
}

/// MODE 0: out *= v   MODE 1: out += v   MODE 2: out = v   on every (pair, time point)
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
hier_apply_kernel(const int64_t* __restrict__ ptr, const int* __restrict__ cond, const int* __restrict__ obs,
                  const int64_t* __restrict__ t_offsets, int ny, const double* __restrict__ vals,
                  double* __restrict__ out) {
    const int p = blockIdx.x;
    const double v = vals[p];
    for (int64_t e = ptr[p] + threadIdx.x; e < ptr[p + 1]; e += THREADS) {
        const int c = cond[e], o = obs[e];
        for (int64_t it = t_offsets[c]; it < t_offsets[c + 1]; ++it) {
            double& x = out[it * ny + o];
            if (MODE == 0) x *= v;
            else if (MODE == 1) x += v;
            else x = v;
        }
    }
}

/// computeNegLogLikelihood over the data points [first, last) (those of this rank's conditions): per-block partials
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
hier_nllh_kernel(int64_t first, int64_t last, const double* __restrict__ mes, const double* __restrict__ sim,
                 const double* __restrict__ sig, double* __restrict__ partial) {
    __shared__ double r0[THREADS];
    const int tid = threadIdx.x;
    double acc = 0.0;
    for (int64_t i = first + (int64_t)blockIdx.x * THREADS + tid; i < last; i += (int64_t)gridDim.x * THREADS) {
        const double m = mes[i];
        if (isnan(m)) continue;
        const double s = sim[i];
        const double s2 = sig[i] * sig[i];
        if (isnan(s) || isnan(s2) || s2 < 0.0) {
            acc = NAN;
            continue;
        }
        const double d = m - s;
        acc += log(2.0 * 3.14159265358979323846 * s2) + d * d / s2;
    }
    r0[tid] = acc;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) r0[tid] += r0[tid + s];
        __syncthreads();
    }
    if (tid == 0) partial[blockIdx.x] = r0[0];
}

inline bool hier_build_kind(HierKind& k, const int32_t* idx, int n, const int32_t* rows, int n_rows, int C, int ny,
                            int n_opt, std::string* err) {
    k.n = n;
    k.idx.assign(idx, idx + n);
    for (int i = 0; i < n; ++i) {
        if (idx[i] < 0 || idx[i] >= n_opt) { *err = "analytical parameter index out of range"; return false; }
        if (i > 0 && idx[i] <= idx[i - 1]) { *err = "analytical parameter indices must be ascending"; return false; }
    }
    // AnalyticalParameterHdf5Reader::readParameterConditionObservableMappingFromFile
    std::vector<std::map<int, std::vector<int>>> mapping(n);
    for (int i = 0; i < n_rows; ++i) {
        const int p = rows[i * 3], c = rows[i * 3 + 1], o = rows[i * 3 + 2];
        if (p < 0 || p >= n || c < 0 || c >= C || o < 0 || o >= ny) { *err = "analytical parameter map row out of range"; return false; }
        mapping[p][c].push_back(o);
    }
    k.ptr.assign(1, 0);
    for (int p = 0; p < n; ++p) {
        for (auto const& kv : mapping[p])
            for (int o : kv.second) {
                k.cond.push_back(kv.first);
                k.obs.push_back(o);
            }
        k.ptr.push_back((int64_t)k.cond.size());
    }
    auto up = [&](void** d, const void* s, size_t bytes) {
        if (cudaMalloc(d, bytes ? bytes : 8) != cudaSuccess) return false;
        return bytes == 0 || cudaMemcpy(*d, s, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up((void**)&k.d_ptr, k.ptr.data(), sizeof(int64_t) * k.ptr.size()) ||
        !up((void**)&k.d_cond, k.cond.data(), sizeof(int) * k.cond.size()) ||
        !up((void**)&k.d_obs, k.obs.data(), sizeof(int) * k.obs.size()) ||
        cudaMalloc((void**)&k.d_sums, sizeof(double) * 3 * HIER_SLICES * (n ? n : 1)) != cudaSuccess ||
        cudaMalloc((void**)&k.d_vals, sizeof(double) * (n ? n : 1)) != cudaSuccess) {
        *err = "cudaMalloc failed for hierarchical tables";
        return false;
    }
    return true;
}

inline void hier_free_kind(HierKind& k) {
    cudaFree(k.d_ptr); cudaFree(k.d_cond); cudaFree(k.d_obs); cudaFree(k.d_sums); cudaFree(k.d_vals);
}

inline void hier_destroy(HierDevice* h) {
    if (!h) return;
    hier_free_kind(h->scaling); hier_free_kind(h->offset); hier_free_kind(h->sigma);
    cudaFree(h->d_t_offsets);
    cudaFree(h->d_partial);
    delete h;
}

inline HierDevice* hier_create(const pb200_hier_desc& hd, int C, int n_opt, int ny, const int64_t* t_offsets,
                               const int* pscale_opt, std::string* err) {
    auto* h = new HierDevice;
    h->C = C; h->n_opt = n_opt; h->ny = ny;
    h->pscale_opt.assign(pscale_opt, pscale_opt + n_opt);
    h->nt_total = t_offsets[C];
    h->t_offsets.assign(t_offsets, t_offsets + C + 1);
    bool ok = hier_build_kind(h->scaling, hd.scaling_idx, hd.n_scaling, hd.scaling_map, hd.n_scaling_rows, C, ny, n_opt, err) &&
              hier_build_kind(h->offset, hd.offset_idx, hd.n_offset, hd.offset_map, hd.n_offset_rows, C, ny, n_opt, err) &&
              hier_build_kind(h->sigma, hd.sigma_idx, hd.n_sigma, hd.sigma_map, hd.n_sigma_rows, C, ny, n_opt, err);
    if (ok) {
        // getDefaultScalingFactor / getDefaultOffsetParameter accept only lin and log10
        for (auto* k : {&h->scaling, &h->offset, &h->sigma})
            for (int i : k->idx)
                if (pscale_opt[i] == PB200_SCALE_LN) { *err = "Parameter scaling must be log10 or none."; ok = false; }
    }
    if (ok) {
        h->inner_sorted = h->scaling.idx;
        h->inner_sorted.insert(h->inner_sorted.end(), h->offset.idx.begin(), h->offset.idx.end());
        h->inner_sorted.insert(h->inner_sorted.end(), h->sigma.idx.begin(), h->sigma.idx.end());
        std::sort(h->inner_sorted.begin(), h->inner_sorted.end());
        for (size_t i = 1; i < h->inner_sorted.size(); ++i)
            if (h->inner_sorted[i] == h->inner_sorted[i - 1]) { *err = "analytical parameter listed twice"; ok = false; }
    }
    h->n_partial = 296;
    if (ok && (cudaMalloc((void**)&h->d_t_offsets, sizeof(int64_t) * (C + 1)) != cudaSuccess ||
               cudaMemcpy(h->d_t_offsets, t_offsets, sizeof(int64_t) * (C + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
               cudaMalloc((void**)&h->d_partial, sizeof(double) * h->n_partial) != cudaSuccess)) {
        *err = "cudaMalloc failed for hierarchical tables";
        ok = false;
    }
    if (!ok) { hier_destroy(h); return nullptr; }
    return h;
}

inline int hier_num_inner(const HierDevice* h) { return (int)h->inner_sorted.size(); }

namespace hier_detail {
inline double unscale(double v, int s) { return s == PB200_SCALE_LOG10 ? std::pow(10, v) : s == PB200_SCALE_LN ? std::exp(v) : v; }
inline double scale(double v, int s) { return s == PB200_SCALE_LOG10 ? std::log10(v) : s == PB200_SCALE_LN ? std::log(v) : v; }
}  // namespace hier_detail

/// spliceParameters (hierarchicalOptimization.cpp:989-1029); inner = scalings, offsets, sigmas
inline void hier_splice(const HierDevice* h, const double* outer, const double* inner, double* full) {
    size_t is = 0, io = 0, ig = 0, ir = 0;
    const size_t ns = h->scaling.idx.size(), no = h->offset.idx.size();
    for (int i = 0; i < h->n_opt; ++i) {
        if (is < ns && h->scaling.idx[is] == i) full[i] = inner[is++];
        else if (io < no && h->offset.idx[io] == i) full[i] = inner[ns + io++];
        else if (ig < h->sigma.idx.size() && h->sigma.idx[ig] == i) full[i] = inner[ns + no + ig++];
        else full[i] = outer[ir++];
    }
}

/// getDefault{ScalingFactors,OffsetParameters,SigmaParameters} (:240-278, :672-700) + splice
inline void hier_splice_defaults(const HierDevice* h, const double* outer, double* full) {
    std::vector<double> inner;
    for (int i : h->scaling.idx) inner.push_back(h->pscale_opt[i] == PB200_SCALE_LOG10 ? 0.0 : 1.0);
    for (int i : h->offset.idx)
        inner.push_back(h->pscale_opt[i] == PB200_SCALE_LOG10 ? -std::numeric_limits<double>::infinity() : 0.0);
    for (int i : h->sigma.idx) inner.push_back(h->pscale_opt[i] == PB200_SCALE_LOG10 ? 0.0 : 1.0);
    hier_splice(h, outer, inner.data(), full);
}

/// fillFilteredParams (:649-669)
inline void hier_filter(const HierDevice* h, const double* full, double* outer) {
    size_t nf = 0, ri = 0;
    for (int i = 0; i < h->n_opt; ++i) {
        if (nf < h->inner_sorted.size() && h->inner_sorted[nf] == i) ++nf;
        else outer[ri++] = full[i];
    }
}

/// Steps 3-4 of HierarchicalOptimizationWrapper::evaluate (:184-238).  d_y (pass-1 outputs)
/// is scaled/offset in place; d_sig receives the analytic sigmas when the analytic nllh is
/// requested.  inner_scaled: scalings, offsets, sigmas in the optimiser's parameter scale.
/// reduce(buf, n, op): element-wise all-reduce over the ranks of a partitioned evaluation (null: one rank);
/// [c_lo, c_hi): the conditions this rank simulated (the others' outputs are not touched: a genuine NaN
/// output of an owned condition propagates, as in the reference)
inline bool hier_compute(HierDevice* h, const double* d_mes, double* d_y, double* d_sig, cudaStream_t stream,
                         bool want_nllh, std::vector<double>* inner_scaled, double* nllh, std::string* err,
                         const std::function<void(double*, int, int)>& reduce = nullptr, int c_lo = 0, int c_hi = -1) {
    if (c_hi < 0) c_hi = h->C;
    // sums of one kind over all ranks: entries 0,1 are sums, entry 2 a maximum
    auto reduce_kind = [&](std::vector<double>& host, int n) {
        if (!reduce || n == 0) return;
        std::vector<double> sm(2 * (size_t)n), mx(n);
        for (int i = 0; i < n; ++i) { sm[2 * i] = host[3 * i]; sm[2 * i + 1] = host[3 * i + 1]; mx[i] = host[3 * i + 2]; }
        reduce(sm.data(), 2 * n, PB200_REDUCE_SUM);
        reduce(mx.data(), n, PB200_REDUCE_MAX);
        for (int i = 0; i < n; ++i) { host[3 * i] = sm[2 * i]; host[3 * i + 1] = sm[2 * i + 1]; host[3 * i + 2] = mx[i]; }
    };
    using namespace hier_detail;
    constexpr int T = 256;
    auto check = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess) { *err = std::string(what) + ": " + cudaGetErrorString(e); return false; }
        return true;
    };
    auto fetch = [&](HierKind& k, std::vector<double>& host) {
        host.resize((size_t)3 * k.n);
        if (k.n == 0) return true;
        std::vector<double> part((size_t)3 * HIER_SLICES * k.n);
        if (!(check(cudaMemcpyAsync(part.data(), k.d_sums, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, stream), "memcpy") &&
              check(cudaStreamSynchronize(stream), "sync")))
            return false;
        for (int p = 0; p < k.n; ++p) {  // slices in fixed order
            double a = 0.0, b = 0.0, mx = 0.0;
            for (int sl = 0; sl < HIER_SLICES; ++sl) {
                const double* o = &part[((size_t)p * HIER_SLICES + sl) * 3];
                a += o[0];
                b += o[1];
                mx = std::max(mx, o[2]);
            }
            host[3 * p] = a; host[3 * p + 1] = b; host[3 * p + 2] = mx;
        }
        return true;
    };
    auto push = [&](HierKind& k, const std::vector<double>& lin) {
        if (k.n == 0) return true;
        return check(cudaMemcpyAsync(k.d_vals, lin.data(), sizeof(double) * k.n, cudaMemcpyHostToDevice, stream), "memcpy") &&
               check(cudaStreamSynchronize(stream), "sync");
    };
    HierKind &S = h->scaling, &O = h->offset, &G = h->sigma;
    std::vector<double> hs, ho, hg;
    if (S.n) hier_sums_kernel<HIER_SCALING, T><<<dim3(S.n, HIER_SLICES), T, 0, stream>>>(S.d_ptr, S.d_cond, S.d_obs, h->d_t_offsets, h->ny, d_mes, d_y, S.d_sums, c_lo, c_hi);
    if (O.n) hier_sums_kernel<HIER_OFFSET, T><<<dim3(O.n, HIER_SLICES), T, 0, stream>>>(O.d_ptr, O.d_cond, O.d_obs, h->d_t_offsets, h->ny, d_mes, d_y, O.d_sums, c_lo, c_hi);
    if (!check(cudaGetLastError(), "hier_sums_kernel") || !fetch(S, hs) || !fetch(O, ho)) return false;
    reduce_kind(hs, S.n);
    reduce_kind(ho, O.n);
    std::vector<double> scal(S.n), offs(O.n), sig(G.n), lin;
    for (int i = 0; i < S.n; ++i) {
        const double den = hs[i * 3 + 1];
        const double v = den == 0.0 ? 1.0 : hs[i * 3] / den;
        scal[i] = scale(v, h->pscale_opt[S.idx[i]]);
    }
    for (int i = 0; i < O.n; ++i) {  // from the UNSCALED outputs (quirk Q5)
        const double den = ho[i * 3 + 1];
        const double v = den == 0.0 ? 0.0 : ho[i * 3] / den;
        offs[i] = scale(v, h->pscale_opt[O.idx[i]]);
    }
    // applyOptimalScalings / applyOptimalOffsets use getUnscaledParameter(scaled value)
    lin.resize(S.n);
    for (int i = 0; i < S.n; ++i) lin[i] = unscale(scal[i], h->pscale_opt[S.idx[i]]);
    if (!push(S, lin)) return false;
    if (S.n) hier_apply_kernel<0, T><<<S.n, T, 0, stream>>>(S.d_ptr, S.d_cond, S.d_obs, h->d_t_offsets, h->ny, S.d_vals, d_y);
    lin.resize(O.n);
    for (int i = 0; i < O.n; ++i) lin[i] = unscale(offs[i], h->pscale_opt[O.idx[i]]);
    if (!push(O, lin)) return false;
    if (O.n) hier_apply_kernel<1, T><<<O.n, T, 0, stream>>>(O.d_ptr, O.d_cond, O.d_obs, h->d_t_offsets, h->ny, O.d_vals, d_y);
    if (G.n) hier_sums_kernel<HIER_SIGMA, T><<<dim3(G.n, HIER_SLICES), T, 0, stream>>>(G.d_ptr, G.d_cond, G.d_obs, h->d_t_offsets, h->ny, d_mes, d_y, G.d_sums, c_lo, c_hi);
    if (!check(cudaGetLastError(), "hier kernels") || !fetch(G, hg)) return false;
    reduce_kind(hg, G.n);
    for (int i = 0; i < G.n; ++i) {
        double sigma = std::sqrt(hg[i * 3] / hg[i * 3 + 1]);
        const double eps = std::max(0.01 * hg[i * 3 + 2], 1e-12);
        if (sigma < eps) sigma = eps;
        sig[i] = scale(sigma, h->pscale_opt[G.idx[i]]);
    }
    inner_scaled->clear();
    inner_scaled->insert(inner_scaled->end(), scal.begin(), scal.end());
    inner_scaled->insert(inner_scaled->end(), offs.begin(), offs.end());
    inner_scaled->insert(inner_scaled->end(), sig.begin(), sig.end());
    if (want_nllh) {
        lin.resize(G.n);
        for (int i = 0; i < G.n; ++i) lin[i] = unscale(sig[i], h->pscale_opt[G.idx[i]]);
        if (!push(G, lin)) return false;
        if (G.n) hier_apply_kernel<2, T><<<G.n, T, 0, stream>>>(G.d_ptr, G.d_cond, G.d_obs, h->d_t_offsets, h->ny, G.d_vals, d_sig);
        hier_nllh_kernel<T><<<h->n_partial, T, 0, stream>>>(h->t_offsets[c_lo] * h->ny, h->t_offsets[c_hi] * h->ny, d_mes, d_y, d_sig, h->d_partial);
        std::vector<double> part(h->n_partial);
        if (!check(cudaGetLastError(), "hier_nllh_kernel") ||
            !check(cudaMemcpyAsync(part.data(), h->d_partial, sizeof(double) * h->n_partial, cudaMemcpyDeviceToHost, stream), "memcpy") ||
            !check(cudaStreamSynchronize(stream), "sync"))
            return false;
        double acc = 0.0;
        for (double v : part) acc += v;
        if (reduce) reduce(&acc, 1, PB200_REDUCE_SUM);
        *nllh = acc / 2.0;
    }
    return true;
}

}  // namespace pb200
#endif
