// team.cuh -- the cooperating threads of ONE experimental condition in the large-model
// ("CTA per condition") path: a whole thread block on the device, a single thread in the
// host build that tests/host_harness uses to diff the algorithm against the CVODES oracle.
#ifndef PB200_TEAM_CUH
#define PB200_TEAM_CUH

#include <math.h>
#include <stdint.h>

#include "pb_hd.h"

namespace pb200 {

/// host: one thread does everything (test harness only)
struct HostTeam {
    PB_HD int tid() const { return 0; }
    PB_HD int size() const { return 1; }
    PB_HD int lane() const { return 0; }
    PB_HD int warp() const { return 0; }
    PB_HD int nwarps() const { return 1; }
    static constexpr int WARP = 1;
    PB_HD void sync() const {}
    PB_HD void warp_sync() const {}
    PB_HD double warp_sum(double v) const { return v; }
    PB_HD double sum16(double v) const { return v; }
    PB_HD int warp_max_int(int v) const { return v; }
    PB_HD double sum(double v) const { return v; }
    PB_HD double max(double v) const { return v; }
    PB_HD bool any(bool p) const { return p; }
    /// max of v over the team and the lowest idx attaining it
    PB_HD double argmax(double v, int& idx) const { (void)idx; return v; }
};

#if defined(__CUDACC__)
/// device: the thread block.  Reductions are fixed-order (shuffle tree inside a warp, then one
/// warp over the per-warp partials), so results are run-to-run reproducible.
struct CtaTeam {
    static constexpr int WARP = 32;
    /// shared scratch (the first 64 doubles of the kernel's dynamic shared memory), derived from
    /// the shared-memory base at every use so that the compiler emits LDS/STS
    __device__ __forceinline__ static double* red_base() {
        extern __shared__ __align__(16) double pb_team_smem[];
        return pb_team_smem;
    }
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
    __device__ __forceinline__ int nwarps() const { return blockDim.x >> 5; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    /// sum over the 16-lane half of the warp the lane belongs to (fixed shuffle tree)
    __device__ __forceinline__ double sum16(double v) const {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ int warp_max_int(int v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int u = __shfl_xor_sync(0xffffffffu, v, o);
            v = u > v ? u : v;
        }
        return v;
    }
    __device__ __forceinline__ double warp_max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __noinline__ double sum(double v) const {
        double* const red = red_base();
        v = warp_sum(v);
        __syncthreads();  // red may still be read by the previous reduction
        if (lane() == 0) red[warp()] = v;
        __syncthreads();
        double s = (lane() < nwarps()) ? red[lane()] : 0.0;
        return warp_sum(s);
    }
    __device__ __noinline__ double max(double v) const {
        double* const red = red_base();
        v = warp_max(v);
        __syncthreads();
        if (lane() == 0) red[warp()] = v;
        __syncthreads();
        double s = (lane() < nwarps()) ? red[lane()] : -INFINITY;
        return warp_max(s);
    }
    __device__ __forceinline__ bool any(bool p) const { return __syncthreads_or(p ? 1 : 0) != 0; }
    __device__ __noinline__ double argmax(double v, int& idx) const { return argmax_inline(v, idx); }
    /// the same, inlined (callers that hold many live registers: a call would spill them around it)
    __device__ __forceinline__ double argmax_inline(double v, int& idx) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double vo = __shfl_xor_sync(0xffffffffu, v, o);
            const int io = __shfl_xor_sync(0xffffffffu, idx, o);
            if (vo > v || (vo == v && io < idx)) { v = vo; idx = io; }
        }
        __syncthreads();
        double* const red = red_base();
        int* ired = reinterpret_cast<int*>(red + 32);
        if (lane() == 0) { red[warp()] = v; ired[warp()] = idx; }
        __syncthreads();
        v = (lane() < nwarps()) ? red[lane()] : -INFINITY;
        idx = (lane() < nwarps()) ? ired[lane()] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double vo = __shfl_xor_sync(0xffffffffu, v, o);
            const int io = __shfl_xor_sync(0xffffffffu, idx, o);
            if (vo > v || (vo == v && io < idx)) { v = vo; idx = io; }
        }
        return v;
    }
};
#endif

}  // namespace pb200
#endif
