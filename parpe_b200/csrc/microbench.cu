// microbench.cu -- measured FP64 roofline denominator.
//
// /root/repo/MEASURED_PEAKS.json carries HBM bandwidth and bf16 tensor throughput but no
// FP64 figure; the integrator is FP64-issue bound for small models, so bench.py measures
// the DFMA peak of the device it runs on with this kernel (independent FMA chains, no
// memory traffic) and reports roofline fractions "of measured".
#include <cuda_runtime.h>

#include "../../include/parpe_b200.h"

namespace {
template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;  // keep the chains alive
}
}  // namespace

extern "C" int pb200_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return PB200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PB200_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    constexpr int CHAINS = 16;
    // (long enough that launch ramp and tail are below 1 %: ~35 ms per repetition; the same loop as
    // tools/microbench/dmma_probe.cu's "DFMA registers only" line, which reads 36.6 TFLOP/s on this pool)
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 17;
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return PB200_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<CHAINS><<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return PB200_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * (double)blocks * threads * CHAINS * iters;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return 0;
}
