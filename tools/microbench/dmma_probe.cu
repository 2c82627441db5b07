// dmma_probe.cu -- FP64 tensor-core (mma.sync.m8n8k4.f64) throughput against plain DFMA and
// against DFMA whose second operand is broadcast from shared memory, on the device it runs on.
// Decides how the large-model kernel's GEMM-shaped work (rank-NB update of the blocked
// Gauss-Jordan inversion, solve GEMM) is issued.  build: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int TILES>
__global__ void __launch_bounds__(512, 1) k_dmma(double* out, int iters) {
    double c[TILES][2];
#pragma unroll
    for (int t = 0; t < TILES; ++t) c[t][0] = c[t][1] = threadIdx.x * 1e-6 + t;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < TILES; ++t) dmma(c[t][0], c[t][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1];
    if (s == 123.456) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(512, 1) k_dfma(double* out, int iters) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fma(x[c], 0.999999, 1e-7);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
}

// every FMA takes one operand from shared memory (broadcast, 128-bit loads)
template <int CH>
__global__ void __launch_bounds__(512, 1) k_dfma_lds(double* out, int iters) {
    __shared__ __align__(16) double sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 1e-7 * threadIdx.x;
    __syncthreads();
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int it = 0; it < iters; ++it) {
        const double2* s2 = reinterpret_cast<const double2*>(sh + ((it & 1) * 32));
#pragma unroll
        for (int c = 0; c < CH; c += 2) {
            const double2 v = s2[c / 2];
            x[c] = fma(x[c], 0.999999, v.x);
            x[c + 1] = fma(x[c + 1], 0.999999, v.y);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 1 << 14;
    double* d; cudaMalloc(&d, 8);
    {
        double ms = time_ms([&] { k_dfma<16><<<sms, 512>>>(d, iters); });
        printf("DFMA registers only        : %.2f TFLOP/s\n", 2.0 * sms * 512 * 16 * iters / ms / 1e9);
    }
    {
        double ms = time_ms([&] { k_dfma_lds<16><<<sms, 512>>>(d, iters); });
        printf("DFMA, 1 operand LDS.128 bc : %.2f TFLOP/s\n", 2.0 * sms * 512 * 16 * iters / ms / 1e9);
    }
    {
        double ms = time_ms([&] { k_dmma<2><<<sms, 512>>>(d, iters); });
        printf("DMMA m8n8k4 (2 tiles/warp) : %.2f TFLOP/s\n", 2.0 * sms * 16 * 2 * 256.0 * iters / ms / 1e9);
    }
    {
        double ms = time_ms([&] { k_dmma<4><<<sms, 512>>>(d, iters); });
        printf("DMMA m8n8k4 (4 tiles/warp) : %.2f TFLOP/s\n", 2.0 * sms * 16 * 4 * 256.0 * iters / ms / 1e9);
    }
    {
        double ms = time_ms([&] { k_dmma<8><<<sms, 512>>>(d, iters); });
        printf("DMMA m8n8k4 (8 tiles/warp) : %.2f TFLOP/s\n", 2.0 * sms * 16 * 8 * 256.0 * iters / ms / 1e9);
    }
    {
        double ms = time_ms([&] { k_dmma<16><<<sms, 512>>>(d, iters); });
        printf("DMMA m8n8k4 (16 tiles/warp): %.2f TFLOP/s\n", 2.0 * sms * 16 * 16 * 256.0 * iters / ms / 1e9);
    }
    printf("cuda err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
