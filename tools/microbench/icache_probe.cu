// I-cache capacity probe: K distinct code segments of ~SEGI instructions each, every warp walks the
// segments cyclically starting at a different one.  Prints cycles per instruction per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int S>
__device__ __forceinline__ void seg(float (&a)[8]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], 1.0f + 1e-7f * (S * 37 + i * 8 + j), 1e-9f * (S + j));
    }
}
template <int K>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cyc, int iters, int stagger) {
    float a[8];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    int s = stagger ? (warp * 7) % K : 0;
    long long t0 = clock64();
    for (int it = 0; it < iters * K; ++it) {
        switch (s) {
#define C(n) case n: if (n < K) seg<n>(a); break;
            C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
            C(16) C(17) C(18) C(19) C(20) C(21) C(22) C(23) C(24) C(25) C(26) C(27) C(28) C(29) C(30) C(31)
            C(32) C(33) C(34) C(35) C(36) C(37) C(38) C(39) C(40) C(41) C(42) C(43) C(44) C(45) C(46) C(47)
            C(48) C(49) C(50) C(51) C(52) C(53) C(54) C(55) C(56) C(57) C(58) C(59) C(60) C(61) C(62) C(63)
        }
        s = (s + 1 == K) ? 0 : s + 1;
    }
    long long t1 = clock64();
    float r = 0;
    for (int j = 0; j < 8; ++j) r += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x % 32 == 0) cyc[warp] = t1 - t0;
}
template <int K>
void run(int warps_per_sm, int stagger) {
    int sms = 148;
    float* out; long long* cyc;
    int threads = warps_per_sm * 32;
    cudaMalloc(&out, sms * threads * sizeof(float));
    cudaMalloc(&cyc, sms * warps_per_sm * sizeof(long long));
    int iters = 64;
    probe<K><<<sms, threads>>>(out, cyc, 4, stagger);
    cudaDeviceSynchronize();
    probe<K><<<sms, threads>>>(out, cyc, iters, stagger);
    cudaDeviceSynchronize();
    static long long h[148 * 64];
    cudaMemcpy(h, cyc, sms * warps_per_sm * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < sms * warps_per_sm; ++i) mean += h[i]; mean /= sms * warps_per_sm;
    double instr = (double)iters * K * 256.0;
    printf("K=%2d code=%4d KB warps/SM=%2d stagger=%d: %.2f cycles/instr/warp  (SM IPC %.2f)\n", K, K * 4, warps_per_sm, stagger,
           mean / instr, warps_per_sm * instr / mean);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int st = 0; st <= 1; ++st)
        for (int w : {4, 8, 16, 24}) {
            run<2>(w, st); run<4>(w, st); run<6>(w, st); run<8>(w, st); run<12>(w, st); run<16>(w, st); run<20>(w, st); run<24>(w, st); run<32>(w, st); run<48>(w, st); run<64>(w, st);
        }
    return 0;
}
