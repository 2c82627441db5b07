// Host/device portability macros.  The integrator core (bdf_core.cuh) and the
// generated model functions compile both as sm_100a device code (the product)
// and as plain host C++ (tests/host_harness only: lets the device algorithm be
// stepped and diffed against the CVODES oracle without a GPU; the product library
// never contains or calls the host instantiation).
#ifndef PB200_PB_HD_H
#define PB200_PB_HD_H

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_HD_NOINLINE __host__ __device__ __noinline__
#define PB_UNROLL _Pragma("unroll")
#else
#define PB_HD inline
#define PB_HD_NOINLINE
#define PB_UNROLL
#endif

#endif
