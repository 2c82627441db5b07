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
// loops over the columns of a lane: never unrolled.  With one column per lane the trip count
// is 1 and the loop folds away; with several columns per lane (few lanes per condition) the
// body is emitted once and the columns live in local memory.
#define PB_ROLL _Pragma("unroll 1")
// loops over the state index in the hot path: partially unrolled only -- the kernels are
// instruction-fetch bound when warps on an SM run different conditions, so the static size of
// the step loop matters more than loop overhead (profiles/r01_b_summary.txt)
#ifndef PB_X_UNROLL
#define PB_X_UNROLL 2
#endif
#define PB_STR_(x) #x
#define PB_STR(x) PB_STR_(x)
#define PB_UNROLL_X _Pragma(PB_STR(unroll PB_X_UNROLL))
// loops over the columns in the row-parallel Newton kernels: a few columns in flight per lane
// (independent dependency chains), not the whole loop
#ifndef PB_COL_UNROLL
#define PB_COL_UNROLL 1
#endif
#define PB_UNROLL_COLS _Pragma(PB_STR(unroll PB_COL_UNROLL))
#else
#define PB_HD inline
#define PB_HD_NOINLINE
#define PB_UNROLL
#define PB_ROLL
#define PB_UNROLL_X
#define PB_UNROLL_COLS
#endif

// finite test that does not depend on which of <math.h> / <cmath> the translation unit saw first
#define PB_FINITE(v) (((v) - (v)) == 0.0)

// Out-of-line device functions receive the group's working set through a generic pointer; tell
// the compiler it is shared memory so that it emits LDS/STS (the kernels always place it there).
#if defined(__CUDA_ARCH__)
#define PB_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
// ... and that a pointer into the HBM workspace is global memory (LDG/STG instead of generic LD/ST)
#define PB_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define PB_ASSUME_SHARED(p) ((void)0)
#define PB_ASSUME_GLOBAL(p) ((void)0)
#endif

// Make a value opaque to the optimiser: stops it from cloning a large code region once per
// possible value of a flag (the kernels are instruction-cache sensitive).
#if defined(__CUDA_ARCH__)
#define PB_OPAQUE_INT(v) asm volatile("" : "+r"(v))
#define PB_OPAQUE_F64(v) asm volatile("" : "+d"(v))
#else
#define PB_OPAQUE_INT(v) asm volatile("" : "+r"(v))
#define PB_OPAQUE_F64(v) asm volatile("" : "+x"(v))
#endif

#endif
