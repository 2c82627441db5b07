// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// Plain-loop stand-ins for the three CBLAS entry points AMICI's C++ runtime
// wraps (reference: deps/AMICI/src/cblas.cpp:21-48, interface
// deps/AMICI/include/amici/cblas.h:28-74).  No CBLAS library exists in this
// image, so the reference build under oracle/_ref links these instead of
// deps/AMICI/src/cblas.cpp.  Semantics are those of the reference BLAS:
// column- or row-major, optional transposes, general alpha/beta, strides.
#include "amici/cblas.h"

namespace amici {

namespace {
inline double at(double const* M, int ld, bool colMajor, bool trans, int r, int c) {
    // element (r,c) of op(M)
    if (trans) { int t = r; r = c; c = t; }
    return colMajor ? M[r + (long)c * ld] : M[(long)r * ld + c];
}
} // namespace

void amici_dgemm(BLASLayout layout, BLASTranspose TransA, BLASTranspose TransB,
                 int const M, int const N, int const K, double const alpha,
                 double const* A, int const lda, double const* B, int const ldb,
                 double const beta, double* C, int const ldc) {
    bool const cm = layout == BLASLayout::colMajor;
    bool const ta = TransA != BLASTranspose::noTrans;
    bool const tb = TransB != BLASTranspose::noTrans;
    for (int j = 0; j < N; ++j) {
        for (int i = 0; i < M; ++i) {
            double acc = 0.0;
            for (int k = 0; k < K; ++k)
                acc += at(A, lda, cm, ta, i, k) * at(B, ldb, cm, tb, k, j);
            double& c = cm ? C[i + (long)j * ldc] : C[(long)i * ldc + j];
            c = (beta == 0.0 ? 0.0 : beta * c) + alpha * acc;
        }
    }
}

void amici_dgemv(BLASLayout layout, BLASTranspose TransA, int const M, int const N,
                 double const alpha, double const* A, int const lda,
                 double const* X, int const incX, double const beta, double* Y,
                 int const incY) {
    bool const cm = layout == BLASLayout::colMajor;
    bool const ta = TransA != BLASTranspose::noTrans;
    int const rows = ta ? N : M;
    int const cols = ta ? M : N;
    for (int i = 0; i < rows; ++i) {
        double acc = 0.0;
        for (int k = 0; k < cols; ++k)
            acc += at(A, lda, cm, ta, i, k) * X[(long)k * incX];
        double& y = Y[(long)i * incY];
        y = (beta == 0.0 ? 0.0 : beta * y) + alpha * acc;
    }
}

void amici_daxpy(int n, double alpha, double const* x, int const incx, double* y,
                 int incy) {
    for (int i = 0; i < n; ++i)
        y[(long)i * incy] += alpha * x[(long)i * incx];
}

} // namespace amici
