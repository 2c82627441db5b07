// Device counterparts of the helper functions AMICI's generated model code calls (namespace amici::, declared in
// deps/AMICI/include/amici/symbolic_functions.h): the legacy spline inputs of the classic ("MATLAB-generated")
// layout,
//     amici::spline(t, n, t1, u1, ..., tn, un, ss, dudt)          cubic spline through (ti, ui)
//     amici::spline_pos(t, n, ...)                                 exp of the cubic spline through (ti, log ui)
//     amici::Dspline(id, t, n, ...), amici::Dspline_pos(id, t, n, ...)   derivative w.r.t. the node value that is
//                                                                  argument number `id` (1-based) of the value call
// (deps/AMICI/src/symbolic_functions.cpp:109-269; a vendored model that uses them: deps/AMICI/models/
// model_jakstat_adjoint/w.cpp:14).  The reference's functions are C varargs over the interpolation routine of
// Forsythe, Malcolm & Moler, "Computer Methods for Mathematical Computations" (1977) (deps/AMICI/src/spline.cpp:16-257);
// CUDA has no varargs in device code, so these are variadic templates whose node count is a compile-time constant, and
// the interpolant is written here from the book's formulation: a cubic spline whose end conditions match the third
// derivative of the cubic through the first / last four nodes, or -- first node only, ss != 0 -- a prescribed slope.
// Host and device (the host harness and the oracle comparison run the same code).
#ifndef PB200_AMICI_FUNCTIONS_CUH
#define PB200_AMICI_FUNCTIONS_CUH
#include <math.h>

#include "pb_hd.h"

namespace pb200 {
namespace amifn {

/// Piecewise cubic s(t) = fk[i] + w (c1[i] + w (c2[i] + w c3[i])), w = t - tk[i], on [tk[i], tk[i+1]) -- knots tk, values fk,
/// coefficients c1..c3; the first / last piece is continued outside the knots (as the reference's seval does).
template <int N>
struct Cubic {
    static_assert(N >= 2, "a spline needs two nodes");
    double tk[N], fk[N], c1[N], c2[N], c3[N];

    /// Coefficients from the nodes: tridiagonal system for the scaled second derivatives sigma_i (diagonal in c1,
    /// off-diagonal in c3, right-hand side in c2), eliminated in place.
    PB_HD void fit(bool slope_at_first, double slope_first) {
        if (N == 2) {
            const double s = (fk[1] - fk[0]) / (tk[1] - tk[0]);
            for (int i = 0; i < N; ++i) { c1[i] = s; c2[i] = 0.0; c3[i] = 0.0; }
            return;
        }
        constexpr int L = N - 1;  // last node
        c3[0] = tk[1] - tk[0];
        c2[1] = (fk[1] - fk[0]) / c3[0];
        for (int i = 1; i < L; ++i) {
            c3[i] = tk[i + 1] - tk[i];
            c1[i] = 2.0 * (c3[i - 1] + c3[i]);
            c2[i + 1] = (fk[i + 1] - fk[i]) / c3[i];
            c2[i] = c2[i + 1] - c2[i];
        }
        // end conditions: third derivative from divided differences over the outer four nodes
        c1[0] = -c3[0];
        c1[L] = -c3[L - 1];
        c2[0] = 0.0;
        c2[L] = 0.0;
        if (N > 3) {
            c2[0] = c2[2] / (tk[3] - tk[1]) - c2[1] / (tk[2] - tk[0]);
            c2[L] = c2[L - 1] / (tk[L] - tk[L - 2]) - c2[L - 2] / (tk[L - 1] - tk[L - 3]);
            c2[0] = c2[0] * c3[0] * c3[0] / (tk[3] - tk[0]);
            c2[L] = -c2[L] * c3[L - 1] * c3[L - 1] / (tk[L] - tk[L - 3]);
        }
        if (slope_at_first) {  // ... or the prescribed slope at the first node
            c1[0] = 2.0 * (tk[1] - tk[0]);
            c2[0] = (fk[1] - fk[0]) / (tk[1] - tk[0]) - slope_first;
        }
        for (int i = 1; i < N; ++i) {  // elimination
            const double f = c3[i - 1] / c1[i - 1];
            c1[i] -= f * c3[i - 1];
            c2[i] -= f * c2[i - 1];
        }
        c2[L] = c2[L] / c1[L];  // back substitution: c = sigma
        for (int i = L - 1; i >= 0; --i) c2[i] = (c2[i] - c3[i] * c2[i + 1]) / c1[i];
        // polynomial coefficients of the pieces
        c1[L] = (fk[L] - fk[L - 1]) / c3[L - 1] + c3[L - 1] * (c2[L - 1] + 2.0 * c2[L]);
        for (int i = 0; i < L; ++i) {
            c1[i] = (fk[i + 1] - fk[i]) / c3[i] - c3[i] * (c2[i + 1] + 2.0 * c2[i]);
            c3[i] = (c2[i + 1] - c2[i]) / c3[i];
            c2[i] = 3.0 * c2[i];
        }
        c2[L] = 3.0 * c2[L];
        c3[L] = c3[L - 1];
    }

    PB_HD double operator()(double t) const {
        int i = 0;
        if (t >= tk[N - 1]) i = N - 1;
        else
            for (int j = 1; j < N - 1; ++j)
                if (t >= tk[j]) i = j;
        const double w = t - tk[i];
        return fk[i] + w * (c1[i] + w * (c2[i] + w * c3[i]));
    }
};

namespace detail {
/// (t1, u1, ..., tn, un, ss, dudt) of a generated call, as doubles
template <int M, class... A>
PB_HD void unpack(double (&v)[M], A... args) {
    const double tmp[M] = {static_cast<double>(args)...};
    for (int i = 0; i < M; ++i) v[i] = tmp[i];
}
enum Kind { VALUE, VALUE_POS, DVALUE, DVALUE_POS };

template <int KIND, class... A>
PB_HD double eval(int id, double t, A... args) {
    constexpr int M = (int)sizeof...(A);
    static_assert(M >= 6 && M % 2 == 0, "spline call: n nodes as (t_i, u_i) pairs, then ss and dudt");
    constexpr int N = (M - 2) / 2;
    double v[M];
    unpack<M>(v, args...);
    const bool clamp = static_cast<int>(v[M - 2]) == 1;
    const double slope = v[M - 1];
    Cubic<N> s;
    for (int i = 0; i < N; ++i) s.tk[i] = v[2 * i];
    if (KIND == VALUE) {
        for (int i = 0; i < N; ++i) s.fk[i] = v[2 * i + 1];
        s.fit(clamp, slope);
        return s(t);
    }
    if (KIND == VALUE_POS || KIND == DVALUE_POS) {
        for (int i = 0; i < N; ++i) s.fk[i] = log(v[2 * i + 1]);
        s.fit(clamp, slope);
    }
    if (KIND == VALUE_POS) return exp(s(t));
    // derivative w.r.t. the node value that is argument number `id` (1-based) of the VALUE call (t, n, t1, u1, ...):
    // u_j is argument 4 + 2j.  The spline is linear in its node values, so d s / d u_j is the spline through the
    // unit vector e_j -- with the same end-condition arguments, as the reference passes them
    const int j = id / 2 - 2;
    const double pos = (KIND == DVALUE_POS) ? exp(s(t)) : 1.0;
    for (int i = 0; i < N; ++i) s.fk[i] = (i == j) ? 1.0 : 0.0;
    s.fit(clamp, slope);
    const double ds = s(t);
    // positive spline: d exp(S(log u)) / d u_j = exp(S) * dS/d(log u_j) / u_j
    return (KIND == DVALUE_POS) ? ds * pos / v[2 * j + 1] : ds;
}
}  // namespace detail

template <class... A>
PB_HD double spline(double t, int /*num*/, A... args) { return detail::eval<detail::VALUE>(0, t, args...); }
template <class... A>
PB_HD double spline_pos(double t, int /*num*/, A... args) { return detail::eval<detail::VALUE_POS>(0, t, args...); }
template <class... A>
PB_HD double Dspline(int id, double t, int /*num*/, A... args) { return detail::eval<detail::DVALUE>(id, t, args...); }
template <class... A>
PB_HD double Dspline_pos(int id, double t, int /*num*/, A... args) { return detail::eval<detail::DVALUE_POS>(id, t, args...); }

}  // namespace amifn
}  // namespace pb200
#endif
