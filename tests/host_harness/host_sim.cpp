// TEST INFRASTRUCTURE -- host instantiation of the device integrator core.
//
// Compiles parpe_b200/csrc/bdf_core.cuh and the generated model functions as plain
// C++ with the SingleLane group, so the algorithm that runs on the GPU can be
// stepped on the CPU and diffed against the CVODES oracle (tests/test_host_core.py).
// It is never linked into, or called by, the product library (libparpe_b200.so).
#define PB200_HOST_TRACE 1
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../parpe_b200/csrc/bdf_core.cuh"
#include "../../parpe_b200/csrc/models/boehm.cuh"
#include "../../parpe_b200/csrc/models/fujita.cuh"
#include "../../parpe_b200/csrc/models/steadystate.cuh"
#include "../../parpe_b200/csrc/models/steadystate_scaled.cuh"
// the reference's own AMICI-generated model, re-homed by parpe_b200/codegen/amici_import.py at build time (git-ignored)
#if __has_include("../../parpe_b200/csrc/models/imported/amici_steadystate.cuh")
#include "../../parpe_b200/csrc/models/imported/amici_steadystate.cuh"
#define PB200_HAVE_AMICI_STEADYSTATE 1
#endif
#if __has_include("../../parpe_b200/csrc/models/imported/amici_jakstat_adjoint.cuh")
#include "../../parpe_b200/csrc/models/imported/amici_jakstat_adjoint.cuh"
#define PB200_HAVE_AMICI_JAKSTAT 1
#endif
// models written by AMICI's python code generator at build time (parpe_b200/codegen/amici_pygen.py), re-homed likewise
#if __has_include("../../parpe_b200/csrc/models/imported/pygen_steadystate.cuh")
#include "../../parpe_b200/csrc/models/imported/pygen_steadystate.cuh"
#define PB200_HAVE_PYGEN_STEADYSTATE 1
#endif
#if __has_include("../../parpe_b200/csrc/models/imported/pygen_boehm_w.cuh")
#include "../../parpe_b200/csrc/models/imported/pygen_boehm_w.cuh"
#define PB200_HAVE_PYGEN_BOEHM_W 1
#endif

using namespace pb200;

template <class M>
static int run(const double* p_unscaled, const double* k, const int* plist, int ns, int nt, const double* t,
               const double* my, const double* sd, double t0, double rtol, double atol, int maxsteps, int stldet,
               int errconS, double* llh, double* sllh, int64_t* stats, double* y_out, double* sigmay_out, int* trace) {
    static thread_local GroupShared<M> sh;
    static thread_local Bdf<M, SingleLane> bdf;
    bdf.sh = &sh;
    for (int i = 0; i < M::NP; ++i) sh.p[i] = p_unscaled[i];
    for (int i = 0; i < M::NK; ++i) sh.k[i] = k[i];
    for (int i = 0; i < ns; ++i) sh.plist[i] = plist[i];
    bdf.opt.t0 = t0;
    bdf.opt.maxsteps = maxsteps;
    bdf.opt.stldet = stldet;
    bdf.opt.errconS = errconS;
    bdf.opt.newton_maxsteps = 0;
    ConditionIO io{nt, t, my, sd, ns, y_out, sigmay_out, trace};
    int status = bdf.simulate(io, rtol, atol, *llh);
    for (int j = 0; j < ns; ++j) sllh[j] = sh.sllh[j + 1];
    bdf.export_stats(stats, 1);
    return status;
}

/// The steady-state instantiation (Bdf<M, Grp, true>): t = inf output times and, with k_preeq != nullptr,
/// a pre-equilibration phase followed by the re-initialisation of the states in reinit_mask -- the
/// sequence simulate_kernel's steady-state variant runs per condition.
template <class M>
static int run_ss(const double* p_unscaled, const double* k, const double* k_preeq, unsigned reinit_mask,
                  const int* plist, int ns, int nt, const double* t, const double* my, const double* sd, double t0,
                  double rtol, double atol, int maxsteps, int stldet, int errconS, double* llh, double* sllh,
                  int64_t* stats, double* y_out, double* sigmay_out, int* trace, double* x_preeq, int newton_maxsteps) {
    static thread_local GroupShared<M> sh;
    static thread_local Bdf<M, SingleLane, true> bdf;
    using B = Bdf<M, SingleLane, true>;
    bdf.sh = &sh;
    for (int i = 0; i < M::NP; ++i) sh.p[i] = p_unscaled[i];
    for (int i = 0; i < ns; ++i) sh.plist[i] = plist[i];
    bdf.opt.t0 = t0;
    bdf.opt.maxsteps = maxsteps;
    bdf.opt.stldet = stldet;
    bdf.opt.errconS = errconS;
    bdf.opt.newton_maxsteps = newton_maxsteps;
    ConditionIO io{nt, t, my, sd, ns, y_out, sigmay_out, trace};
    int status;
    if (k_preeq) {
        for (int i = 0; i < M::NK; ++i) sh.k[i] = k_preeq[i];
        bdf.begin(io, rtol, atol, B::SS_PREEQ);
        while (!bdf.advance()) {}
        status = bdf.status();
        stats[16] = bdf.nst;  // steps of the pre-equilibration
        if (x_preeq)
            for (int i = 0; i < M::NX * bdf.ncol; ++i) x_preeq[i] = sh.zn[0][i];
        if (status == SIM_SUCCESS) {
            for (int i = 0; i < M::NK; ++i) sh.k[i] = k[i];
            bdf.begin(io, rtol, atol, B::SS_POSTEQ, reinit_mask);
            while (!bdf.advance()) {}
            status = bdf.status();
        }
    } else {
        for (int i = 0; i < M::NK; ++i) sh.k[i] = k[i];
        bdf.begin(io, rtol, atol);
        while (!bdf.advance()) {}
        status = bdf.status();
    }
    *llh = bdf.llh();
    for (int j = 0; j < ns; ++j) sllh[j] = sh.sllh[j + 1];
    const int64_t npre = stats[16];
    bdf.export_stats(stats, 1);
    stats[16] = k_preeq ? npre : 0;
    return status;
}

extern "C" int hs_simulate_ss(const char* model, const double* p_unscaled, const double* k, const double* k_preeq,
                              unsigned reinit_mask, const int* plist, int ns, int nt, const double* t,
                              const double* my, const double* sd, double t0, double rtol, double atol, int maxsteps,
                              int stldet, int errconS, double* llh, double* sllh, int64_t* stats, double* y_out,
                              double* sigmay_out, int* trace, double* x_preeq, int newton_maxsteps) {
#define DISPATCH_SS(NAME)                                                                                          \
    if (!std::strcmp(model, #NAME))                                                                                \
        return run_ss<models::Model_##NAME>(p_unscaled, k, k_preeq, reinit_mask, plist, ns, nt, t, my, sd, t0, rtol, \
                                            atol, maxsteps, stldet, errconS, llh, sllh, stats, y_out, sigmay_out,  \
                                            trace, x_preeq, newton_maxsteps);
    DISPATCH_SS(steadystate)
    DISPATCH_SS(steadystate_scaled)
    DISPATCH_SS(boehm)
    DISPATCH_SS(fujita)
#if defined(PB200_HAVE_AMICI_STEADYSTATE)
    DISPATCH_SS(amici_steadystate)
#endif
#if defined(PB200_HAVE_AMICI_JAKSTAT)
    DISPATCH_SS(amici_jakstat_adjoint)
#endif
#if defined(PB200_HAVE_PYGEN_STEADYSTATE)
    DISPATCH_SS(pygen_steadystate)
#endif
#if defined(PB200_HAVE_PYGEN_BOEHM_W)
    DISPATCH_SS(pygen_boehm_w)
#endif
    return -12345;
}

extern "C" int hs_simulate(const char* model, const double* p_unscaled, const double* k, const int* plist, int ns,
                           int nt, const double* t, const double* my, const double* sd, double t0, double rtol,
                           double atol, int maxsteps, int stldet, int errconS, double* llh, double* sllh,
                           int64_t* stats, double* y_out, double* sigmay_out, int* trace) {
#define DISPATCH(NAME)                                                                                      \
    if (!std::strcmp(model, #NAME))                                                                         \
        return run<models::Model_##NAME>(p_unscaled, k, plist, ns, nt, t, my, sd, t0, rtol, atol, maxsteps, \
                                         stldet, errconS, llh, sllh, stats, y_out, sigmay_out, trace);
    DISPATCH(steadystate)
    DISPATCH(steadystate_scaled)
    DISPATCH(boehm)
    DISPATCH(fujita)
#if defined(PB200_HAVE_AMICI_STEADYSTATE)
    DISPATCH(amici_steadystate)
#endif
#if defined(PB200_HAVE_AMICI_JAKSTAT)
    DISPATCH(amici_jakstat_adjoint)
#endif
#if defined(PB200_HAVE_PYGEN_STEADYSTATE)
    DISPATCH(pygen_steadystate)
#endif
#if defined(PB200_HAVE_PYGEN_BOEHM_W)
    DISPATCH(pygen_boehm_w)
#endif
    return -12345;
}
