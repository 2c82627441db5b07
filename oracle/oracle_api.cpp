// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// C API over the oracle (reference AMICI/CVODES + restated parPE glue) so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can drive it
// through ctypes with the same tables they hand to the CUDA library.
#include <cstring>
#include <map>
#include <string>

#include "../include/parpe_b200.h"
#include "hierarchical.hpp"
#include "parpe_glue.hpp"

#include "models/model_boehm.h"
#include "models/model_fujita.h"
#include "models/model_steadystate.h"
#include "models/model_steadystate_scaled.h"
#include "net_model.hpp"
#include "../parpe_b200/csrc/models/net_chen.h"
#include "../parpe_b200/csrc/models/net_froehlich.h"
#include "../parpe_b200/csrc/models/net_net31.h"
#include "../parpe_b200/csrc/models/net_net32.h"

#ifdef PB200_ORACLE_WITH_VENDORED
// the reference's own pre-generated model (deps/AMICI/models/model_steadystate),
// compiled in place by the Makefile: cross-checks our generator
#include "model_steadystate.h"
// ... and its model with a spline input (deps/AMICI/models/model_jakstat_adjoint)
#include "model_jakstat_adjoint.h"
#endif

#ifdef PB200_ORACLE_WITH_PYGEN
// models written by AMICI's python code generator at build time (parpe_b200/codegen/amici_pygen.py)
std::unique_ptr<amici::Model> pb200_pygen_model(const std::string& name);
#endif

using namespace pb200_oracle;

namespace {
std::unique_ptr<amici::Model> makeModel(const std::string& name) {
    using namespace pb200_oracle_models;
    if (name == "steadystate") return create_steadystate();
    if (name == "steadystate_scaled") return create_steadystate_scaled();
    if (name == "boehm") return create_boehm();
    if (name == "fujita") return create_fujita();
    if (name == "net32") return std::make_unique<NetworkModel>(pb200::nets::net32::tables());
    if (name == "net31") return std::make_unique<NetworkModel>(pb200::nets::net31::tables());
    if (name == "chen") return std::make_unique<NetworkModel>(pb200::nets::chen::tables());
    if (name == "froehlich") return std::make_unique<NetworkModel>(pb200::nets::froehlich::tables());
#ifdef PB200_ORACLE_WITH_VENDORED
    if (name == "steadystate_vendored")
        return std::make_unique<amici::model_model_steadystate::Model_model_steadystate>();
    if (name == "jakstat_vendored")
        return std::make_unique<amici::model_model_jakstat_adjoint::Model_model_jakstat_adjoint>();
#endif
#ifdef PB200_ORACLE_WITH_PYGEN
    if (name.rfind("pygen_", 0) == 0) return pb200_pygen_model(name);
#endif
    return nullptr;
}
thread_local std::string g_err;
}  // namespace

struct ora_problem {
    std::unique_ptr<Problem> p;
    std::unique_ptr<HierarchicalSpec> hier;
};

extern "C" {

const char* ora_last_error() { return g_err.c_str(); }

int ora_has_model(const char* name) { return makeModel(name) ? 1 : 0; }

// ism: 1 simultaneous / 2 staggered; linsol: 1 dense / 9 KLU
ora_problem* ora_create(const pb200_problem_desc* d, int ism, int linsol) {
    try {
        auto model = makeModel(d->model_name);
        if (!model) {
            g_err = std::string("unknown model ") + d->model_name;
            return nullptr;
        }
        Tables t;
        t.n_conditions = d->n_conditions;
        t.n_opt = d->n_opt;
        t.np = d->np;
        t.nk = d->nk;
        t.ny = d->ny;
        const int C = d->n_conditions;
        t.t_offsets.assign(d->t_offsets, d->t_offsets + C + 1);
        int64_t nt_total = t.t_offsets[C];
        t.t.assign(d->t, d->t + nt_total);
        t.y.assign(d->y, d->y + nt_total * d->ny);
        t.sigma.assign(d->sigma, d->sigma + nt_total * d->ny);
        t.k.assign(d->k, d->k + (size_t)C * d->nk);
        t.mapping.assign(d->mapping, d->mapping + (size_t)C * d->np);
        t.pscale_sim.assign(d->pscale_sim, d->pscale_sim + (size_t)C * d->np);
        t.pscale_opt.assign(d->pscale_opt, d->pscale_opt + d->n_opt);
        if (d->overrides)
            t.overrides.assign(d->overrides, d->overrides + (size_t)C * d->np);
        t.t0 = d->t0;
        t.atol = d->atol;
        t.rtol = d->rtol;
        t.maxsteps = d->maxsteps;
        t.stldet = d->stldet;
        t.num_trials = d->num_trials;
        t.relaxation_factor = d->relaxation_factor;
        t.newton_maxsteps = d->newton_maxsteps;
        t.newton_only = d->newton_only;
        t.steady_sens_newton = d->steady_sens_newton;
        t.ism = d->ism == 1 ? 1 : ism;  // (the descriptor's request for the simultaneous corrector wins)
        t.linsol = linsol;
        model->setT0(d->t0);
        auto* h = new ora_problem;
        h->p = std::make_unique<Problem>(std::move(model), std::move(t));
        return h;
    } catch (std::exception const& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ora_destroy(ora_problem* h) { delete h; }

// counterpart of pb200_set_preequilibration (include/parpe_b200.h)
int ora_set_preequilibration(ora_problem* h, const int32_t* has_preeq, const double* k_preeq,
                             const int64_t* reinit_offsets, const int32_t* reinit_idx) {
    Tables& t = h->p->mutableTables();
    t.has_preeq.clear();
    t.k_preeq.clear();
    t.reinit_idx.clear();
    if (!has_preeq)
        return 0;
    const int C = t.n_conditions;
    t.has_preeq.assign(has_preeq, has_preeq + C);
    t.k_preeq.assign(k_preeq, k_preeq + (size_t)C * t.nk);
    t.reinit_idx.resize(C);
    if (reinit_offsets)
        for (int c = 0; c < C; ++c)
            t.reinit_idx[c].assign(reinit_idx + reinit_offsets[c], reinit_idx + reinit_offsets[c + 1]);
    return 0;
}

// returns 0/1 (FunctionEvaluationStatus) or -1 on exception
int ora_evaluate(ora_problem* h, const double* theta, const int* idx, int n_idx, double* fval,
                 double* grad, int nthreads, int collect_stats, double* llh_c, double* sllh_c,
                 int* status_c, int64_t* stats_c, double* seconds_c) {
    try {
        std::vector<int> vidx(idx, idx + n_idx);
        std::vector<ConditionResult> res;
        int rc = h->p->evaluate(theta, vidx, *fval, grad, nthreads, &res, collect_stats != 0);
        const int np = h->p->tables().np;
        for (int r = 0; r < n_idx; ++r) {
            if (llh_c) llh_c[r] = res[r].llh;
            if (status_c) status_c[r] = res[r].status;
            if (seconds_c) seconds_c[r] = res[r].seconds;
            if (sllh_c) {
                for (int j = 0; j < np; ++j)
                    sllh_c[(size_t)r * np + j] = j < (int)res[r].gradient.size() ? res[r].gradient[j] : 0.0;
            }
            if (stats_c) {
                int64_t* s = stats_c + (size_t)r * PB200_NSTATS;
                for (int j = 0; j < PB200_NSTATS; ++j) s[j] = -1;
                s[PB200_STAT_NST] = res[r].nst;
                s[PB200_STAT_NFE] = res[r].nfe;
                s[PB200_STAT_NETF] = res[r].netf;
                s[PB200_STAT_NCFN] = res[r].ncfn;
                s[PB200_STAT_TRIALS] = res[r].trials;
                s[14] = res[r].preeq_steps;   // steps of the pre-equilibration simulation
                s[15] = res[r].posteq_steps;  // steps of the post-equilibration simulation
            }
        }
        return rc;
    } catch (std::exception const& e) {
        g_err = e.what();
        return -1;
    }
}

// y_out / sigmay_out in the layout of desc->y
int ora_get_model_outputs(ora_problem* h, const double* theta, double* y_out, double* sigmay_out,
                          int nthreads) {
    try {
        std::vector<std::vector<double>> y, s;
        int rc = h->p->getModelOutputsAndSigmas(theta, y, s, nthreads);
        auto const& tab = h->p->tables();
        for (int c = 0; c < tab.n_conditions; ++c) {
            size_t off = (size_t)tab.t_offsets[c] * tab.ny;
            size_t n = (size_t)(tab.t_offsets[c + 1] - tab.t_offsets[c]) * tab.ny;
            for (size_t i = 0; i < n; ++i) {
                y_out[off + i] = i < y[c].size() ? y[c][i] : NAN;
                sigmay_out[off + i] = i < s[c].size() ? s[c][i] : NAN;
            }
        }
        return rc;
    } catch (std::exception const& e) {
        g_err = e.what();
        return -1;
    }
}

// debugging aid: cumulative step count and order at every output time of condition c
int ora_debug_numsteps(ora_problem* h, const double* theta, int c, int want_grad, int* numsteps, int* order) {
    try {
        std::vector<int> vidx{c};
        std::vector<ConditionResult> res;
        double fval;
        std::vector<double> grad(h->p->tables().n_opt);
        h->p->evaluate(theta, vidx, fval, want_grad ? grad.data() : nullptr, 1, &res, true);
        for (size_t i = 0; i < res[0].numsteps.size(); ++i) {
            numsteps[i] = res[0].numsteps[i];
            order[i] = res[0].order[i];
        }
        return (int)res[0].numsteps.size();
    } catch (std::exception const& e) {
        g_err = e.what();
        return -1;
    }
}

// A5 helpers exposed for the bit-exact indexing tests
void ora_map_parameters(ora_problem* h, int c, const double* theta, double* p_sim) {
    std::vector<double> p;
    h->p->mapAndSetOptimizationToSimulationVariables(c, theta, p);
    std::memcpy(p_sim, p.data(), p.size() * sizeof(double));
}
int ora_plist(ora_problem* h, int c, int* plist) {
    auto pl = h->p->plistFor(h->p->mappingFor(c));
    std::memcpy(plist, pl.data(), pl.size() * sizeof(int));
    return (int)pl.size();
}
double ora_get_unscaled(double v, int scale) { return getUnscaledParameter(v, scale); }
double ora_get_scaled(double v, int scale) { return getScaledParameter(v, scale); }
double ora_apply_chain_rule(double g, double p, int oldScale, int newScale) {
    return applyChainRule(g, p, oldScale, newScale);
}

// ---- hierarchical free functions on caller-supplied outputs (known-answer tests) ----------
// outputs / measurements: [n_conditions][nt_c * ny] flattened with offsets[n_conditions+1]
namespace {
Matrix toMatrix(const double* flat, const int64_t* off, int C) {
    Matrix m(C);
    for (int c = 0; c < C; ++c) m[c].assign(flat + off[c], flat + off[c + 1]);
    return m;
}
AnalyticalParameterProvider provider(int n_param, const int32_t* rows, int n_rows) {
    std::vector<int32_t> idx(n_param);
    for (int i = 0; i < n_param; ++i) idx[i] = i;
    return AnalyticalParameterProvider::fromRows(idx.data(), n_param, rows, n_rows);
}
}  // namespace

// kind: 0 scaling, 1 offset, 2 sigma
double ora_hier_compute_one(int kind, int param_idx, int n_param, const int32_t* rows, int n_rows, int C,
                            const int64_t* off, const double* outputs, const double* measurements, int ny) {
    auto prov = provider(n_param, rows, n_rows);
    auto out = toMatrix(outputs, off, C), mes = toMatrix(measurements, off, C);
    if (kind == 0) return computeAnalyticalScalings(param_idx, out, mes, prov, ny);
    if (kind == 1) return computeAnalyticalOffsets(param_idx, out, mes, prov, ny);
    return computeAnalyticalSigmas(param_idx, out, mes, prov, ny);
}
// mode: 0 scaling (multiply), 1 offset (add); outputs modified in place
void ora_hier_apply_one(int mode, int param_idx, double value, int n_param, const int32_t* rows, int n_rows, int C,
                        const int64_t* off, double* outputs, int ny) {
    auto prov = provider(n_param, rows, n_rows);
    auto out = toMatrix(outputs, off, C);
    if (mode == 0) applyOptimalScaling(param_idx, value, out, prov, ny);
    else applyOptimalOffset(param_idx, value, out, prov, ny);
    for (int c = 0; c < C; ++c) std::memcpy(outputs + off[c], out[c].data(), out[c].size() * sizeof(double));
}
double ora_hier_nllh(int C, const int64_t* off, const double* measurements, const double* outputs,
                     const double* sigmas) {
    return computeNegLogLikelihood(toMatrix(measurements, off, C), toMatrix(outputs, off, C), toMatrix(sigmas, off, C));
}
void ora_hier_splice(const double* reduced, int n_reduced, const int32_t* si, int ns, const int32_t* oi, int no,
                     const int32_t* gi, int ng, const double* sv, const double* ov, const double* gv, double* full) {
    auto f = spliceParameters(std::vector<double>(reduced, reduced + n_reduced), std::vector<int>(si, si + ns),
                              std::vector<int>(oi, oi + no), std::vector<int>(gi, gi + ng),
                              std::vector<double>(sv, sv + ns), std::vector<double>(ov, ov + no),
                              std::vector<double>(gv, gv + ng));
    std::memcpy(full, f.data(), f.size() * sizeof(double));
}
void ora_hier_filter(const double* values, int n, const int32_t* excl, int n_excl, double* result) {
    fillFilteredParams(std::vector<double>(values, values + n), std::vector<int>(excl, excl + n_excl), result);
}
double ora_hier_default_scaling(int scale) { return getDefaultScalingFactor(scale); }
double ora_hier_default_offset(int scale) { return getDefaultOffsetParameter(scale); }

// ---- hierarchical (A15) --------------------------------------------------
int ora_hier_attach(ora_problem* h, const pb200_hier_desc* hd) {
    try {
        h->hier = std::make_unique<HierarchicalSpec>(HierarchicalSpec::fromDesc(*hd, h->p->tables()));
        return 0;
    } catch (std::exception const& e) {
        g_err = e.what();
        return -1;
    }
}
int ora_hier_num_outer(ora_problem* h) { return h->hier ? h->hier->numOuter() : -1; }
int ora_hier_evaluate(ora_problem* h, const double* theta_outer, double* fval, double* grad_outer,
                      double* full_parameters, int nthreads) {
    try {
        return hierarchicalEvaluate(*h->p, *h->hier, theta_outer, *fval, grad_outer, full_parameters, nthreads);
    } catch (std::exception const& e) {
        g_err = e.what();
        return -1;
    }
}
}  // extern "C"
