// TEST INFRASTRUCTURE (oracle) -- not part of the product path.  See parpe_glue.hpp.
#include "parpe_glue.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <numeric>
#include <stdexcept>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace pb200_oracle {

// deps/AMICI/src/misc.cpp:26-37
double getUnscaledParameter(double v, int scale) {
    switch (scale) {
    case 2: return std::pow(10, v);
    case 1: return std::exp(v);
    case 0: return v;
    }
    throw std::invalid_argument("Invalid value for ParameterScaling.");
}

// deps/AMICI/src/misc.cpp:52-63
double getScaledParameter(double v, int scale) {
    switch (scale) {
    case 2: return std::log10(v);
    case 1: return std::log(v);
    case 0: return v;
    }
    throw std::invalid_argument("Invalid value for ParameterScaling.");
}

// src/parpeamici/multiConditionDataProvider.cpp:762-796 (quirk Q4 kept: the
// parameter is unscaled with the OLD scale in both branches)
double applyChainRule(double gradient, double parameter, int oldScale, int newScale) {
    if (oldScale == newScale)
        return gradient;
    switch (oldScale) {
    case 2: gradient /= getUnscaledParameter(parameter, oldScale) * std::log(10); break;
    case 1: gradient /= getUnscaledParameter(parameter, oldScale); break;
    default: break;
    }
    switch (newScale) {
    case 2: gradient *= getUnscaledParameter(parameter, oldScale) * std::log(10); break;
    case 1: gradient *= getUnscaledParameter(parameter, oldScale); break;
    default: break;
    }
    return gradient;
}

Problem::Problem(std::unique_ptr<amici::Model> model, Tables tables)
    : model_(std::move(model)), tab_(std::move(tables)) {
    if (model_->np() != tab_.np || model_->nk() != tab_.nk || model_->ny != tab_.ny)
        throw std::invalid_argument("table dimensions do not match the model");
    // Newton's method for steady states needs more than Solver::newton_maxsteps > 0: the MODEL's modes must allow
    // it (AMICI's default for both is integrationOnly, model.h:2063-2072; /amiciOptions attributes
    // steadyStateComputationMode / steadyStateSensitivityMode, deps/AMICI/src/hdf5.cpp:1433-1448).  The tables'
    // newton_maxsteps > 0 stands for "both modes integrateIfNewtonFails with this many Newton steps".
    if (tab_.newton_maxsteps > 0) {
        model_->setSteadyStateComputationMode(tab_.newton_only ? amici::SteadyStateComputationMode::newtonOnly
                                                               : amici::SteadyStateComputationMode::integrateIfNewtonFails);
        model_->setSteadyStateSensitivityMode(amici::SteadyStateSensitivityMode::integrateIfNewtonFails);
    }
    // steady_sens_newton: sensitivity mode newtonOnly; with newton_maxsteps = 0 the computation mode stays at AMICI's
    // default integrationOnly (steady state by simulation without sensitivities, then the linear solve)
    if (tab_.steady_sens_newton)
        model_->setSteadyStateSensitivityMode(amici::SteadyStateSensitivityMode::newtonOnly);
}

// multiConditionDataProvider.cpp:84-102 (column `c` of optimizationSimulationMapping)
std::vector<int> Problem::mappingFor(int c) const {
    return std::vector<int>(tab_.mapping.begin() + (size_t)c * tab_.np,
                            tab_.mapping.begin() + (size_t)(c + 1) * tab_.np);
}

// multiConditionDataProvider.cpp:323-339
std::vector<int> Problem::plistFor(const std::vector<int>& mapping) const {
    std::vector<int> plist;
    for (int i_p = 0; i_p < tab_.np; ++i_p)
        if (mapping[i_p] >= 0)
            plist.push_back(i_p);
    return plist;
}

// multiConditionDataProvider.cpp:126-169
void Problem::mapAndSetOptimizationToSimulationVariables(int c, const double* theta,
                                                         std::vector<double>& simulation) const {
    auto mapping = mappingFor(c);
    simulation.resize(tab_.np);
    for (int i = 0; i < tab_.np; ++i) {
        if (mapping[i] >= 0) {
            simulation[i] = getScaledParameter(
                getUnscaledParameter(theta[mapping[i]], tab_.pscale_opt[mapping[i]]),
                tab_.pscale_sim[(size_t)c * tab_.np + i]);
        } else if (!tab_.overrides.empty()) {
            simulation[i] = tab_.overrides[(size_t)c * tab_.np + i];
        } else {
            simulation[i] = NAN;
        }
    }
}

// what MultiConditionDataProviderHDF5::getSolver + /amiciOptions yield
// (multiConditionDataProvider.cpp:434-449) and setSensitivityOptions
// (multiConditionProblem.cpp:826-835)
std::unique_ptr<amici::Solver> Problem::makeSolver(bool sensi) const {
    auto solver = model_->getSolver();
    solver->setAbsoluteTolerance(tab_.atol);
    solver->setRelativeTolerance(tab_.rtol);
    solver->setMaxSteps(tab_.maxsteps);
    solver->setStabilityLimitFlag(tab_.stldet != 0);
    solver->setNewtonMaxSteps(tab_.newton_maxsteps);
    solver->setLinearSolver(static_cast<amici::LinearSolver>(tab_.linsol));
    solver->setInternalSensitivityMethod(static_cast<amici::InternalSensitivityMethod>(tab_.ism));
    if (sensi) {
        solver->setSensitivityOrder(amici::SensitivityOrder::first);
        solver->setSensitivityMethod(amici::SensitivityMethod::forward);
    } else {
        solver->setSensitivityOrder(amici::SensitivityOrder::none);
        solver->setSensitivityMethod(amici::SensitivityMethod::none);
    }
    return solver;
}

// multiConditionProblem.cpp:327-504
ConditionResult Problem::runAndLogSimulation(amici::Model& model, int c, bool sensi,
                                             bool want_outputs, bool collect_stats) const {
    auto t_start = std::chrono::steady_clock::now();
    auto solverTemplate = makeSolver(sensi);

    // getExperimentalDataForCondition, multiConditionDataProvider.cpp:302-321
    amici::ExpData edata(model);
    int64_t lo = tab_.t_offsets[c], hi = tab_.t_offsets[c + 1];
    edata.setTimepoints(std::vector<double>(tab_.t.begin() + lo, tab_.t.begin() + hi));
    edata.setObservedData(std::vector<double>(tab_.y.begin() + lo * tab_.ny, tab_.y.begin() + hi * tab_.ny));
    edata.setObservedDataStdDev(
        std::vector<double>(tab_.sigma.begin() + lo * tab_.ny, tab_.sigma.begin() + hi * tab_.ny));
    edata.fixedParameters.assign(tab_.k.begin() + (size_t)c * tab_.nk, tab_.k.begin() + (size_t)(c + 1) * tab_.nk);
    edata.plist = plistFor(mappingFor(c));
    // updateFixedSimulationParameters, multiConditionDataProvider.cpp:224-250
    if (!tab_.has_preeq.empty() && tab_.has_preeq[c]) {
        edata.fixedParametersPreequilibration.assign(tab_.k_preeq.begin() + (size_t)c * tab_.nk,
                                                     tab_.k_preeq.begin() + (size_t)(c + 1) * tab_.nk);
        edata.reinitialization_state_idxs_sim = tab_.reinit_idx[c];
    }

    std::unique_ptr<amici::ReturnData> rdata;
    int trial = 1;
    for (; trial <= tab_.num_trials; ++trial) {
        auto solver = std::unique_ptr<amici::Solver>(solverTemplate->clone());
        if (collect_stats || want_outputs) {
            solver->setReturnDataReportingMode(amici::RDataReporting::full);
        } else if (sensi) {
            solver->setReturnDataReportingMode(amici::RDataReporting::likelihood);
        } else {
            solver->setSensitivityMethod(amici::SensitivityMethod::none);
            solver->setReturnDataReportingMode(amici::RDataReporting::residuals);
        }
        if (rdata) {
            bool forwardFailed = std::isnan(rdata->llh);
            if (forwardFailed) {
                solver->setAbsoluteTolerance(std::pow(tab_.relaxation_factor, trial - 1) * solver->getAbsoluteTolerance());
                solver->setRelativeTolerance(std::pow(tab_.relaxation_factor, trial - 1) * solver->getRelativeTolerance());
            }
        }
        try {
            rdata = amici::runAmiciSimulation(*solver, &edata, model, false);
        } catch (std::exception const& e) {
            std::fprintf(stderr, "[oracle] error during simulation of condition %d: %s\n", c, e.what());
        }
        if (rdata && rdata->status == amici::AMICI_SUCCESS)
            break;
    }
    ConditionResult r;
    r.trials = std::min(trial, tab_.num_trials);
    r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    if (rdata) {
        r.llh = rdata->llh;
        if (sensi)
            r.gradient = rdata->sllh;
        r.modelOutput = rdata->y;
        r.modelSigmas = rdata->sigmay;
        r.status = rdata->status;
        if (!rdata->numsteps.empty()) {  // quirk Q8: empty in likelihood mode
            r.nst = rdata->numsteps.back();
            r.nfe = rdata->numrhsevals.back();
            r.netf = rdata->numerrtestfails.back();
            r.ncfn = rdata->numnonlinsolvconvfails.back();
            r.numsteps = rdata->numsteps;
            r.order = rdata->order;
        }
        if (rdata->preeq_numsteps.size() >= 2)
            r.preeq_steps = rdata->preeq_numsteps[1];
        if (rdata->posteq_numsteps.size() >= 2)
            r.posteq_steps = rdata->posteq_numsteps[1];
        return r;
    }
    r.llh = NAN;
    if (sensi)
        r.gradient.assign(edata.plist.size(), NAN);
    r.status = amici::AMICI_UNRECOVERABLE_ERROR;
    return r;
}

// A1 + runSimulations + aggregateLikelihood:
// multiConditionProblem.cpp:647-675, 717-773, 775-824.  The OpenMP loop is the
// structure of the reference's non-MPI path (amiciSimulationRunner.cpp:107-131)
// but uses the TRUE condition index as the MPI path does (:64-76; quirk Q1 of
// the shared-memory path is not reproduced).
int Problem::evaluate(const double* theta, const std::vector<int>& idx, double& fval, double* grad,
                      int nthreads, std::vector<ConditionResult>* per_condition, bool collect_stats) const {
    const bool sensi = grad != nullptr;
    fval = 0.0;
    if (grad)
        std::fill(grad, grad + tab_.n_opt, 0.0);
    const int n = (int)idx.size();
    std::vector<ConditionResult> results(n);

#ifdef _OPENMP
    if (nthreads <= 0)
        nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        // messageHandler clones model (+solver) per job, multiConditionProblem.cpp:577-578
        std::unique_ptr<amici::Model> model(model_->clone());
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int r = 0; r < n; ++r) {
            int c = idx[r];
            // updateSimulationParametersAndScale, multiConditionDataProvider.cpp:451-469
            std::vector<amici::ParameterScaling> scaleSim(tab_.np);
            for (int i = 0; i < tab_.np; ++i)
                scaleSim[i] = static_cast<amici::ParameterScaling>(tab_.pscale_sim[(size_t)c * tab_.np + i]);
            std::vector<double> p;
            model->setParameterScale(scaleSim);
            mapAndSetOptimizationToSimulationVariables(c, theta, p);
            model->setParameters(p);
            results[r] = runAndLogSimulation(*model, c, sensi, false, collect_stats);
        }
    }

    // aggregateLikelihood, multiConditionProblem.cpp:775-824 (fixed order: index-list order)
    int errors = 0;
    for (int r = 0; r < n; ++r) {
        int c = idx[r];
        auto const& res = results[r];
        errors += res.status != amici::AMICI_SUCCESS;
        fval -= res.llh;
        if (grad) {
            std::vector<double> p;
            mapAndSetOptimizationToSimulationVariables(c, theta, p);
            // mapSimulationToOptimizationGradientAddMultiply, multiConditionDataProvider.cpp:104-124
            auto mapping = mappingFor(c);
            auto plist = plistFor(mapping);
            for (size_t i_plist = 0; i_plist < plist.size(); ++i_plist) {
                int i_p = plist[i_plist];
                double newGrad = applyChainRule(res.gradient[i_plist], p[i_p],
                                                tab_.pscale_sim[(size_t)c * tab_.np + i_p],
                                                tab_.pscale_opt[mapping[i_p]]);
                grad[mapping[i_p]] += -1.0 * newGrad;
            }
        }
    }
    if (per_condition)
        *per_condition = std::move(results);
    if (errors || !std::isfinite(fval)) {
        fval = std::numeric_limits<double>::infinity();
        return 1;  // functionEvaluationFailure
    }
    return 0;
}

// multiConditionProblem.cpp:506-563
int Problem::getModelOutputsAndSigmas(const double* theta, std::vector<std::vector<double>>& y,
                                      std::vector<std::vector<double>>& sigmay, int nthreads) const {
    const int n = tab_.n_conditions;
    y.assign(n, {});
    sigmay.assign(n, {});
    int errors = 0;
#ifdef _OPENMP
    if (nthreads <= 0)
        nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads) reduction(+ : errors)
#endif
    {
        std::unique_ptr<amici::Model> model(model_->clone());
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int c = 0; c < n; ++c) {
            std::vector<amici::ParameterScaling> scaleSim(tab_.np);
            for (int i = 0; i < tab_.np; ++i)
                scaleSim[i] = static_cast<amici::ParameterScaling>(tab_.pscale_sim[(size_t)c * tab_.np + i]);
            std::vector<double> p;
            model->setParameterScale(scaleSim);
            mapAndSetOptimizationToSimulationVariables(c, theta, p);
            model->setParameters(p);
            auto res = runAndLogSimulation(*model, c, false, true, false);
            errors += res.status;
            y[c] = res.modelOutput;
            sigmay[c] = res.modelSigmas;
        }
    }
    return errors == 0 ? 0 : 1;
}

}  // namespace pb200_oracle
