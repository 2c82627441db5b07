// b200SummedGradientFunction.h -- header-only C++ host side of the drop-in boundary.
//
// parpe::B200SummedGradientFunction implements parpe::SummedGradientFunction<int> with the public
// surface of parpe::AmiciSummedGradientFunction that callers use
// (include/parpeamici/multiConditionProblem.h:95-237): evaluate (vector and single-index
// overloads), numParameters, getParameterIds, getModelOutputsAndSigmas, getAllMeasurements,
// getParameterScaling.  parpe::B200HierarchicalOptimizationWrapper implements
// parpe::GradientFunction with the behaviour of parpe::HierarchicalOptimizationWrapper::evaluate
// (src/parpeamici/hierarchicalOptimization.cpp:142-238).  Both only forward to the C ABI
// (include/parpe_b200.h); all numerics run in libparpe_b200.so on the GPU.
//
// Error conventions are the reference's: a failed simulation yields fval = +infinity and
// functionEvaluationFailure (multiConditionProblem.cpp:669-672); API misuse (wrong vector
// sizes) throws, as gsl Expects / ParPEException do in parPE.
#ifndef PB200_B200_SUMMED_GRADIENT_FUNCTION_H
#define PB200_B200_SUMMED_GRADIENT_FUNCTION_H

#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/parpe_b200.h"
#include "parpe_interfaces.h"

namespace parpe {

/// In-memory problem description: what MultiConditionDataProviderHDF5 serves from the HDF5 input
/// (doc/hdf5.md), owned as std::vectors; desc() views it as the C struct.
struct B200ProblemTables {
    std::string modelName;
    int numConditions = 0, numOptimizationParameters = 0, np = 0, nk = 0, ny = 0;
    std::vector<int64_t> tOffsets;            // [C+1]
    std::vector<double> t, y, sigma;          // y/sigma: row-major obs + it*ny per condition
    std::vector<double> k;                    // [C*nk]
    std::vector<int32_t> mapping, pscaleSim;  // [C*np]
    std::vector<int32_t> pscaleOpt;           // [n_opt]
    std::vector<double> overrides;            // [C*np] or empty
    std::vector<std::string> parameterIds;    // /parameters/parameterNames
    std::vector<double> lowerBound, upperBound, nominalValues;  // /parameters/{lowerBound,upperBound,nominalValues}
    // pre-equilibration (simulationConditions column 0, reinitializationIndices); hasPreeq empty = none
    std::vector<int32_t> hasPreeq;                        // [C]
    std::vector<double> kPreeq;                           // [C*nk]
    std::vector<std::vector<int>> reinitializationIndices; // [C] ragged
    double t0 = 0.0, atol = 1e-16, rtol = 1e-8;
    int maxSteps = 10000;
    bool stabilityLimitDetection = true;
    int numSimulationTrials = 6;                         // PARPE_NUM_SIMULATION_TRIALS
    double integrationToleranceRelaxationFactor = 100.0; // PARPE_INTEGRATION_TOLERANCE_RELAXATION_FACTOR
    int lanesPerCondition = 0;
    int newtonMaxSteps = 0;  // amici::Solver::setNewtonMaxSteps (steady states; 0 = by simulation only)
    bool newtonOnly = false; // steadyStateComputationMode newtonOnly: no simulation if Newton's method fails
    bool steadySensNewton = false; // steadyStateSensitivityMode newtonOnly with a simulation to steady state
    int internalSensitivityMethod = 0; // 0 / 2 staggered, 1 simultaneous
    int linearSolver = 0;    // large models: 0 automatic, 1 dense (LinearSolver::dense), 2 sparse (LinearSolver::KLU)

    /// The tables of a parPE parameter-estimation input file (doc/hdf5.md), assembled by the library's built-in
    /// HDF5 reader (pb200_h5_load_problem) -- what MultiConditionDataProviderHDF5 serves per condition, read once.
    /// Hierarchical index / map tables, if the file has them, are returned through `hierarchical`.
    static B200ProblemTables fromHdf5(std::string const& path, std::string const& modelName,
                                      struct B200HierarchicalTables* hierarchical = nullptr);

    pb200_problem_desc desc() const {
        pb200_problem_desc d{};
        d.abi_version = PB200_ABI_VERSION;
        d.model_name = modelName.c_str();
        d.n_conditions = numConditions;
        d.n_opt = numOptimizationParameters;
        d.np = np; d.nk = nk; d.ny = ny;
        d.t_offsets = tOffsets.data();
        d.t = t.data(); d.y = y.data(); d.sigma = sigma.data(); d.k = k.data();
        d.mapping = mapping.data(); d.pscale_sim = pscaleSim.data(); d.pscale_opt = pscaleOpt.data();
        d.overrides = overrides.empty() ? nullptr : overrides.data();
        d.t0 = t0; d.atol = atol; d.rtol = rtol;
        d.maxsteps = maxSteps;
        d.stldet = stabilityLimitDetection ? 1 : 0;
        d.sensi_errcon = 1;
        d.num_trials = numSimulationTrials;
        d.relaxation_factor = integrationToleranceRelaxationFactor;
        d.lanes_per_condition = lanesPerCondition;
        d.newton_maxsteps = newtonMaxSteps;
        d.newton_only = newtonOnly ? 1 : 0;
        d.linsol = linearSolver;
        d.steady_sens_newton = steadySensNewton ? 1 : 0;
        d.ism = internalSensitivityMethod;
        return d;
    }
};

class B200SummedGradientFunction : public SummedGradientFunction<int> {
  public:
    explicit B200SummedGradientFunction(B200ProblemTables tables, int device = 0) : tables_(std::move(tables)) {
        pb200_problem_desc d = tables_.desc();
        if (pb200_create(&d, device, &handle_) != 0)
            throw std::runtime_error(std::string("pb200_create: ") + pb200_last_error());
        if (!tables_.hasPreeq.empty()) {
            try {
                setPreequilibration(tables_.hasPreeq, tables_.kPreeq, tables_.reinitializationIndices);
            } catch (...) {
                pb200_destroy(handle_);
                throw;
            }
        }
    }
    ~B200SummedGradientFunction() override { pb200_destroy(handle_); }
    B200SummedGradientFunction(const B200SummedGradientFunction&) = delete;
    B200SummedGradientFunction& operator=(const B200SummedGradientFunction&) = delete;

    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, int dataset, double& fval,
                                      gsl::span<double> gradient, Logger* logger, double* cpuTime) const override {
        return evaluate(parameters, std::vector<int>{dataset}, fval, gradient, logger, cpuTime);
    }

    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, std::vector<int> datasets, double& fval,
                                      gsl::span<double> gradient, Logger* /*logger*/,
                                      double* cpuTime) const override {
        if (!gradient.empty() && gradient.size() != parameters.size())
            throw std::invalid_argument("gradient size does not match parameter vector");
        double seconds = 0.0;
        const int rc = pb200_evaluate(handle_, parameters.data(), static_cast<int>(parameters.size()),
                                      datasets.data(), static_cast<int>(datasets.size()), &fval,
                                      gradient.empty() ? nullptr : gradient.data(), &seconds);
        if (rc < 0) throw std::runtime_error(std::string("pb200_evaluate: ") + pb200_last_error());
        if (cpuTime) *cpuTime = seconds;
        return rc == PB200_EVAL_SUCCESS ? functionEvaluationSuccess : functionEvaluationFailure;
    }

    int numParameters() const override { return tables_.numOptimizationParameters; }
    std::vector<std::string> getParameterIds() const override { return tables_.parameterIds; }

    /// AmiciSummedGradientFunction::getModelOutputsAndSigmas (multiConditionProblem.cpp:506-563)
    virtual FunctionEvaluationStatus getModelOutputsAndSigmas(gsl::span<const double> parameters,
                                                              std::vector<std::vector<double>>& modelOutputs,
                                                              std::vector<std::vector<double>>& modelSigmas,
                                                              Logger* /*logger*/, double* /*cpuTime*/) const {
        const size_t n = tables_.y.size();
        std::vector<double> y(n), s(n);
        const int rc = pb200_get_model_outputs(handle_, parameters.data(), static_cast<int>(parameters.size()),
                                               y.data(), s.data());
        if (rc < 0) throw std::runtime_error(std::string("pb200_get_model_outputs: ") + pb200_last_error());
        modelOutputs.assign(tables_.numConditions, {});
        modelSigmas.assign(tables_.numConditions, {});
        for (int c = 0; c < tables_.numConditions; ++c) {
            const auto lo = tables_.tOffsets[c] * tables_.ny, hi = tables_.tOffsets[c + 1] * tables_.ny;
            modelOutputs[c].assign(y.begin() + lo, y.begin() + hi);
            modelSigmas[c].assign(s.begin() + lo, s.begin() + hi);
        }
        return rc == PB200_EVAL_SUCCESS ? functionEvaluationSuccess : functionEvaluationFailure;
    }

    virtual std::vector<std::vector<double>> getAllMeasurements() const {
        std::vector<std::vector<double>> m(tables_.numConditions);
        for (int c = 0; c < tables_.numConditions; ++c)
            m[c].assign(tables_.y.begin() + tables_.tOffsets[c] * tables_.ny,
                        tables_.y.begin() + tables_.tOffsets[c + 1] * tables_.ny);
        return m;
    }

    /// amici::ParameterScaling of an optimisation parameter: 0 none, 1 ln, 2 log10
    virtual int getParameterScaling(int parameterIndex) const { return tables_.pscaleOpt.at(parameterIndex); }

    /// all condition indices, as MultiConditionProblem::getTrainingData returns them
    std::vector<int> getTrainingData() const {
        std::vector<int> idx(tables_.numConditions);
        for (int i = 0; i < tables_.numConditions; ++i) idx[i] = i;
        return idx;
    }

    /// This process simulates the share `rank` of `world` of every index list (what
    /// LoadBalancerMaster/Worker distribute over MPI ranks in the reference).  evaluate() still returns
    /// the GLOBAL fval / gradient / status on every rank -- as the reference master does after summing
    /// all workers' results -- and therefore needs a collective: initNccl() or setAllreduce().
    /// Pre-equilibration (include/parpe_b200.h: pb200_set_preequilibration) -- what
    /// MultiConditionDataProviderHDF5::updateFixedSimulationParameters puts into
    /// ExpData::fixedParametersPreequilibration / reinitialization_state_idxs_sim
    /// (src/parpeamici/multiConditionDataProvider.cpp:224-250): per simulation condition a flag, the constants of
    /// the pre-equilibration condition (row-major [C x nk]) and the list of states re-initialised afterwards.
    void setPreequilibration(std::vector<int32_t> const& hasPreeq, std::vector<double> const& kPreeq,
                             std::vector<std::vector<int>> const& reinitializationIndices = {}) {
        std::vector<int64_t> off(hasPreeq.size() + 1, 0);
        std::vector<int32_t> idx;
        for (size_t c = 0; c < hasPreeq.size(); ++c) {
            if (c < reinitializationIndices.size())
                idx.insert(idx.end(), reinitializationIndices[c].begin(), reinitializationIndices[c].end());
            off[c + 1] = static_cast<int64_t>(idx.size());
        }
        if (idx.empty()) idx.push_back(0);
        if (pb200_set_preequilibration(handle_, hasPreeq.data(), kPreeq.data(), off.data(), idx.data()) != 0)
            throw std::runtime_error(std::string("pb200_set_preequilibration: ") + pb200_last_error());
    }
    void setPartition(int rank, int world) {
        if (pb200_set_partition(handle_, rank, world) != 0)
            throw std::runtime_error(std::string("pb200_set_partition: ") + pb200_last_error());
    }
    /// Host-transport collective (include/parpe_b200.h: pb200_set_allreduce), e.g. a function calling
    /// MPI_Allreduce in place: used by evaluate() and the hierarchical wrapper on a partitioned handle.
    void setAllreduce(pb200_allreduce_fn fn, void* user) {
        if (pb200_set_allreduce(handle_, fn, user) != 0)
            throw std::runtime_error(std::string("pb200_set_allreduce: ") + pb200_last_error());
    }
    /// 128-byte NCCL id, to be created on rank 0 and distributed to all ranks (MPI_Bcast in parPE)
    static std::vector<char> ncclUniqueId() {
        std::vector<char> id(128);
        if (pb200_nccl_unique_id(id.data()) != 0)
            throw std::runtime_error(std::string("pb200_nccl_unique_id: ") + pb200_last_error());
        return id;
    }
    /// In-library collective: joins the NCCL communicator of the `world` handles of this problem and sets
    /// the partition; from then on every evaluation all-reduces (nllh, #errors, gradient) on the device.
    /// Must be called by all ranks concurrently (it blocks until the communicator is complete).
    void initNccl(std::vector<char> const& id, int rank, int world) {
        if (id.size() != 128) throw std::invalid_argument("NCCL unique id must have 128 bytes");
        if (pb200_nccl_init(handle_, id.data(), rank, world) != 0)
            throw std::runtime_error(std::string("pb200_nccl_init: ") + pb200_last_error());
    }

    /// The standalone simulator's result datasets (src/parpeamici/standaloneSimulator.cpp:57-270; doc/hdf5.md:228-252):
    /// llh, parameters, t/<m>, yMes/<m>, ySim/<m> under `group` of a new HDF5 file (pb200_write_simulation_results).
    FunctionEvaluationStatus writeSimulationResults(gsl::span<const double> parameters, std::string const& path,
                                                    std::string const& group = "/multistarts/0") const {
        const int rc = pb200_write_simulation_results(handle_, parameters.data(), static_cast<int>(parameters.size()),
                                                      path.c_str(), group.c_str());
        if (rc < 0) throw std::runtime_error(std::string("pb200_write_simulation_results: ") + pb200_last_error());
        return rc == PB200_EVAL_SUCCESS ? functionEvaluationSuccess : functionEvaluationFailure;
    }
    pb200_handle* handle() const { return handle_; }
    const B200ProblemTables& tables() const { return tables_; }

  private:
    B200ProblemTables tables_;
    pb200_handle* handle_ = nullptr;
};

/// The same objective over SEVERAL GPUs of one process -- what the reference reaches with MPI ranks
/// (LoadBalancerMaster distributing packages to LoadBalancerWorkers, loadBalancerMaster.cpp:104-134,
/// 200-244): one handle per device, every index list sharded contiguously over them, the partial
/// (nllh, #errors, gradient) combined by ONE ncclAllReduce inside the library.  evaluate() drives all
/// devices (a host thread per extra device for the duration of the call) and returns the global
/// result; Ipopt / Ceres / the minibatch optimizer use it like any parpe::SummedGradientFunction<int>.
/// Per-simulation result logging: what parPE's logSimulation / saveSimulation writes per simulation job
/// (src/parpeamici/multiConditionProblem.cpp:254-325; doc/hdf5.md:190-226), collected in memory after each
/// evaluate() and written once (pb200_result_log_*).
class B200ResultLog {
  public:
    B200ResultLog() {
        if (pb200_result_log_create(&log_) != 0)
            throw std::runtime_error(std::string("pb200_result_log_create: ") + pb200_last_error());
    }
    ~B200ResultLog() { pb200_result_log_destroy(log_); }
    B200ResultLog(const B200ResultLog&) = delete;
    B200ResultLog& operator=(const B200ResultLog&) = delete;
    /// the per-condition results of `fun`'s last evaluate as records of `group`; returns their number
    int append(B200SummedGradientFunction const& fun, std::string const& group, int optimizationIdx, int iterationIdx) {
        const int n = pb200_result_log_append(log_, fun.handle(), group.c_str(), optimizationIdx, iterationIdx);
        if (n < 0) throw std::runtime_error(std::string("pb200_result_log_append: ") + pb200_last_error());
        return n;
    }
    int numRecords(std::string const& group) const { return pb200_result_log_num_records(log_, group.c_str()); }
    void write(std::string const& path) const {
        if (pb200_result_log_write(log_, path.c_str()) != 0)
            throw std::runtime_error(std::string("pb200_result_log_write: ") + pb200_last_error());
    }

  private:
    pb200_result_log* log_ = nullptr;
};

class B200MultiGpuSummedGradientFunction : public SummedGradientFunction<int> {
  public:
    B200MultiGpuSummedGradientFunction(B200ProblemTables const& tables, std::vector<int> const& devices) {
        if (devices.empty()) throw std::invalid_argument("no devices");
        for (int d : devices) ranks_.push_back(std::make_unique<B200SummedGradientFunction>(tables, d));
        const int world = static_cast<int>(devices.size());
        if (world > 1) {
            const std::vector<char> id = B200SummedGradientFunction::ncclUniqueId();
            std::vector<std::string> errors(world);
            std::vector<std::thread> th;
            for (int r = 0; r < world; ++r)
                th.emplace_back([&, r] {
                    try { ranks_[r]->initNccl(id, r, world); } catch (std::exception const& e) { errors[r] = e.what(); }
                });
            for (auto& t : th) t.join();
            for (auto& e : errors)
                if (!e.empty()) throw std::runtime_error(e);
        }
    }

    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, int dataset, double& fval,
                                      gsl::span<double> gradient, Logger* logger, double* cpuTime) const override {
        return evaluate(parameters, std::vector<int>{dataset}, fval, gradient, logger, cpuTime);
    }

    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, std::vector<int> datasets, double& fval,
                                      gsl::span<double> gradient, Logger* logger, double* cpuTime) const override {
        const int world = static_cast<int>(ranks_.size());
        std::vector<FunctionEvaluationStatus> status(world, functionEvaluationFailure);
        std::vector<double> fvals(world, 0.0), times(world, 0.0);
        std::vector<std::vector<double>> grads(world, std::vector<double>(gradient.size()));
        std::vector<std::string> errors(world);
        auto run = [&](int r) {
            try {
                status[r] = ranks_[r]->evaluate(parameters, datasets, fvals[r],
                                                gsl::span<double>(grads[r].data(), grads[r].size()), logger, &times[r]);
            } catch (std::exception const& e) { errors[r] = e.what(); }
        };
        std::vector<std::thread> th;
        for (int r = 1; r < world; ++r) th.emplace_back(run, r);
        run(0);
        for (auto& t : th) t.join();
        for (auto& e : errors)
            if (!e.empty()) throw std::runtime_error(e);
        // every rank holds the same all-reduced numbers; rank 0's are returned
        fval = fvals[0];
        for (size_t j = 0; j < gradient.size(); ++j) gradient[j] = grads[0][j];
        if (cpuTime) {
            double t = 0.0;
            for (double v : times) t = v > t ? v : t;
            *cpuTime = t;
        }
        return status[0];
    }

    int numParameters() const override { return ranks_[0]->numParameters(); }
    std::vector<std::string> getParameterIds() const override { return ranks_[0]->getParameterIds(); }
    std::vector<int> getTrainingData() const { return ranks_[0]->getTrainingData(); }
    int numDevices() const { return static_cast<int>(ranks_.size()); }
    B200SummedGradientFunction& rank(int r) const { return *ranks_.at(r); }

  private:
    std::vector<std::unique_ptr<B200SummedGradientFunction>> ranks_;
};

/// (inner parameter, condition, observable) row tables of the hierarchical problem
struct B200HierarchicalTables {
    std::vector<int32_t> scalingIdx, offsetIdx, sigmaIdx;  // optimisation-parameter indices, ascending
    std::vector<int32_t> scalingMap, offsetMap, sigmaMap;  // rows of 3, row-major
    pb200_hier_desc desc() const {
        pb200_hier_desc d{};
        d.n_scaling = static_cast<int32_t>(scalingIdx.size());
        d.n_offset = static_cast<int32_t>(offsetIdx.size());
        d.n_sigma = static_cast<int32_t>(sigmaIdx.size());
        d.scaling_idx = scalingIdx.data(); d.offset_idx = offsetIdx.data(); d.sigma_idx = sigmaIdx.data();
        d.n_scaling_rows = static_cast<int32_t>(scalingMap.size() / 3);
        d.n_offset_rows = static_cast<int32_t>(offsetMap.size() / 3);
        d.n_sigma_rows = static_cast<int32_t>(sigmaMap.size() / 3);
        d.scaling_map = scalingMap.data(); d.offset_map = offsetMap.data(); d.sigma_map = sigmaMap.data();
        return d;
    }
};

inline B200ProblemTables B200ProblemTables::fromHdf5(std::string const& path, std::string const& modelName,
                                                     B200HierarchicalTables* hierarchical) {
    pb200_h5_problem* P = nullptr;
    if (pb200_h5_load_problem(path.c_str(), modelName.c_str(), &P) != 0)
        throw std::runtime_error(std::string("pb200_h5_load_problem: ") + pb200_last_error());
    const pb200_problem_desc& d = *pb200_h5_problem_desc(P);
    B200ProblemTables T;
    T.modelName = modelName;
    T.numConditions = d.n_conditions;
    T.numOptimizationParameters = d.n_opt;
    T.np = d.np; T.nk = d.nk; T.ny = d.ny;
    const size_t C = static_cast<size_t>(d.n_conditions);
    T.tOffsets.assign(d.t_offsets, d.t_offsets + C + 1);
    const size_t nt = static_cast<size_t>(T.tOffsets.back());
    T.t.assign(d.t, d.t + nt);
    T.y.assign(d.y, d.y + nt * d.ny);
    T.sigma.assign(d.sigma, d.sigma + nt * d.ny);
    T.k.assign(d.k, d.k + C * d.nk);
    T.mapping.assign(d.mapping, d.mapping + C * d.np);
    T.pscaleSim.assign(d.pscale_sim, d.pscale_sim + C * d.np);
    T.pscaleOpt.assign(d.pscale_opt, d.pscale_opt + d.n_opt);
    if (d.overrides) T.overrides.assign(d.overrides, d.overrides + C * d.np);
    T.t0 = d.t0; T.atol = d.atol; T.rtol = d.rtol;
    T.maxSteps = d.maxsteps;
    T.stabilityLimitDetection = d.stldet != 0;
    T.newtonMaxSteps = d.newton_maxsteps;
    T.newtonOnly = d.newton_only != 0;
    T.linearSolver = d.linsol;
    T.steadySensNewton = d.steady_sens_newton != 0;
    T.internalSensitivityMethod = d.ism;
    {
        const double* lo = nullptr; const double* up = nullptr; const double* nom = nullptr; const char* names = nullptr;
        const int n = pb200_h5_problem_parameters(P, &lo, &up, &nom, &names);
        if (lo) T.lowerBound.assign(lo, lo + n);
        if (up) T.upperBound.assign(up, up + n);
        if (nom) T.nominalValues.assign(nom, nom + n);
        if (names) {
            std::string cur;
            for (const char* c = names; *c; ++c) {
                if (*c == '\n') { T.parameterIds.push_back(cur); cur.clear(); }
                else cur.push_back(*c);
            }
        }
    }
    {
        const int32_t* flag = nullptr; const double* kp = nullptr; const int64_t* off = nullptr; const int32_t* idx = nullptr;
        if (pb200_h5_problem_preeq(P, &flag, &kp, &off, &idx)) {
            T.hasPreeq.assign(flag, flag + C);
            T.kPreeq.assign(kp, kp + C * d.nk);
            T.reinitializationIndices.resize(C);
            for (size_t c = 0; c < C; ++c) T.reinitializationIndices[c].assign(idx + off[c], idx + off[c + 1]);
        }
    }
    if (hierarchical) {
        *hierarchical = B200HierarchicalTables();
        if (const pb200_hier_desc* h = pb200_h5_problem_hier(P)) {
            hierarchical->scalingIdx.assign(h->scaling_idx, h->scaling_idx + h->n_scaling);
            hierarchical->offsetIdx.assign(h->offset_idx, h->offset_idx + h->n_offset);
            hierarchical->sigmaIdx.assign(h->sigma_idx, h->sigma_idx + h->n_sigma);
            hierarchical->scalingMap.assign(h->scaling_map, h->scaling_map + 3 * h->n_scaling_rows);
            hierarchical->offsetMap.assign(h->offset_map, h->offset_map + 3 * h->n_offset_rows);
            hierarchical->sigmaMap.assign(h->sigma_map, h->sigma_map + 3 * h->n_sigma_rows);
        }
    }
    pb200_h5_problem_free(P);
    return T;
}

class B200HierarchicalOptimizationWrapper : public GradientFunction {
  public:
    B200HierarchicalOptimizationWrapper(B200SummedGradientFunction* wrapped, const B200HierarchicalTables& h)
        : wrapped_(wrapped) {
        pb200_hier_desc d = h.desc();
        if (pb200_hier_attach(wrapped_->handle(), &d) != 0)
            throw std::runtime_error(std::string("pb200_hier_attach: ") + pb200_last_error());
        numOuter_ = pb200_hier_num_outer(wrapped_->handle());
        std::vector<int32_t> inner(h.scalingIdx);
        inner.insert(inner.end(), h.offsetIdx.begin(), h.offsetIdx.end());
        inner.insert(inner.end(), h.sigmaIdx.begin(), h.sigmaIdx.end());
        auto all = wrapped_->getParameterIds();
        for (int i = 0; i < static_cast<int>(all.size()); ++i) {
            bool isInner = false;
            for (auto j : inner) isInner = isInner || (j == i);
            if (!isInner) outerIds_.push_back(all[i]);
        }
    }
    using GradientFunction::evaluate;
    FunctionEvaluationStatus evaluate(gsl::span<double const> reducedParameters, double& fval,
                                      gsl::span<double> gradient, Logger* /*logger*/,
                                      double* cpuTime) const override {
        std::vector<double> full;
        return evaluate(reducedParameters, fval, gradient, full, cpuTime);
    }
    /// also returns the spliced full parameter vector (HierarchicalOptimizationWrapper::evaluate
    /// with fullParameters, hierarchicalOptimization.cpp:142-150)
    FunctionEvaluationStatus evaluate(gsl::span<double const> reducedParameters, double& fval,
                                      gsl::span<double> gradient, std::vector<double>& fullParameters,
                                      double* cpuTime) const {
        if (static_cast<int>(reducedParameters.size()) != numOuter_)
            throw std::invalid_argument("Reduced parameter vector size " + std::to_string(reducedParameters.size()) +
                                        " does not match numParameters " + std::to_string(numOuter_));
        fullParameters.resize(wrapped_->numParameters());
        double seconds = 0.0;
        const int rc = pb200_hier_evaluate(wrapped_->handle(), reducedParameters.data(), numOuter_, &fval,
                                           gradient.empty() ? nullptr : gradient.data(), fullParameters.data(),
                                           &seconds);
        if (rc < 0) throw std::runtime_error(std::string("pb200_hier_evaluate: ") + pb200_last_error());
        if (cpuTime) *cpuTime += seconds;
        return rc == PB200_EVAL_SUCCESS ? functionEvaluationSuccess : functionEvaluationFailure;
    }
    int numParameters() const override { return numOuter_; }
    std::vector<std::string> getParameterIds() const override { return outerIds_; }

  private:
    B200SummedGradientFunction* wrapped_;
    int numOuter_ = 0;
    std::vector<std::string> outerIds_;
};

}  // namespace parpe
#endif
