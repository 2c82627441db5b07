// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// CPU restatement of parPE's glue around AMICI for the hot path
// AmiciSummedGradientFunction::evaluate, over in-memory tables instead of HDF5.
// parPE's own translation units need HDF5 C++, Boost.Serialization and MPI
// (none in this image), so their arithmetic/indexing is restated here; the
// numerical path underneath is the reference's own AMICI/CVODES/KLU code
// compiled from /root/reference into oracle/_ref/libamici_ref.a.
//
// Parity status: the numerical path is the reference itself.  The glue (A1, A2,
// A5, A6, A8 of SURVEY.md section 8) has no test in the reference ("parity
// unpinned" by reference tests); it is pinned by tests/test_oracle_glue.py
// through finite-difference gradient checks (the reference's own method,
// src/parpeoptimization/optimizationProblem.cpp:50-131) and by the known-answer
// cases of examples/parpeamici/steadystate/exampleSteadystateScaledTest.h.
#ifndef PB200_ORACLE_PARPE_GLUE_HPP
#define PB200_ORACLE_PARPE_GLUE_HPP

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "amici/amici.h"

namespace pb200_oracle {

/** In-memory stand-in for parpe::MultiConditionDataProviderHDF5
 *  (include/parpeamici/multiConditionDataProvider.h:203-443). */
struct Tables {
    int n_conditions = 0, n_opt = 0, np = 0, nk = 0, ny = 0;
    std::vector<int64_t> t_offsets;     // [C+1]
    std::vector<double> t, y, sigma;    // y/sigma: row-major obs + it*ny
    std::vector<double> k;              // [C*nk]
    std::vector<int> mapping;           // [C*np]
    std::vector<int> pscale_sim;        // [C*np]
    std::vector<int> pscale_opt;        // [n_opt]
    std::vector<double> overrides;      // [C*np] or empty
    // pre-equilibration (multiConditionDataProvider.cpp:224-250): per simulation the constants of the
    // pre-equilibration condition and the states re-initialised afterwards; has_preeq empty = none
    std::vector<int> has_preeq;                  // [C]
    std::vector<double> k_preeq;                 // [C*nk]
    std::vector<std::vector<int>> reinit_idx;    // [C] ragged
    double t0 = 0.0, atol = 1e-16, rtol = 1e-8;
    int maxsteps = 10000;
    int stldet = 1;
    int num_trials = 6;
    double relaxation_factor = 100.0;
    int newton_maxsteps = 0;  // amici::Solver::setNewtonMaxSteps (AMICI default 0)
    int newton_only = 0;      // 1 = steadyStateComputationMode newtonOnly (with newton_maxsteps > 0)
    int steady_sens_newton = 0;  // 1 = steadyStateSensitivityMode newtonOnly
    int ism = 2;     // amici::InternalSensitivityMethod (1 simultaneous, 2 staggered)
    int linsol = 9;  // amici::LinearSolver (1 dense, 9 KLU)
};

struct ConditionResult {  // AmiciResultPackageSimple, amiciSimulationRunner.h:157-182
    double llh = 0.0;
    double seconds = 0.0;
    std::vector<double> gradient;      // sllh, plist-compacted
    std::vector<double> modelOutput;   // y  [nt*ny]
    std::vector<double> modelSigmas;   // sigmay [nt*ny]
    int status = 0;
    int trials = 0;
    long nst = -1, nfe = -1, nsetups = -1, nni = -1, netf = -1, ncfn = -1;
    std::vector<int> numsteps, order;  // per output time (full reporting only)
    long preeq_steps = -1, posteq_steps = -1;  // internal steps of the simulations to steady state (full reporting only)
};

class Problem {
  public:
    Problem(std::unique_ptr<amici::Model> model, Tables tables);

    /** A1: AmiciSummedGradientFunction::evaluate, multiConditionProblem.cpp:647-675 */
    int evaluate(const double* theta, const std::vector<int>& idx, double& fval,
                 double* grad /*nullable*/, int nthreads, std::vector<ConditionResult>* per_condition,
                 bool collect_stats) const;

    /** getModelOutputsAndSigmas, multiConditionProblem.cpp:506-563 */
    int getModelOutputsAndSigmas(const double* theta, std::vector<std::vector<double>>& y,
                                 std::vector<std::vector<double>>& sigmay, int nthreads) const;

    // A5 pieces (public for unit tests)
    std::vector<int> mappingFor(int c) const;
    std::vector<int> plistFor(const std::vector<int>& mapping) const;
    void mapAndSetOptimizationToSimulationVariables(int c, const double* theta, std::vector<double>& p_sim) const;

    const Tables& tables() const { return tab_; }
    Tables& mutableTables() { return tab_; }
    const amici::Model& model() const { return *model_; }

  private:
    ConditionResult runAndLogSimulation(amici::Model& model, int c, bool sensi, bool want_outputs,
                                        bool collect_stats) const;
    std::unique_ptr<amici::Solver> makeSolver(bool sensi) const;

    std::unique_ptr<amici::Model> model_;
    Tables tab_;
};

// free functions restated from the reference (unit-tested against its known answers)
double getUnscaledParameter(double v, int scale);  // deps/AMICI/src/misc.cpp:26-43
double getScaledParameter(double v, int scale);    // deps/AMICI/src/misc.cpp:45-63
double applyChainRule(double gradient, double parameter, int oldScale, int newScale);  // multiConditionDataProvider.cpp:762-796

}  // namespace pb200_oracle
#endif
