// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// CPU restatement of parpe::HierarchicalOptimizationWrapper's analytic inner
// problem (src/parpeamici/hierarchicalOptimization.cpp), row A15 of SURVEY.md
// section 8.  Pinned by the reference's own known-answer tests
// (tests/parpeamici/hierarchicalOptimizationTest.cpp:201-223,265-504,580-590),
// restated in tests/test_hierarchical_oracle.py.
#ifndef PB200_ORACLE_HIERARCHICAL_HPP
#define PB200_ORACLE_HIERARCHICAL_HPP

#include <map>
#include <vector>

#include "../include/parpe_b200.h"
#include "parpe_glue.hpp"

namespace pb200_oracle {

/** AnalyticalParameterHdf5Reader semantics over an in-memory (param, condition,
 *  observable) row table (hierarchicalOptimizationAnalyticalParameterProvider.cpp). */
struct AnalyticalParameterProvider {
    std::vector<int> optimizationParameterIndices;
    std::vector<std::map<int, std::vector<int>>> mapping;  // [param][condition] -> observables

    std::vector<int> getConditionsForParameter(int parameterIndex) const;
    const std::vector<int>& getObservablesForParameter(int parameterIndex, int conditionIdx) const;
    static AnalyticalParameterProvider fromRows(const int32_t* idx, int n, const int32_t* rows, int n_rows);
};

struct HierarchicalSpec {
    AnalyticalParameterProvider scaling, offset, sigma;
    int numObservables = 0;
    int numConditions = 0;
    int nOpt = 0;
    std::vector<int> pscale_opt;
    int numOuter() const {
        return nOpt - (int)scaling.optimizationParameterIndices.size() -
               (int)offset.optimizationParameterIndices.size() - (int)sigma.optimizationParameterIndices.size();
    }
    static HierarchicalSpec fromDesc(const pb200_hier_desc& hd, const Tables& tab);
};

using Matrix = std::vector<std::vector<double>>;

// hierarchicalOptimization.cpp:672-700
double getDefaultScalingFactor(int scaling);
double getDefaultOffsetParameter(int scaling);
// :702-794, :796-851, :853-930
double computeAnalyticalScalings(int scalingIdx, const Matrix& modelOutputsUnscaled, const Matrix& measurements,
                                 const AnalyticalParameterProvider& scalingReader, int numObservables);
double computeAnalyticalOffsets(int offsetIdx, const Matrix& modelOutputsUnscaled, const Matrix& measurements,
                                const AnalyticalParameterProvider& offsetReader, int numObservables);
double computeAnalyticalSigmas(int sigmaIdx, const Matrix& modelOutputsScaled, const Matrix& measurements,
                               const AnalyticalParameterProvider& sigmaReader, int numObservables,
                               double epsilonAbs = 1e-12, double epsilonRel = 0.01);
// :932-987
void applyOptimalScaling(int scalingIdx, double scalingLin, Matrix& modelOutputs,
                         const AnalyticalParameterProvider& scalingReader, int numObservables);
void applyOptimalOffset(int offsetIdx, double offsetLin, Matrix& modelOutputs,
                        const AnalyticalParameterProvider& offsetReader, int numObservables);
// :989-1029
std::vector<double> spliceParameters(const std::vector<double>& reducedParameters,
                                     const std::vector<int>& proportionalityFactorIndices,
                                     const std::vector<int>& offsetParameterIndices,
                                     const std::vector<int>& sigmaParameterIndices,
                                     const std::vector<double>& scalingFactors,
                                     const std::vector<double>& offsetParameters,
                                     const std::vector<double>& sigmaParameters);
// :649-669
void fillFilteredParams(const std::vector<double>& valuesToFilter, const std::vector<int>& sortedIndicesToExclude,
                        double* result);
// :1032-1046, :1095-1134
double computeNegLogLikelihood(const Matrix& measurements, const Matrix& modelOutputsScaled, const Matrix& sigmas);
double computeNegLogLikelihood(const std::vector<double>& measurements, const std::vector<double>& modelOutputsScaled,
                               const std::vector<double>& sigmas);

/** HierarchicalOptimizationWrapper::evaluate, :142-238 (+ :280-309, :443-492) */
int hierarchicalEvaluate(const Problem& problem, const HierarchicalSpec& spec, const double* reducedParameters,
                         double& fval, double* gradient /*nullable*/, double* fullParametersOut /*nullable*/,
                         int nthreads);

}  // namespace pb200_oracle
#endif
