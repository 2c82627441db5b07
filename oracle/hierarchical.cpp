// TEST INFRASTRUCTURE (oracle) -- not part of the product path.  See hierarchical.hpp.
#include "hierarchical.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace pb200_oracle {

std::vector<int> AnalyticalParameterProvider::getConditionsForParameter(int parameterIndex) const {
    std::vector<int> result;
    for (auto const& kv : mapping[parameterIndex])
        result.push_back(kv.first);
    return result;
}

const std::vector<int>& AnalyticalParameterProvider::getObservablesForParameter(int parameterIndex,
                                                                                int conditionIdx) const {
    return mapping[parameterIndex].at(conditionIdx);
}

AnalyticalParameterProvider AnalyticalParameterProvider::fromRows(const int32_t* idx, int n, const int32_t* rows,
                                                                  int n_rows) {
    AnalyticalParameterProvider p;
    p.optimizationParameterIndices.assign(idx, idx + n);
    p.mapping.resize(n);
    for (int i = 0; i < n_rows; ++i) {
        int scalingIdx = rows[i * 3 + 0], conditionIdx = rows[i * 3 + 1], observableIdx = rows[i * 3 + 2];
        p.mapping.at(scalingIdx)[conditionIdx].push_back(observableIdx);
    }
    return p;
}

HierarchicalSpec HierarchicalSpec::fromDesc(const pb200_hier_desc& hd, const Tables& tab) {
    HierarchicalSpec s;
    s.scaling = AnalyticalParameterProvider::fromRows(hd.scaling_idx, hd.n_scaling, hd.scaling_map, hd.n_scaling_rows);
    s.offset = AnalyticalParameterProvider::fromRows(hd.offset_idx, hd.n_offset, hd.offset_map, hd.n_offset_rows);
    s.sigma = AnalyticalParameterProvider::fromRows(hd.sigma_idx, hd.n_sigma, hd.sigma_map, hd.n_sigma_rows);
    s.numObservables = tab.ny;
    s.numConditions = tab.n_conditions;
    s.nOpt = tab.n_opt;
    s.pscale_opt = tab.pscale_opt;
    return s;
}

double getDefaultScalingFactor(int scaling) {
    switch (scaling) {
    case 0: return 1.0;
    case 2: return 0.0;
    default: throw std::invalid_argument("Parameter scaling must be log10 or none.");
    }
}

double getDefaultOffsetParameter(int scaling) {
    switch (scaling) {
    case 0: return 0.0;
    case 2: return -std::numeric_limits<double>::infinity();  // quirk Q6
    default: throw std::invalid_argument("Parameter scaling must be log10 or none.");
    }
}

double computeAnalyticalScalings(int scalingIdx, const Matrix& modelOutputsUnscaled, const Matrix& measurements,
                                 const AnalyticalParameterProvider& scalingReader, int numObservables) {
    double enumerator = 0.0, denominator = 0.0;
    for (int conditionIdx : scalingReader.getConditionsForParameter(scalingIdx)) {
        auto const& dependentObservables = scalingReader.getObservablesForParameter(scalingIdx, conditionIdx);
        int numTimepoints = (int)measurements[conditionIdx].size() / numObservables;
        for (int observableIdx : dependentObservables) {
            for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx) {
                double mes = measurements[conditionIdx][observableIdx + timeIdx * numObservables];
                if (!std::isnan(mes)) {
                    double sim = modelOutputsUnscaled[conditionIdx][observableIdx + timeIdx * numObservables];
                    if (sim < 0 && sim > -1e-18)
                        sim = 0.0;
                    enumerator += sim * mes;
                    denominator += sim * sim;
                }
            }
        }
    }
    if (denominator == 0.0)
        return 1.0;
    return enumerator / denominator;
}

double computeAnalyticalOffsets(int offsetIdx, const Matrix& modelOutputsUnscaled, const Matrix& measurements,
                                const AnalyticalParameterProvider& offsetReader, int numObservables) {
    double enumerator = 0.0, denominator = 0.0;
    for (int conditionIdx : offsetReader.getConditionsForParameter(offsetIdx)) {
        auto const& dependentObservables = offsetReader.getObservablesForParameter(offsetIdx, conditionIdx);
        int numTimepoints = (int)measurements[conditionIdx].size() / numObservables;
        for (int observableIdx : dependentObservables) {
            for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx) {
                double mes = measurements[conditionIdx][observableIdx + timeIdx * numObservables];
                if (!std::isnan(mes)) {
                    double sim = modelOutputsUnscaled[conditionIdx][observableIdx + timeIdx * numObservables];
                    enumerator += mes - sim;
                    denominator += 1.0;
                }
            }
        }
    }
    if (denominator == 0.0)
        return 0.0;
    return enumerator / denominator;
}

double computeAnalyticalSigmas(int sigmaIdx, const Matrix& modelOutputsScaled, const Matrix& measurements,
                               const AnalyticalParameterProvider& sigmaReader, int numObservables, double epsilonAbs,
                               double epsilonRel) {
    double enumerator = 0.0, denominator = 0.0, maxAbsMeasurement = 0.0;
    for (int conditionIdx : sigmaReader.getConditionsForParameter(sigmaIdx)) {
        auto const& dependentObservables = sigmaReader.getObservablesForParameter(sigmaIdx, conditionIdx);
        int numTimepoints = (int)measurements[conditionIdx].size() / numObservables;
        for (int observableIdx : dependentObservables) {
            if (observableIdx >= numObservables)
                throw std::out_of_range("computeAnalyticalSigmas: Invalid observableIdx >= numObservables.");
            for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx) {
                const int flat_index = observableIdx + timeIdx * numObservables;
                double mes = measurements[conditionIdx][flat_index];
                if (!std::isnan(mes)) {
                    double scaledSim = modelOutputsScaled[conditionIdx][flat_index];
                    enumerator += (mes - scaledSim) * (mes - scaledSim);
                    denominator += 1.0;
                    maxAbsMeasurement = std::max(maxAbsMeasurement, std::abs(mes));
                }
            }
        }
    }
    double sigma = std::sqrt(enumerator / denominator);
    epsilonAbs = std::max(epsilonRel * maxAbsMeasurement, epsilonAbs);
    if (sigma < epsilonAbs)
        return epsilonAbs;
    return sigma;
}

void applyOptimalScaling(int scalingIdx, double scalingLin, Matrix& modelOutputs,
                         const AnalyticalParameterProvider& scalingReader, int numObservables) {
    for (int conditionIdx : scalingReader.getConditionsForParameter(scalingIdx)) {
        int numTimepoints = (int)modelOutputs[conditionIdx].size() / numObservables;
        for (int observableIdx : scalingReader.getObservablesForParameter(scalingIdx, conditionIdx)) {
            if (observableIdx >= numObservables)
                throw std::out_of_range("applyOptimalScaling: Invalid observableIdx >= numObservables.");
            for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx)
                modelOutputs[conditionIdx][observableIdx + timeIdx * numObservables] *= scalingLin;
        }
    }
}

void applyOptimalOffset(int offsetIdx, double offsetLin, Matrix& modelOutputs,
                        const AnalyticalParameterProvider& offsetReader, int numObservables) {
    for (int conditionIdx : offsetReader.getConditionsForParameter(offsetIdx)) {
        int numTimepoints = (int)modelOutputs[conditionIdx].size() / numObservables;
        for (int observableIdx : offsetReader.getObservablesForParameter(offsetIdx, conditionIdx)) {
            if (observableIdx >= numObservables)
                throw std::out_of_range("applyOptimalOffset: Invalid observableIdx >= numObservables.");
            for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx)
                modelOutputs[conditionIdx][observableIdx + timeIdx * numObservables] += offsetLin;
        }
    }
}

std::vector<double> spliceParameters(const std::vector<double>& reducedParameters,
                                     const std::vector<int>& proportionalityFactorIndices,
                                     const std::vector<int>& offsetParameterIndices,
                                     const std::vector<int>& sigmaParameterIndices,
                                     const std::vector<double>& scalingFactors,
                                     const std::vector<double>& offsetParameters,
                                     const std::vector<double>& sigmaParameters) {
    std::vector<double> fullParameters(reducedParameters.size() + scalingFactors.size() + offsetParameters.size() +
                                       sigmaParameters.size());
    size_t idxScaling = 0, idxOffset = 0, idxSigma = 0, idxRegular = 0;
    for (int i = 0; i < (int)fullParameters.size(); ++i) {
        if (idxScaling < proportionalityFactorIndices.size() && proportionalityFactorIndices[idxScaling] == i)
            fullParameters[i] = scalingFactors.at(idxScaling++);
        else if (idxOffset < offsetParameterIndices.size() && offsetParameterIndices[idxOffset] == i)
            fullParameters[i] = offsetParameters.at(idxOffset++);
        else if (idxSigma < sigmaParameterIndices.size() && sigmaParameterIndices[idxSigma] == i)
            fullParameters[i] = sigmaParameters.at(idxSigma++);
        else if (idxRegular < reducedParameters.size())
            fullParameters[i] = reducedParameters[idxRegular++];
        else
            throw std::logic_error("spliceParameters: index lists inconsistent");
    }
    if (idxScaling != proportionalityFactorIndices.size() || idxOffset != offsetParameterIndices.size() ||
        idxSigma != sigmaParameterIndices.size() || idxRegular != reducedParameters.size())
        throw std::logic_error("spliceParameters: postcondition violated");
    return fullParameters;
}

void fillFilteredParams(const std::vector<double>& valuesToFilter, const std::vector<int>& sortedIndicesToExclude,
                        double* result) {
    size_t nextFilterIdx = 0, resultIdx = 0;
    for (int i = 0; i < (int)valuesToFilter.size(); ++i) {
        if (nextFilterIdx < sortedIndicesToExclude.size() && sortedIndicesToExclude[nextFilterIdx] == i)
            ++nextFilterIdx;
        else
            result[resultIdx++] = valuesToFilter[i];
    }
}

double computeNegLogLikelihood(const std::vector<double>& measurements, const std::vector<double>& modelOutputsScaled,
                               const std::vector<double>& sigmas) {
    double nllh = 0.0;
    for (size_t i = 0; i < measurements.size(); ++i) {
        double mes = measurements[i];
        if (!std::isnan(mes)) {
            double sim = modelOutputsScaled[i];
            double sigmaSquared = sigmas[i] * sigmas[i];
            if (std::isnan(sim) || std::isnan(sigmaSquared) || sigmaSquared < 0.0)
                return std::numeric_limits<double>::quiet_NaN();
            double diff = mes - sim;
            diff *= diff;
            nllh += std::log(2.0 * M_PI * sigmaSquared) + diff / sigmaSquared;
        }
    }
    nllh /= 2.0;
    return nllh;
}

double computeNegLogLikelihood(const Matrix& measurements, const Matrix& modelOutputsScaled, const Matrix& sigmas) {
    double nllh = 0.0;
    for (size_t conditionIdx = 0; conditionIdx < measurements.size(); ++conditionIdx) {
        nllh += computeNegLogLikelihood(measurements[conditionIdx], modelOutputsScaled[conditionIdx], sigmas[conditionIdx]);
        if (std::isnan(nllh))
            return nllh;
    }
    return nllh;
}

int hierarchicalEvaluate(const Problem& problem, const HierarchicalSpec& spec, const double* reducedParameters,
                         double& fval, double* gradient, double* fullParametersOut, int nthreads) {
    auto const& tab = problem.tables();
    const int nOuter = spec.numOuter();
    std::vector<double> reduced(reducedParameters, reducedParameters + nOuter);
    auto const& sIdx = spec.scaling.optimizationParameterIndices;
    auto const& oIdx = spec.offset.optimizationParameterIndices;
    auto const& gIdx = spec.sigma.optimizationParameterIndices;

    // getUnscaledModelOutputsAndSigmas, :280-309 with the defaults of :240-278
    std::vector<double> scalingDummy(sIdx.size()), offsetDummy(oIdx.size()), sigmaDummy(gIdx.size());
    for (size_t i = 0; i < sIdx.size(); ++i) scalingDummy[i] = getDefaultScalingFactor(spec.pscale_opt[sIdx[i]]);
    for (size_t i = 0; i < oIdx.size(); ++i) offsetDummy[i] = getDefaultOffsetParameter(spec.pscale_opt[oIdx[i]]);
    for (size_t i = 0; i < gIdx.size(); ++i) sigmaDummy[i] = getDefaultScalingFactor(spec.pscale_opt[gIdx[i]]);
    auto fullParameters = spliceParameters(reduced, sIdx, oIdx, gIdx, scalingDummy, offsetDummy, sigmaDummy);

    Matrix modelOutputs, modelSigmas;
    if (problem.getModelOutputsAndSigmas(fullParameters.data(), modelOutputs, modelSigmas, nthreads) != 0)
        return 1;

    // getAllMeasurements
    Matrix measurements(tab.n_conditions);
    for (int c = 0; c < tab.n_conditions; ++c)
        measurements[c].assign(tab.y.begin() + tab.t_offsets[c] * tab.ny, tab.y.begin() + tab.t_offsets[c + 1] * tab.ny);

    // :311-333, :355-401
    std::vector<double> scalings(sIdx.size()), offsets(oIdx.size()), sigmas(gIdx.size());
    for (size_t i = 0; i < sIdx.size(); ++i)
        scalings[i] = getScaledParameter(
            computeAnalyticalScalings((int)i, modelOutputs, measurements, spec.scaling, spec.numObservables),
            spec.pscale_opt[sIdx[i]]);
    for (size_t i = 0; i < oIdx.size(); ++i)  // quirk Q5: from the unscaled outputs
        offsets[i] = getScaledParameter(
            computeAnalyticalOffsets((int)i, modelOutputs, measurements, spec.offset, spec.numObservables),
            spec.pscale_opt[oIdx[i]]);
    for (size_t i = 0; i < sIdx.size(); ++i)
        applyOptimalScaling((int)i, getUnscaledParameter(scalings[i], spec.pscale_opt[sIdx[i]]), modelOutputs,
                            spec.scaling, spec.numObservables);
    for (size_t i = 0; i < oIdx.size(); ++i)
        applyOptimalOffset((int)i, getUnscaledParameter(offsets[i], spec.pscale_opt[oIdx[i]]), modelOutputs, spec.offset,
                           spec.numObservables);
    for (size_t i = 0; i < gIdx.size(); ++i)
        sigmas[i] = getScaledParameter(
            computeAnalyticalSigmas((int)i, modelOutputs, measurements, spec.sigma, spec.numObservables),
            spec.pscale_opt[gIdx[i]]);

    fullParameters = spliceParameters(reduced, sIdx, oIdx, gIdx, scalings, offsets, sigmas);
    if (fullParametersOut)
        std::copy(fullParameters.begin(), fullParameters.end(), fullParametersOut);

    // evaluateWithOptimalParameters, :443-492
    if (gradient) {
        fval = NAN;
        std::vector<int> dataIndices(tab.n_conditions);
        for (int i = 0; i < tab.n_conditions; ++i) dataIndices[i] = i;
        std::vector<double> fullGradient(fullParameters.size());
        int status = problem.evaluate(fullParameters.data(), dataIndices, fval, fullGradient.data(), nthreads, nullptr, false);
        if (status != 0)
            return status;
        std::vector<int> analytical(sIdx);
        analytical.insert(analytical.end(), oIdx.begin(), oIdx.end());
        analytical.insert(analytical.end(), gIdx.begin(), gIdx.end());
        std::sort(analytical.begin(), analytical.end());
        fillFilteredParams(fullGradient, analytical, gradient);
    } else {
        // fillInAnalyticalSigmas, :403-441
        for (size_t sp = 0; sp < gIdx.size(); ++sp) {
            double sigmaParameterValue = getUnscaledParameter(sigmas[sp], spec.pscale_opt[gIdx[sp]]);
            for (int conditionIdx : spec.sigma.getConditionsForParameter((int)sp)) {
                int numTimepoints = (int)modelSigmas[conditionIdx].size() / spec.numObservables;
                for (int observableIdx : spec.sigma.getObservablesForParameter((int)sp, conditionIdx))
                    for (int timeIdx = 0; timeIdx < numTimepoints; ++timeIdx)
                        modelSigmas[conditionIdx][observableIdx + timeIdx * spec.numObservables] = sigmaParameterValue;
            }
        }
        fval = computeNegLogLikelihood(measurements, modelOutputs, modelSigmas);
    }
    return std::isfinite(fval) ? 0 : 1;
}

}  // namespace pb200_oracle
