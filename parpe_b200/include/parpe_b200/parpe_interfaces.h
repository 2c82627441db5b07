// parpe_interfaces.h -- the abstract interfaces of parPE that the B200 hot path sits behind.
//
// When building inside parPE, define PB200_USE_PARPE_HEADERS: the reference's own
// <parpecommon/functions.h> (include/parpecommon/functions.h:16-198) is included and the
// adapter classes of b200SummedGradientFunction.h derive from the reference's types, so
// OptimizationReporter, the optimizer adapters, MinibatchOptimizer and
// HierarchicalOptimizationProblemWrapper see ordinary parpe:: objects.
//
// For a standalone build (tests on a box without the parPE tree) the few declarations this
// path needs are provided here under the same names and signatures: FunctionEvaluationStatus,
// GradientFunction, SummedGradientFunction<T>, SummedGradientFunctionGradientFunctionAdapter<T>,
// an opaque Logger, and a pointer+length gsl::span.
#ifndef PB200_PARPE_INTERFACES_H
#define PB200_PARPE_INTERFACES_H

#if defined(PB200_USE_PARPE_HEADERS)
#include <parpecommon/functions.h>
#else

#include <cstddef>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace gsl {
// pointer + length view; only what the interface signatures use
template <class T>
class span {
  public:
    span() = default;
    span(T* p, std::size_t n) : p_(p), n_(n) {}
    template <class V, class = decltype(std::declval<V&>().data())>
    span(V& v) : p_(v.data()), n_(v.size()) {}
    T* data() const { return p_; }
    std::size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T& operator[](std::size_t i) const { return p_[i]; }
    T* begin() const { return p_; }
    T* end() const { return p_ + n_; }

  private:
    T* p_ = nullptr;
    std::size_t n_ = 0;
};
}  // namespace gsl

namespace parpe {

class Logger;  // opaque here; the hot path only passes the pointer through

enum FunctionEvaluationStatus { functionEvaluationSuccess, functionEvaluationFailure };

// cost function with gradient (reference: functions.h:26-56)
class GradientFunction {
  public:
    virtual ~GradientFunction() = default;
    virtual FunctionEvaluationStatus evaluate(gsl::span<double const> parameters, double& fval,
                                              gsl::span<double> gradient) const {
        return evaluate(parameters, fval, gradient, nullptr, nullptr);
    }
    virtual FunctionEvaluationStatus evaluate(gsl::span<double const> parameters, double& fval,
                                              gsl::span<double> gradient, Logger* logger, double* cpuTime) const = 0;
    virtual int numParameters() const = 0;
    virtual std::vector<std::string> getParameterIds() const = 0;
};

// cost function that is a sum over data sets (reference: functions.h:66-118)
template <typename T>
class SummedGradientFunction {
  public:
    virtual ~SummedGradientFunction() = default;
    virtual FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, T dataset, double& fval,
                                              gsl::span<double> gradient, Logger* logger, double* cpuTime) const = 0;
    virtual FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, std::vector<T> datasets,
                                              double& fval, gsl::span<double> gradient, Logger* logger,
                                              double* cpuTime) const = 0;
    virtual int numParameters() const = 0;
    virtual std::vector<std::string> getParameterIds() const = 0;
};

// binds a SummedGradientFunction to a fixed list of data sets (reference: functions.h:127-198)
template <typename T>
class SummedGradientFunctionGradientFunctionAdapter : public GradientFunction, public SummedGradientFunction<T> {
  public:
    SummedGradientFunctionGradientFunctionAdapter(std::unique_ptr<SummedGradientFunction<T>> gradFun,
                                                  std::vector<T> datasets)
        : grad_fun_(std::move(gradFun)), datasets_(std::move(datasets)) {}
    using GradientFunction::evaluate;
    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, double& fval, gsl::span<double> gradient,
                                      Logger* logger = nullptr, double* cpuTime = nullptr) const override {
        return grad_fun_->evaluate(parameters, datasets_, fval, gradient, logger, cpuTime);
    }
    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, T dataset, double& fval,
                                      gsl::span<double> gradient, Logger* logger, double* cpuTime) const override {
        return grad_fun_->evaluate(parameters, dataset, fval, gradient, logger, cpuTime);
    }
    FunctionEvaluationStatus evaluate(gsl::span<const double> parameters, std::vector<T> datasets, double& fval,
                                      gsl::span<double> gradient, Logger* logger, double* cpuTime) const override {
        return grad_fun_->evaluate(parameters, std::move(datasets), fval, gradient, logger, cpuTime);
    }
    int numParameters() const override { return grad_fun_->numParameters(); }
    std::vector<std::string> getParameterIds() const override { return grad_fun_->getParameterIds(); }
    SummedGradientFunction<T>* getWrappedFunction() const { return grad_fun_.get(); }

  private:
    std::unique_ptr<SummedGradientFunction<T>> grad_fun_;
    std::vector<T> datasets_;
};

}  // namespace parpe
#endif  // PB200_USE_PARPE_HEADERS
#endif
