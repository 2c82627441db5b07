// Drives the CUDA hot path with the REFERENCE'S OWN, UNMODIFIED parpe::MinibatchOptimizer<int>
// (include/parpeoptimization/minibatchOptimization.h:396-538 and
// src/parpeoptimization/minibatchOptimization.cpp, compiled in place from /root/reference by
// tests/cpp/build_minibatch.sh together with the reference's logging.cpp / misc.cpp /
// parpeException.cpp): three epochs of mini-batch gradient descent on the steady-state example
// through parpe::B200SummedGradientFunction, which here derives from the reference's own
// parpe::SummedGradientFunction<int> (PB200_USE_PARPE_HEADERS).  SURVEY.md 7.2's minimum slice.
//
// HDF5 is absent from this image: the reference headers are parsed against the type stubs in
// tests/cpp/stubs/, the optimizer runs without reporter / result writer (reporter == nullptr is a
// supported argument of MinibatchOptimizer::optimize), and the two non-virtual result-writer members
// its template body names are defined below as never-called stubs so that the program links.
#include <parpeoptimization/minibatchOptimization.h>

#include <cmath>
#include <cstdio>

#include "../../parpe_b200/include/parpe_b200/b200SummedGradientFunction.h"

namespace parpe {
// never called (reporter == nullptr); see the header comment
void OptimizationResultWriter::setLoggingEachIteration(bool) { std::abort(); }
void OptimizationResultWriter::setLoggingEachFunctionEvaluation(bool, bool) { std::abort(); }
}  // namespace parpe

using namespace parpe;

static B200ProblemTables steadystate(int C) {
    B200ProblemTables T;
    T.modelName = "steadystate";
    T.numConditions = C;
    T.np = 5; T.nk = 4; T.ny = 3;
    T.numOptimizationParameters = 5;
    const int nt = 20;
    // "measurements": the model's own steady state (the reference's golden vector,
    // examples/parpeamici/steadystate/exampleSteadystateScaledTest.h:40-48) at the late time points, a
    // smooth approach before -- positive pseudo-data, enough for a descent direction to exist
    const double xss[3] = {0.456644592142075, 0.437977375496898, 0.033333333333333};
    for (int c = 0; c <= C; ++c) T.tOffsets.push_back((int64_t)c * nt);
    for (int c = 0; c < C; ++c) {
        for (int i = 0; i < nt; ++i) T.t.push_back(std::pow(10.0, -5.0 + 6.0 * i / (nt - 1)));
        for (int i = 0; i < nt; ++i)
            for (int o = 0; o < 3; ++o) {
                const double t = T.t[(size_t)c * nt + i];
                T.y.push_back(0.1 + (xss[o] - 0.1) * (1.0 - std::exp(-t)) * (1.0 + 0.02 * ((i + o + c) % 3 - 1)));
                T.sigma.push_back(0.1);
            }
        const double k[4] = {0.1, 0.4, 0.7, 1.0};
        T.k.insert(T.k.end(), k, k + 4);
        for (int i = 0; i < 5; ++i) {
            T.mapping.push_back(i);
            T.pscaleSim.push_back(PB200_SCALE_LOG10);
        }
    }
    T.pscaleOpt.assign(5, PB200_SCALE_LOG10);
    T.parameterIds = {"p1", "p2", "p3", "p4", "p5"};
    T.atol = 1e-12;
    return T;
}

int main() {
    try {
        if (pb200_device_count() < 1) {
            std::printf("{\"skipped\": \"no CUDA device\"}\n");
            return 0;
        }
        const int C = 12;
        B200SummedGradientFunction fun(steadystate(C));
        const std::vector<int> data = fun.getTrainingData();
        // start away from the nominal parameters {1, 0.5, 0.4, 2, 0.1}
        std::vector<double> start = {0.3, -0.6, -0.1, 0.0, -0.7};
        std::vector<double> lower(5, -3.0), upper(5, 3.0);
        double cost0 = 0.0;
        fun.evaluate(gsl::span<const double>(start), data, cost0, gsl::span<double>(), nullptr, nullptr);
        cost0 /= C;

        MinibatchOptimizer<int> mb;  // the reference's optimizer, default updaters (Adam, learning-rate decay)
        mb.maxEpochs = 3;
        mb.batchSize = 4;
        mb.parameterUpdater = std::make_unique<ParameterUpdaterAdam>();
        auto [status, cost, parameters] = mb.optimize(fun, gsl::span<const int>(data), gsl::span<const double>(start),
                                                      gsl::span<const double>(lower), gsl::span<const double>(upper),
                                                      nullptr, nullptr);
        std::printf("{\"exit_status\": %d, \"epochs\": %d, \"batch_size\": %d, \"cost_start\": %.17g, \"cost_final\": %.17g, "
                    "\"parameters\": [%.17g, %.17g, %.17g, %.17g, %.17g], \"ok\": %s}\n",
                    status, mb.maxEpochs, mb.batchSize, cost0, cost, parameters[0], parameters[1], parameters[2], parameters[3],
                    parameters[4], (std::isfinite(cost) && cost < cost0) ? "true" : "false");
        return (std::isfinite(cost) && cost < cost0) ? 0 : 1;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
