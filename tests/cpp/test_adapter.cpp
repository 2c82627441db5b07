// Drives the CUDA hot path through the reference-facing C++ interface: a
// parpe::SummedGradientFunction<int> bound to all conditions by
// parpe::SummedGradientFunctionGradientFunctionAdapter (what OptimizationReporter / the optimizer
// adapters call), the single-index overload, a minibatch subset (what MinibatchOptimizer passes),
// cost-only evaluation (gradient.empty()), and the hierarchical wrapper.  Prints JSON; pytest
// compares with the Python path and the oracle.
#include <cmath>
#include <cstdio>
#include <memory>

#include "../../parpe_b200/include/parpe_b200/b200SummedGradientFunction.h"

using namespace parpe;

static B200ProblemTables steadystate(int C) {
    B200ProblemTables T;
    T.modelName = "steadystate";
    T.numConditions = C;
    T.np = 5; T.nk = 4; T.ny = 3;
    T.numOptimizationParameters = 5;
    const int nt = 20;
    for (int c = 0; c <= C; ++c) T.tOffsets.push_back((int64_t)c * nt);
    for (int c = 0; c < C; ++c) {
        for (int i = 0; i < nt; ++i) T.t.push_back(std::pow(10.0, -5.0 + 6.0 * i / (nt - 1)));
        for (int i = 0; i < nt * 3; ++i) {
            T.y.push_back(0.4 + 0.01 * ((i + 7 * c) % 5));
            T.sigma.push_back(0.1);
        }
        const double k[4] = {0.1, 0.4, 0.7, 1.0 + 0.1 * c};
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

static void print_vec(const char* name, const std::vector<double>& v) {
    std::printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
    std::printf("]");
}

int main() {
    try {
        const int C = 4;
        auto fun = std::make_unique<B200SummedGradientFunction>(steadystate(C));
        B200SummedGradientFunction* raw = fun.get();
        std::vector<double> theta = {0.0, std::log10(0.5), std::log10(0.4), std::log10(2.0), -1.0};
        std::vector<double> grad(5), grad1(5), gradb(5), none;
        double fval = 0, fval1 = 0, fvalb = 0, fval0 = 0, cpu = 0;
        // the adapter binds the full index list, as MultiConditionProblem does (multiConditionProblem.cpp:41-69)
        SummedGradientFunctionGradientFunctionAdapter<int> adapter(std::move(fun), raw->getTrainingData());
        GradientFunction& gf = adapter;
        auto st = gf.evaluate(gsl::span<const double>(theta), fval, gsl::span<double>(grad), nullptr, &cpu);
        auto st0 = gf.evaluate(gsl::span<const double>(theta), fval0, gsl::span<double>());
        SummedGradientFunction<int>& sf = adapter;
        auto st1 = sf.evaluate(gsl::span<const double>(theta), 2, fval1, gsl::span<double>(grad1), nullptr, nullptr);
        auto stb = sf.evaluate(gsl::span<const double>(theta), std::vector<int>{3, 1}, fvalb,
                               gsl::span<double>(gradb), nullptr, nullptr);
        std::vector<std::vector<double>> y, s;
        raw->getModelOutputsAndSigmas(gsl::span<const double>(theta), y, s, nullptr, nullptr);
        std::printf("{\"status\": [%d, %d, %d, %d], \"numParameters\": %d, \"fval\": %.17g, \"fval_costonly\": %.17g, "
                    "\"fval_single\": %.17g, \"fval_batch\": %.17g, \"y00\": %.17g, \"n_outputs\": %zu, ",
                    (int)st, (int)st0, (int)st1, (int)stb, gf.numParameters(), fval, fval0, fval1, fvalb, y[0][0],
                    y[3].size());
        print_vec("grad", grad);
        std::printf(", ");
        print_vec("grad_single", grad1);
        std::printf(", ");
        print_vec("grad_batch", gradb);
        // wrong size: reference throws (gsl Expects / ParPEException)
        bool threw = false;
        try {
            std::vector<double> bad(4);
            double f;
            gf.evaluate(gsl::span<const double>(bad), f, gsl::span<double>());
        } catch (std::exception const&) { threw = true; }
        std::printf(", \"throws_on_size_mismatch\": %s}\n", threw ? "true" : "false");
        return 0;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
