// No Python, no torch: drives TWO (or more) GPUs of this process through the reference-facing C++
// interface -- parpe::B200MultiGpuSummedGradientFunction behind the reference's
// SummedGradientFunctionGradientFunctionAdapter -- and compares the global result with the one-GPU
// result of the same problem.  The collective is the library's own ncclAllReduce (pb200_nccl_init).
// Usage: test_multi_gpu [n_devices]    (prints JSON; exit code 0 iff the results agree)
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
        const double k[4] = {0.1, 0.4, 0.7, 1.0 + 0.01 * (c % 37)};
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

int main(int argc, char** argv) {
    try {
        const int ndev = argc > 1 ? std::atoi(argv[1]) : 2;
        if (pb200_device_count() < ndev) {
            std::printf("{\"skipped\": \"needs %d GPUs, %d visible\"}\n", ndev, pb200_device_count());
            return 0;
        }
        const int C = 301;  // odd: uneven shares
        const B200ProblemTables tables = steadystate(C);
        std::vector<double> theta = {0.0, std::log10(0.5), std::log10(0.4), std::log10(2.0), -1.0};
        // one GPU
        B200SummedGradientFunction single(tables, 0);
        std::vector<double> g1(5), gm(5), gb1(5), gbm(5);
        double f1 = 0, fm = 0, fb1 = 0, fbm = 0, f0 = 0, f01 = 0;
        auto st1 = single.evaluate(gsl::span<const double>(theta), single.getTrainingData(), f1, gsl::span<double>(g1),
                                   nullptr, nullptr);
        // (cost only integrates without the sensitivities in the error test: its own reference value)
        single.evaluate(gsl::span<const double>(theta), single.getTrainingData(), f01, gsl::span<double>(), nullptr, nullptr);
        // several GPUs behind the reference's adapter
        std::vector<int> devices;
        for (int d = 0; d < ndev; ++d) devices.push_back(d);
        auto multi = std::make_unique<B200MultiGpuSummedGradientFunction>(tables, devices);
        B200MultiGpuSummedGradientFunction* raw = multi.get();
        SummedGradientFunctionGradientFunctionAdapter<int> adapter(std::move(multi), raw->getTrainingData());
        GradientFunction& gf = adapter;
        auto stm = gf.evaluate(gsl::span<const double>(theta), fm, gsl::span<double>(gm), nullptr, nullptr);
        auto st0 = gf.evaluate(gsl::span<const double>(theta), f0, gsl::span<double>());
        // a minibatch (what MinibatchOptimizer passes), with duplicates
        std::vector<int> batch = {7, 300, 12, 12, 150, 3, 299, 41, 0};
        single.evaluate(gsl::span<const double>(theta), batch, fb1, gsl::span<double>(gb1), nullptr, nullptr);
        SummedGradientFunction<int>& sf = adapter;
        sf.evaluate(gsl::span<const double>(theta), batch, fbm, gsl::span<double>(gbm), nullptr, nullptr);
        double err = std::fabs(fm - f1) / std::fabs(f1) + std::fabs(f0 - f01) / std::fabs(f01) + std::fabs(fbm - fb1) / std::fabs(fb1);
        double gscale = 0, gerr = 0;
        for (int j = 0; j < 5; ++j) gscale = std::fmax(gscale, std::fabs(g1[j]));
        for (int j = 0; j < 5; ++j) gerr = std::fmax(gerr, std::fabs(gm[j] - g1[j]) / gscale);
        for (int j = 0; j < 5; ++j) gerr = std::fmax(gerr, std::fabs(gbm[j] - gb1[j]) / gscale);
        const bool ok = st1 == functionEvaluationSuccess && stm == functionEvaluationSuccess &&
                        st0 == functionEvaluationSuccess && err < 1e-11 && gerr < 1e-11;
        std::printf("{\"devices\": %d, \"conditions\": %d, \"fval_1gpu\": %.17g, \"fval_multi\": %.17g, \"fval_costonly\": %.17g, "
                    "\"fval_batch_1gpu\": %.17g, \"fval_batch_multi\": %.17g, \"rel_err\": %.3g, \"grad_rel_err\": %.3g, "
                    "\"ok\": %s}\n", ndev, C, f1, fm, f0, fb1, fbm, err, gerr, ok ? "true" : "false");
        return ok ? 0 : 1;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
