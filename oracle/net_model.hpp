// TEST INFRASTRUCTURE (oracle) -- not part of the product path.
//
// amici::Model_ODE subclass (MATLAB-generated layout, pythonGenerated=false; cf.
// deps/AMICI/models/model_steadystate/model_steadystate.h) over the tables of a generated
// mass-action network (parpe_b200/codegen/netgen.py): the reference AMICI/CVODES runtime then
// integrates exactly the model the CUDA path evaluates from the same tables.
#ifndef PB200_ORACLE_NET_MODEL_HPP
#define PB200_ORACLE_NET_MODEL_HPP
#include <cmath>
#include <memory>
#include <string>
#include <vector>

#include <sunmatrix/sunmatrix_sparse.h>
#include "amici/model_ode.h"
#include "amici/solver_cvodes.h"

#include "../parpe_b200/csrc/net_model.cuh"
#include "../parpe_b200/csrc/team.cuh"

namespace pb200_oracle_models {
using amici::realtype;

class NetworkModel : public amici::Model_ODE {
  public:
    explicit NetworkModel(pb200::NetTables const& T)
        : amici::Model_ODE(
              amici::ModelDimensions(T.nx, T.nx, T.nx, T.nx, 0, T.np, T.nk, T.ny, T.ny, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0,
                                     {}, 0, 0, 0, T.nnz, T.nx - 1, T.nx - 1),
              amici::SimulationParameters(std::vector<realtype>(T.nk, 1.0), std::vector<realtype>(T.np, 1.0)),
              amici::SecondOrderMode::none, std::vector<realtype>(T.nx, 0.0), std::vector<int>{}),
          T_(T), vb_(T.nr), v_(T.nr), jv_(T.nnz) {}
    amici::Model* clone() const override { return new NetworkModel(*this); }
    std::string getAmiciCommit() const override { return "pb200-netgen"; }
    std::string getName() const override { return T_.name; }

    void fxdot(realtype* xdot, const realtype, const realtype* x, const realtype* p, const realtype*, const realtype*,
               const realtype*) override {
        pb200::HostTeam tm;
        pb200::net::rates(T_, tm, x, p, vb_.data(), v_.data());
        pb200::net::xdot_from_rates(T_, tm, v_.data(), xdot);
    }
    void fJSparse(SUNMatrixContent_Sparse JSparse, const realtype, const realtype* x, const realtype* p,
                  const realtype*, const realtype*, const realtype*, const realtype*) override {
        pb200::HostTeam tm;
        pb200::net::jac_values(T_, tm, x, p, jv_.data());
        for (int c = 0; c <= T_.nx; ++c) JSparse->indexptrs[c] = T_.csc_ptr[c];
        for (int e = 0; e < T_.nnz; ++e) {
            JSparse->indexvals[e] = T_.csc_row[e];
            JSparse->data[e] = jv_[T_.csc_from_csr[e]];
        }
    }
    void fdxdotdp(realtype* dxdotdp, const realtype, const realtype* x, const realtype* p, const realtype*,
                  const realtype*, const int ip, const realtype*, const realtype*) override {
        pb200::HostTeam tm;
        pb200::net::rates(T_, tm, x, p, vb_.data(), (double*)nullptr);
        for (int i = 0; i < T_.nx; ++i)
            for (int e = T_.s_ptr[i]; e < T_.s_ptr[i + 1]; ++e)
                if (T_.pidx[T_.s_reac[e]] == ip) dxdotdp[i] += T_.s_coef[e] * vb_[T_.s_reac[e]];
    }
    void fx0(realtype* x0, const realtype, const realtype*, const realtype* k) override {
        pb200::HostTeam tm;
        pb200::net::x0(T_, tm, k, x0);
    }
    void fsx0(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const int) override {}
    // re-initialisation after pre-equilibration, in the layout of the python-generated models
    // (deps/AMICI/python/sdist/amici/de_export.py:688-749): the listed states whose initial value depends on a
    // constant take that value again, their sensitivities sx0 = 0
    void fx0_fixedParameters(realtype* x0, const realtype, const realtype*, const realtype* k,
                             gsl::span<const int> reinitialization_state_idxs) override {
        for (int i : reinitialization_state_idxs)
            if (T_.x0_kidx[i] >= 0) x0[i] = T_.x0_base[i] * k[T_.x0_kidx[i]];
    }
    void fsx0_fixedParameters(realtype* sx0, const realtype, const realtype*, const realtype*, const realtype*,
                              const int, gsl::span<const int> reinitialization_state_idxs) override {
        for (int i : reinitialization_state_idxs)
            if (T_.x0_kidx[i] >= 0) sx0[i] = 0.0;
    }
    void fy(double* y, const realtype, const realtype* x, const realtype*, const realtype*, const realtype*,
            const realtype*) override {
        for (int j = 0; j < T_.ny; ++j) y[j] = pb200::net::observable(T_, j, x);
    }
    void fdydx(double* dydx, const realtype, const realtype*, const realtype*, const realtype*, const realtype*,
               const realtype*, const realtype*) override {
        for (int j = 0; j < T_.ny; ++j)
            for (int e = T_.o_ptr[j]; e < T_.o_ptr[j + 1]; ++e) dydx[j + T_.o_idx[e] * T_.ny] += T_.o_w[e];
    }
    void fdydp(double*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*, const int,
               const realtype*, const realtype*) override {}
    void fsigmay(double* sigmay, const realtype, const realtype*, const realtype*, const realtype*) override {
        for (int j = 0; j < T_.ny; ++j) sigmay[j] = T_.sigma0[j];
    }
    void fdsigmaydp(double*, const realtype, const realtype*, const realtype*, const realtype*, const int) override {}
    // Gaussian objective, the formulas of the reference's generated Jy/dJydy/dJydsigma
    // (deps/AMICI/models/model_steadystate/Jy.cpp:16-26)
    void fJy(double* nllh, const int iy, const realtype*, const realtype*, const double* y, const double* sigmay,
             const double* my) override {
        nllh[0] = std::log((sigmay[iy] * sigmay[iy]) * 3.141592653589793 * 2.0) * 5.0E-1 +
                  1.0 / (sigmay[iy] * sigmay[iy]) * std::pow(my[iy] - y[iy], 2.0) * 5.0E-1;
    }
    void fdJydy(double* dJydy, const int iy, const realtype*, const realtype*, const double* y, const double* sigmay,
                const double* my) override {
        dJydy[iy] = 1.0 / (sigmay[iy] * sigmay[iy]) * (my[iy] * 2.0 - y[iy] * 2.0) * -5.0E-1;
    }
    void fdJydsigma(double* dJydsigma, const int iy, const realtype*, const realtype*, const double* y,
                    const double* sigmay, const double* my) override {
        dJydsigma[iy] = 1.0 / (sigmay[iy] * sigmay[iy] * sigmay[iy]) * std::pow(my[iy] - y[iy], 2.0) * -1.0 + 1.0 / sigmay[iy];
    }
    void fw(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*,
            const realtype*, const realtype*, bool) override {}
    void fdwdx(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*,
               const realtype*, const realtype*, const realtype*, bool) override {}
    void fdwdp(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*,
               const realtype*, const realtype*, const realtype*, const realtype*, const realtype*, bool) override {}
    void froot(realtype*, const realtype, const realtype*, const realtype*, const realtype*, const realtype*,
               const realtype*) override {}

  private:
    pb200::NetTables T_;
    std::vector<double> vb_, v_, jv_;
};
}  // namespace pb200_oracle_models
#endif
