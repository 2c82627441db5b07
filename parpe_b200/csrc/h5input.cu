// h5input.cu -- the data-provider role: parPE's HDF5 parameter-estimation input file -> the tables of
// pb200_problem_desc (host code only; no CUDA in this file).
//
// What it replaces in the reference: MultiConditionDataProviderHDF5's constructor and getters
// (src/parpeamici/multiConditionDataProvider.cpp:42-60 dataset paths; :84-102 mapping column of a condition;
// :194-206 pscaleSimulation row; :224-250 simulationConditions -> column of /fixedParameters/k; :302-321
// /measurements/{t,y,ysigma}/<i>; parameterOverrides :126-169) and AnalyticalParameterHdf5Reader
// (hierarchicalOptimizationAnalyticalParameterProvider.cpp: /{scaling,offset,sigma}ParameterIndices and
// ...MapToObservables), reading the layout of doc/hdf5.md:1-172 once, at create time.  The file is parsed by
// h5min.h (no HDF5 library in this image).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/parpe_b200.h"
#include "h5min.h"

namespace pb200 {
int fail_message(int code, const std::string& msg);  // parpe_b200.cu: sets pb200_last_error
}

using pb200::H5Error;
using pb200::H5File;

struct pb200_h5 {
    std::unique_ptr<H5File> f;
};

struct pb200_h5_problem {
    std::string model_name;
    std::vector<int64_t> t_offsets;
    std::vector<double> t, y, sigma, k, overrides;
    std::vector<int32_t> mapping, pscale_sim, pscale_opt;
    std::vector<int32_t> scaling_idx, offset_idx, sigma_idx, scaling_map, offset_map, sigma_map;
    // pre-equilibration (simulationConditions column 0, reinitializationIndices)
    // /parameters/{lowerBound, upperBound, nominalValues, parameterNames} (what MultiConditionProblem's
    // fillParametersMin / fillParametersMax and getProblemParameterIds read); empty when the file has none
    std::vector<double> lower, upper, nominal;
    std::string names;  // newline separated
    std::vector<int32_t> has_preeq, reinit_idx;
    std::vector<double> k_preeq;
    std::vector<int64_t> reinit_off;
    bool any_preeq = false;
    pb200_problem_desc desc{};
    pb200_hier_desc hier{};
    bool has_hier = false;
};

namespace {
template <class F>
int guarded(F&& fn) {
    try {
        return fn();
    } catch (std::exception const& e) {
        return pb200::fail_message(PB200_ERR_ARG, e.what());
    }
}

std::vector<int32_t> to_i32(const std::vector<int64_t>& v) { return std::vector<int32_t>(v.begin(), v.end()); }
}  // namespace

extern "C" {

int pb200_h5_open(const char* path, pb200_h5** out) {
    if (!path || !out) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto h = std::make_unique<pb200_h5>();
        h->f = std::make_unique<H5File>(path);
        *out = h.release();
        return 0;
    });
}
void pb200_h5_close(pb200_h5* h) { delete h; }

int pb200_h5_exists(pb200_h5* h, const char* path) {
    if (!h || !path) return 0;
    try {
        return h->f->exists(path) ? 1 : 0;
    } catch (std::exception const&) {
        return 0;
    }
}

int pb200_h5_dataset_info(pb200_h5* h, const char* path, int* rank, int64_t* dims, int* type_class, int* type_size) {
    if (!h || !path) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    return guarded([&] {
        const H5File::Dataset d = h->f->info(path);
        if (rank) *rank = d.rank;
        if (dims)
            for (int i = 0; i < d.rank && i < 8; ++i) dims[i] = (int64_t)d.dims[i];
        if (type_class) *type_class = d.type_class;
        if (type_size) *type_size = d.type_size;
        return 0;
    });
}

int64_t pb200_h5_read_double(pb200_h5* h, const char* path, double* out, int64_t capacity) {
    if (!h || !path) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    int64_t n = 0;
    const int rc = guarded([&] {
        const std::vector<double> v = h->f->read_double(path);
        n = (int64_t)v.size();
        if (out) {
            if (capacity < n) throw H5Error(std::string(path) + ": output buffer too small");
            std::memcpy(out, v.data(), sizeof(double) * v.size());
        }
        return 0;
    });
    return rc < 0 ? rc : n;
}

int64_t pb200_h5_read_raw(pb200_h5* h, const char* path, char* out, int64_t capacity) {
    if (!h || !path) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    int64_t n = 0;
    const int rc = guarded([&] {
        const std::vector<uint8_t> v = h->f->read_bytes(path);
        n = (int64_t)v.size();
        if (out) {
            if (capacity < n) throw H5Error(std::string(path) + ": output buffer too small");
            std::memcpy(out, v.data(), v.size());
        }
        return 0;
    });
    return rc < 0 ? rc : n;
}

int pb200_h5_attribute_double(pb200_h5* h, const char* obj, const char* name, double* out, int capacity) {
    if (!h || !obj || !name) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    int n = 0;
    const int rc = guarded([&] {
        const std::vector<double> v = h->f->attribute(obj, name);
        n = (int)v.size();
        if (out) {
            if (capacity < n) throw H5Error(std::string(name) + ": output buffer too small");
            std::memcpy(out, v.data(), sizeof(double) * v.size());
        }
        return 0;
    });
    return rc < 0 ? rc : n;
}

int pb200_h5_list(pb200_h5* h, const char* group, char* buf, int capacity) {
    if (!h || !group) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    int n = 0;
    const int rc = guarded([&] {
        std::string all;
        for (auto& s : h->f->list(group)) {
            all += s;
            all += '\n';
            ++n;
        }
        if (buf) {
            if ((int)all.size() + 1 > capacity) throw H5Error("list: output buffer too small");
            std::memcpy(buf, all.c_str(), all.size() + 1);
        }
        return 0;
    });
    return rc < 0 ? rc : n;
}

int pb200_h5_load_problem(const char* path, const char* model_name, pb200_h5_problem** out) {
    if (!path || !model_name || !out) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    *out = nullptr;
    int nx = 0, np = 0, nk = 0, ny = 0;
    if (pb200_model_info(model_name, &nx, &np, &nk, &ny) != 0) return PB200_ERR_MODEL;
    return guarded([&] {
        H5File f(path);
        auto P = std::make_unique<pb200_h5_problem>();
        P->model_name = model_name;
        // ---- /fixedParameters: simulationConditions [C x 2] = (preequilibration column, simulation column) of k [nk x nVec]
        const auto sc_info = f.info("/fixedParameters/simulationConditions");
        if (sc_info.rank != 2 || sc_info.dims[1] != 2) throw H5Error("/fixedParameters/simulationConditions must be [n x 2]");
        const int C = (int)sc_info.dims[0];
        const auto sc = f.read_int("/fixedParameters/simulationConditions");
        std::vector<double> K;
        uint64_t nvec = 0;
        if (nk > 0) {
            const auto k_info = f.info("/fixedParameters/k");
            if (k_info.rank != 2 || (int)k_info.dims[0] != nk) throw H5Error("/fixedParameters/k must be [nk x nConditionVectors] with the model's nk");
            nvec = k_info.dims[1];
            K = f.read_double("/fixedParameters/k");
        }
        P->k.resize((size_t)C * nk);
        P->has_preeq.assign(C, 0);
        P->k_preeq.assign((size_t)C * nk, 0.0);
        P->reinit_off.assign(C + 1, 0);
        // ragged list of the states re-initialised after pre-equilibration (getReinitializationIndices,
        // multiConditionDataProvider.cpp:497-524): a 1-D dataset of variable-length int sequences, one per simulation
        std::vector<std::vector<int64_t>> reinit;
        if (f.exists("/fixedParameters/reinitializationIndices")) {
            reinit = f.read_vlen_int("/fixedParameters/reinitializationIndices");
            if ((int)reinit.size() != C) throw H5Error("/fixedParameters/reinitializationIndices must have one entry per simulation condition");
        }
        for (int c = 0; c < C; ++c) {
            const int64_t col = sc[2 * c + 1];
            if (col < 0 || (nk > 0 && (uint64_t)col >= nvec)) throw H5Error("simulationConditions refers to a missing column of k");
            for (int i = 0; i < nk; ++i) P->k[(size_t)c * nk + i] = K[(size_t)i * nvec + col];
            // column 0: the pre-equilibration condition, -1 = none (updateFixedSimulationParameters,
            // multiConditionDataProvider.cpp:224-250)
            const int64_t pre = sc[2 * c];
            if (pre >= 0) {
                if (nk > 0 && (uint64_t)pre >= nvec) throw H5Error("simulationConditions refers to a missing column of k (pre-equilibration)");
                P->has_preeq[c] = 1;
                P->any_preeq = true;
                for (int i = 0; i < nk; ++i) P->k_preeq[(size_t)c * nk + i] = K[(size_t)i * nvec + pre];
                if (!reinit.empty())
                    for (int64_t v : reinit[c]) {
                        if (v < 0 || v >= nx) throw H5Error("reinitializationIndices: state index out of range");
                        P->reinit_idx.push_back((int32_t)v);
                    }
            }
            P->reinit_off[c + 1] = (int64_t)P->reinit_idx.size();
        }
        if (P->reinit_idx.empty()) P->reinit_idx.push_back(0);
        // ---- /measurements/{t,y,ysigma}/<c>
        P->t_offsets.assign(1, 0);
        for (int c = 0; c < C; ++c) {
            const std::string id = std::to_string(c);
            const auto tc = f.read_double("/measurements/t/" + id);
            const auto yc = f.read_double("/measurements/y/" + id);
            const auto sg = f.read_double("/measurements/ysigma/" + id);
            if (yc.size() != tc.size() * (size_t)ny || sg.size() != yc.size())
                throw H5Error("/measurements/{y,ysigma}/" + id + " must be [nt x ny] with the model's ny");
            P->t.insert(P->t.end(), tc.begin(), tc.end());
            P->y.insert(P->y.end(), yc.begin(), yc.end());
            P->sigma.insert(P->sigma.end(), sg.begin(), sg.end());
            P->t_offsets.push_back((int64_t)P->t.size());
        }
        // ---- /parameters
        P->pscale_opt = to_i32(f.read_int("/parameters/pscaleOptimization"));
        const int n_opt = (int)P->pscale_opt.size();
        P->mapping.resize((size_t)C * np);
        if (f.exists("/parameters/optimizationSimulationMapping")) {
            const auto mi = f.info("/parameters/optimizationSimulationMapping");
            if (mi.rank != 2 || (int)mi.dims[0] != np || (int)mi.dims[1] != C)
                throw H5Error("/parameters/optimizationSimulationMapping must be [np x nConditions]");
            const auto M = f.read_int("/parameters/optimizationSimulationMapping");
            for (int c = 0; c < C; ++c)
                for (int i = 0; i < np; ++i) P->mapping[(size_t)c * np + i] = (int32_t)M[(size_t)i * C + c];
        } else {  // trivial default mapping (multiConditionDataProvider.cpp:97-101)
            for (int c = 0; c < C; ++c)
                for (int i = 0; i < np; ++i) P->mapping[(size_t)c * np + i] = i;
        }
        {
            const auto si = f.info("/parameters/pscaleSimulation");
            if (si.rank != 2 || (int)si.dims[0] != C || (int)si.dims[1] != np)
                throw H5Error("/parameters/pscaleSimulation must be [nConditions x np]");
            P->pscale_sim = to_i32(f.read_int("/parameters/pscaleSimulation"));
        }
        if (f.exists("/parameters/parameterOverrides")) {
            const auto oi = f.info("/parameters/parameterOverrides");
            if (oi.rank != 2 || (int)oi.dims[0] != np || (int)oi.dims[1] != C)
                throw H5Error("/parameters/parameterOverrides must be [np x nConditions]");
            const auto O = f.read_double("/parameters/parameterOverrides");
            P->overrides.resize((size_t)C * np);
            for (int c = 0; c < C; ++c)
                for (int i = 0; i < np; ++i) P->overrides[(size_t)c * np + i] = O[(size_t)i * C + c];
        }
        // ---- bounds, nominal values and names of the optimisation parameters (optional)
        auto opt_vec = [&](const char* p, std::vector<double>& v) {
            if (!f.exists(p)) return;
            v = f.read_double(p);
            if ((int)v.size() != n_opt) throw H5Error(std::string(p) + " must have one entry per optimisation parameter");
        };
        opt_vec("/parameters/lowerBound", P->lower);
        opt_vec("/parameters/upperBound", P->upper);
        opt_vec("/parameters/nominalValues", P->nominal);
        if (f.exists("/parameters/parameterNames")) {
            const auto nm = f.read_strings("/parameters/parameterNames");
            if ((int)nm.size() != n_opt) throw H5Error("/parameters/parameterNames must have one entry per optimisation parameter");
            for (auto& s : nm) { P->names += s; P->names += '\n'; }
        }
        // ---- /amiciOptions attributes (deps/AMICI/src/hdf5.cpp readSolverSettingsFromHDF5)
        pb200_problem_desc& d = P->desc;
        d.abi_version = PB200_ABI_VERSION;
        d.model_name = P->model_name.c_str();
        d.n_conditions = C;
        d.n_opt = n_opt;
        d.np = np; d.nk = nk; d.ny = ny;
        d.t0 = 0.0;
        auto opt = [&](const char* name, double dflt) {
            return f.has_attribute("/amiciOptions", name) ? f.attribute("/amiciOptions", name).at(0) : dflt;
        };
        d.atol = opt("atol", 1e-16);
        d.rtol = opt("rtol", 1e-8);
        d.maxsteps = (int32_t)opt("maxsteps", 10000);
        d.stldet = (int32_t)opt("stldet", 1);
        // Options that select an ALGORITHM (deps/AMICI/src/hdf5.cpp readSolverSettingsFromHDF5; enums in
        // deps/AMICI/include/amici/defines.h).  What this path implements is accepted, what would change the results
        // beyond the path's tolerance is refused with a message, never silently replaced:
        //   lmm   2 = BDF (1 = Adams refused)             iter  2 = Newton (1 = functional refused)
        //   ism   1 simultaneous: honoured (the full-feature instantiation of the model's kernels); 2 staggered;
        //         3 staggered1: runs the staggered corrector --
        //         the reference's own variants agree to ~1e-9, far inside the path's tolerance
        //   sensi_meth / sensi_meth_preeq  1 forward, 2 adjoint (parPE's generator writes adjoint,
        //         python/parpe/hdf5_pe_input.py:869-871): accepted -- the gradient is the same quantity and is always
        //         computed by forward sensitivities here; sensi (order) is decided per evaluation by gradient.empty()
        //   linsol  1 dense / 3 LAPACKDense -> dense; 9 KLU / 10 SuperLUMT -> sparse where the symbolic analysis allows
        //         (automatic); band / diagonal / iterative solvers (2, 4..8) refused
        //   ss_tol_factor, ss_tol_sensi_factor (100), ss_atol / ss_rtol / ss_atol_sensi / ss_rtol_sensi (unset),
        //   check_sensi_steadystate_conv (1), newton_step_steadystate_conv (0), newton_damping_factor_mode (1 = on),
        //   max_nonlin_iters (3), max_conv_fails (10), max_step_size (0), rtol_fsa / atol_fsa (unset): only AMICI's
        //         defaults are built; a file that sets another value is refused when it matters (the steady-state
        //         ones only with pre-/post-equilibration in the file)
        //   interpType, nmaxevent, ordering, maxtime, maxstepsB, atolB / rtolB, quad_*: no role on this path
        auto opt_int = [&](const char* name, int dflt) { return (int)std::llround(opt(name, (double)dflt)); };
        auto refuse = [&](const char* name, const std::string& why) {
            throw H5Error(std::string("/amiciOptions ") + name + " = " + std::to_string(opt(name, 0.0)) + ": " + why);
        };
        if (opt_int("lmm", 2) != 2) refuse("lmm", "only BDF (2) is implemented");
        if (opt_int("iter", 2) != 2) refuse("iter", "only Newton iteration (2) is implemented");
        {   // (absent = staggered, the north-star configuration -- AMICI's own default would be simultaneous)
            const int ism = opt_int("ism", 2);
            if (ism < 1 || ism > 3) refuse("ism", "not an InternalSensitivityMethod");
            d.ism = ism == 1 ? 1 : 0;  // 3 (staggered1: one sensitivity at a time) runs the staggered corrector
        }
        for (const char* nm : {"sensi_meth", "sensi_meth_preeq"}) {
            const int v = opt_int(nm, 1);
            if (v < 0 || v > 2) refuse(nm, "not a SensitivityMethod");
        }
        switch (opt_int("linsol", 9)) {
        case 1: case 3: d.linsol = 1; break;   // LinearSolver::dense, LAPACKDense
        case 9: case 10: d.linsol = 0; break;  // KLU, SuperLUMT: sparse where the pattern's fill allows (automatic)
        default: refuse("linsol", "only the direct dense (1, 3) and sparse (9, 10) linear solvers have a counterpart");
        }
        if (opt_int("max_nonlin_iters", 3) != 3) refuse("max_nonlin_iters", "only CVODES' default 3 is implemented");
        if (opt_int("max_conv_fails", 10) != 10) refuse("max_conv_fails", "only CVODES' default 10 is implemented");
        if (opt("max_step_size", 0.0) != 0.0) refuse("max_step_size", "a maximum step size is not implemented");
        for (const char* nm : {"rtol_fsa", "atol_fsa"})
            if (f.has_attribute("/amiciOptions", nm) && opt(nm, 0.0) != (nm[0] == 'r' ? d.rtol : d.atol))
                refuse(nm, "sensitivity tolerances other than rtol / atol are not implemented");
        // Steady states (/amiciOptions steadyStateComputationMode / steadyStateSensitivityMode: 0 newtonOnly,
        // 1 integrationOnly -- AMICI's default for both --, 2 integrateIfNewtonFails; newton_maxsteps;
        // deps/AMICI/src/hdf5.cpp:1433-1448).  SteadystateProblem::findSteadyState (steadystateproblem.cpp:120-166)
        // with forward sensitivities: Newton's method is OFF with computation mode 1, newton_maxsteps = 0 or --
        // for evaluations with gradient -- sensitivity mode 1; the simulation is OFF with computation mode 0.
        //   (1, 1|2), (2, 1|2) with newton_maxsteps = 0   steady state by simulation, sensitivities integrated along
        //   (1, 0),   (2, 0)   with newton_maxsteps = 0   ... simulated without them, solved for at the steady state
        //   (2, 2) / (2, 0) with newton_maxsteps > 0      Newton first, simulation if it fails, Newton once more
        //   (0, 0|2) with newton_maxsteps > 0             Newton's method only
        // Refused (only when the file has pre-/post-equilibration; never approximated): (2, 1) with newton_maxsteps > 0
        // (evaluations with gradient would simulate, those without would try Newton first), (0, 1) (the reference
        // throws, steadystateproblem.cpp:73-84) and computation mode 0 without Newton steps (nothing would run).
        {
            const int comp = opt_int("steadyStateComputationMode", 1), sens = opt_int("steadyStateSensitivityMode", 1);
            const int nmax = std::max(0, opt_int("newton_maxsteps", 0));
            bool steady = P->any_preeq;
            for (double tv : P->t) steady = steady || std::isinf(tv);
            d.newton_maxsteps = 0;
            d.newton_only = 0;
            d.steady_sens_newton = 0;
            bool ok = comp >= 0 && comp <= 2 && sens >= 0 && sens <= 2;
            if (ok && comp == 0) {
                ok = nmax > 0 && sens != 1;
                if (ok) { d.newton_maxsteps = nmax; d.newton_only = 1; }
            } else if (ok) {
                const int n = comp == 2 ? nmax : 0;  // (computation mode 1 ignores newton_maxsteps)
                ok = !(n > 0 && sens == 1);
                if (ok) { d.newton_maxsteps = n; d.steady_sens_newton = sens == 0 ? 1 : 0; }
            }
            if (!steady) d.newton_maxsteps = d.newton_only = d.steady_sens_newton = 0;  // (no role without steady states)
            else if (!ok)
                throw H5Error("/amiciOptions steadyStateComputationMode = " + std::to_string(comp) +
                              ", steadyStateSensitivityMode = " + std::to_string(sens) + ", newton_maxsteps = " +
                              std::to_string(nmax) + " with pre-/post-equilibration: this combination is not supported "
                              "(see the table in csrc/h5input.cu)");
            if (steady) {
                if (opt("ss_tol_factor", 100.0) != 100.0) refuse("ss_tol_factor", "only AMICI's default 100 is implemented");
                if (opt("ss_tol_sensi_factor", 100.0) != 100.0) refuse("ss_tol_sensi_factor", "only AMICI's default 100 is implemented");
                for (const char* nm : {"ss_atol", "ss_rtol", "ss_atol_sensi", "ss_rtol_sensi"}) {
                    // (unset = factor x the integration tolerance; a file may spell that default out)
                    const double dflt = 100.0 * (std::strstr(nm, "atol") ? d.atol : d.rtol);
                    if (f.has_attribute("/amiciOptions", nm) && opt(nm, dflt) != dflt)
                        refuse(nm, "steady-state tolerances other than 100 x atol / rtol are not implemented");
                }
                if (opt_int("check_sensi_steadystate_conv", 1) != 1)
                    refuse("check_sensi_steadystate_conv", "only the default (sensitivities are part of the convergence check) is implemented");
                if (opt_int("newton_step_steadystate_conv", 0) != 0)
                    refuse("newton_step_steadystate_conv", "the Newton-step convergence check is not implemented");
                if (d.newton_maxsteps > 0 && opt_int("newton_damping_factor_mode", 1) != 1)
                    refuse("newton_damping_factor_mode", "only the default damped Newton iteration is implemented");
            }
        }
        d.sensi_errcon = 1;
        d.num_trials = 6;
        d.relaxation_factor = 100.0;
        d.t_offsets = P->t_offsets.data();
        d.t = P->t.data(); d.y = P->y.data(); d.sigma = P->sigma.data(); d.k = P->k.data();
        d.mapping = P->mapping.data(); d.pscale_sim = P->pscale_sim.data(); d.pscale_opt = P->pscale_opt.data();
        d.overrides = P->overrides.empty() ? nullptr : P->overrides.data();
        // ---- hierarchical optimisation tables (optional; empty datasets may be omitted: doc/hdf5.md:118-126)
        auto read_idx = [&](const char* p, std::vector<int32_t>& v) {
            if (f.exists(p)) v = to_i32(f.read_int(p));
        };
        auto read_map = [&](const char* p, std::vector<int32_t>& v) {
            if (!f.exists(p)) return;
            const auto mi = f.info(p);
            if (mi.count() == 0) return;
            if (mi.rank != 2 || mi.dims[1] != 3) throw H5Error(std::string(p) + " must be [n x 3]");
            v = to_i32(f.read_int(p));
        };
        read_idx("/scalingParameterIndices", P->scaling_idx);
        read_idx("/offsetParameterIndices", P->offset_idx);
        read_idx("/sigmaParameterIndices", P->sigma_idx);
        read_map("/scalingParametersMapToObservables", P->scaling_map);
        read_map("/offsetParametersMapToObservables", P->offset_map);
        read_map("/sigmaParametersMapToObservables", P->sigma_map);
        P->has_hier = !(P->scaling_idx.empty() && P->offset_idx.empty() && P->sigma_idx.empty());
        pb200_hier_desc& hd = P->hier;
        hd.n_scaling = (int32_t)P->scaling_idx.size(); hd.n_offset = (int32_t)P->offset_idx.size(); hd.n_sigma = (int32_t)P->sigma_idx.size();
        hd.scaling_idx = P->scaling_idx.data(); hd.offset_idx = P->offset_idx.data(); hd.sigma_idx = P->sigma_idx.data();
        hd.n_scaling_rows = (int32_t)P->scaling_map.size() / 3; hd.n_offset_rows = (int32_t)P->offset_map.size() / 3;
        hd.n_sigma_rows = (int32_t)P->sigma_map.size() / 3;
        hd.scaling_map = P->scaling_map.data(); hd.offset_map = P->offset_map.data(); hd.sigma_map = P->sigma_map.data();
        *out = P.release();
        return 0;
    });
}

const pb200_problem_desc* pb200_h5_problem_desc(pb200_h5_problem* p) { return p ? &p->desc : nullptr; }
int pb200_h5_problem_parameters(pb200_h5_problem* p, const double** lower, const double** upper, const double** nominal,
                                const char** names) {
    if (!p) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    if (lower) *lower = p->lower.empty() ? nullptr : p->lower.data();
    if (upper) *upper = p->upper.empty() ? nullptr : p->upper.data();
    if (nominal) *nominal = p->nominal.empty() ? nullptr : p->nominal.data();
    if (names) *names = p->names.empty() ? nullptr : p->names.c_str();
    return p->desc.n_opt;
}
int pb200_h5_read_strings(pb200_h5* h, const char* path, char* buf, int capacity) {
    if (!h || !path) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    int n = 0;
    const int rc = guarded([&] {
        std::string all;
        for (auto& s : h->f->read_strings(path)) {
            all += s;
            all += '\n';
            ++n;
        }
        if (buf) {
            if ((int)all.size() + 1 > capacity) throw H5Error(std::string(path) + ": output buffer too small");
            std::memcpy(buf, all.c_str(), all.size() + 1);
        }
        return 0;
    });
    return rc < 0 ? rc : n;
}
int pb200_h5_problem_preeq(pb200_h5_problem* p, const int32_t** has_preeq, const double** k_preeq,
                           const int64_t** reinit_offsets, const int32_t** reinit_idx) {
    if (!p || !p->any_preeq) return 0;
    if (has_preeq) *has_preeq = p->has_preeq.data();
    if (k_preeq) *k_preeq = p->k_preeq.data();
    if (reinit_offsets) *reinit_offsets = p->reinit_off.data();
    if (reinit_idx) *reinit_idx = p->reinit_idx.data();
    return 1;
}
const pb200_hier_desc* pb200_h5_problem_hier(pb200_h5_problem* p) { return (p && p->has_hier) ? &p->hier : nullptr; }
void pb200_h5_problem_free(pb200_h5_problem* p) { delete p; }

int pb200_create_from_hdf5(const char* path, const char* model_name, int device, pb200_handle** out) {
    if (!out) return pb200::fail_message(PB200_ERR_ARG, "null argument");
    *out = nullptr;
    pb200_h5_problem* P = nullptr;
    int rc = pb200_h5_load_problem(path, model_name, &P);
    if (rc != 0) return rc;
    rc = pb200_create(&P->desc, device, out);
    if (rc == 0 && P->has_hier) rc = pb200_hier_attach(*out, &P->hier);
    if (rc == 0 && P->any_preeq)
        rc = pb200_set_preequilibration(*out, P->has_preeq.data(), P->k_preeq.data(), P->reinit_off.data(), P->reinit_idx.data());
    if (rc != 0 && *out) {
        pb200_destroy(*out);
        *out = nullptr;
    }
    pb200_h5_problem_free(P);  // (the tables were copied to the device)
    return rc;
}

}  // extern "C"
