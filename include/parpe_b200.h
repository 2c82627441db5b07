/*
 * parpe_b200.h -- C ABI of the B200-native objective/gradient hot path of parPE.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types,
 * no exceptions, caller-allocated outputs.  The reference has no C plugin ABI for
 * this path -- it sits behind the C++ classes parpe::SummedGradientFunction<int> /
 * parpe::GradientFunction (include/parpecommon/functions.h:26-118) implemented by
 * parpe::AmiciSummedGradientFunction (include/parpeamici/multiConditionProblem.h:
 * 95-237).  Each entry point below names the reference interface it replaces;
 * the header-only C++ adapter that forwards to these functions with the
 * reference's own signatures is parpe_b200/include/parpe_b200/
 * b200SummedGradientFunction.h, and INTEGRATION.md shows the binding a parPE
 * maintainer would add.
 *
 * All floating point is IEEE double, all indices 32-bit int (as in the reference)
 * except byte/element offsets, which are int64.
 *
 * Thread-compatibility: one handle must not be evaluated from two threads at
 * once (same contract as the reference, whose evaluate() mutates the solver
 * through const: src/parpeamici/multiConditionProblem.cpp:826-835).  Different
 * handles may be used concurrently.
 */
#ifndef PARPE_B200_H
#define PARPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 2

/* parpe::FunctionEvaluationStatus (include/parpecommon/functions.h:16-20) */
#define PB200_EVAL_SUCCESS 0
#define PB200_EVAL_FAILURE 1

/* library-level error codes (negative; never returned by evaluate for a
 * simulation failure -- those are reported the reference's way: fval=+inf and
 * PB200_EVAL_FAILURE, src/parpeamici/multiConditionProblem.cpp:669-672) */
#define PB200_ERR_ARG (-1)
#define PB200_ERR_MODEL (-2)
#define PB200_ERR_CUDA (-3)
#define PB200_ERR_NO_DEVICE (-4)

/* amici::ParameterScaling (deps/AMICI/include/amici/defines.h:118-122) */
#define PB200_SCALE_NONE 0
#define PB200_SCALE_LN 1
#define PB200_SCALE_LOG10 2

/* per-condition simulation status: amici return codes
 * (deps/AMICI/include/amici/defines.h:60-91); 0 = AMICI_SUCCESS */
#define PB200_SIM_SUCCESS 0
#define PB200_SIM_TOO_MUCH_WORK (-1)
#define PB200_SIM_TOO_MUCH_ACC (-2)
#define PB200_SIM_ERR_FAILURE (-3)
#define PB200_SIM_CONV_FAILURE (-4)
#define PB200_SIM_LSETUP_FAIL (-6)
#define PB200_SIM_RHSFUNC_FAIL (-8)
#define PB200_SIM_FIRST_RHSFUNC_ERR (-9)
#define PB200_SIM_REPTD_RHSFUNC_ERR (-10)
#define PB200_SIM_UNREC_RHSFUNC_ERR (-11)
#define PB200_SIM_ILL_INPUT (-22)
#define PB200_SIM_TOO_CLOSE (-27)
#define PB200_SIM_SRHSFUNC_FAIL (-41)
#define PB200_SIM_REPTD_SRHSFUNC_ERR (-43)
#define PB200_SIM_UNREC_SRHSFUNC_ERR (-44)
#define PB200_SIM_ERROR (-99)      /* AMICI_ERROR: e.g. sigma <= 0 or non-finite */

/* per-condition integrator counters returned by pb200_get_condition_results
 * (same counters CVODES exposes; deps/AMICI/src/solver.cpp:342-368) */
enum {
    PB200_STAT_NST = 0,   /* accepted steps                     */
    PB200_STAT_NFE,       /* RHS evaluations                    */
    PB200_STAT_NSETUPS,   /* LU factorisations                  */
    PB200_STAT_NJE,       /* Jacobian evaluations               */
    PB200_STAT_NNI,       /* state Newton iterations            */
    PB200_STAT_NNIS,      /* sensitivity Newton iterations      */
    PB200_STAT_NFSE,      /* sensitivity RHS evaluations        */
    PB200_STAT_NETF,      /* state error-test failures          */
    PB200_STAT_NETFS,     /* sensitivity error-test failures    */
    PB200_STAT_NCFN,      /* state corrector failures           */
    PB200_STAT_NCFNS,     /* sensitivity corrector failures     */
    PB200_STAT_NOR,       /* order reductions by STALD          */
    PB200_STAT_TRIALS,    /* tolerance-relaxation trials used   */
    PB200_STAT_QSUM,      /* sum over accepted steps of order q */
    /* large-model kernel only (0 otherwise): device cycles of the phases of the solve, two 32-bit
     * fields of kilo-cycles each -- diagnostics behind DESIGN.md 3.4's phase split */
    PB200_STAT_CYC_INVERSION, /* (panel factorisation << 32) | rank-NB update */
    PB200_STAT_CYC_NEWTON,    /* (solve GEMMs << 32) | right-hand sides       */
    PB200_STAT_NFALLBACK,     /* setups whose sparse factorisation met a weak diagonal pivot and were
                                 redone by the dense path with partial pivoting */
    PB200_STAT_CYC_SPARSE,    /* (sparse numeric factorisation << 32) | sparse triangular solves */
    PB200_STAT_CYC_STEP,      /* (history sweeps << 32) | whole state phase of the step attempts */
    PB200_STAT_CYC_SENS,      /* (whole sensitivity phase << 32) | whole simulation */
    PB200_NSTATS = 20
};

/*
 * Everything evaluate() needs that the reference reads from its data provider
 * (parpe::MultiConditionDataProvider, include/parpeamici/
 * multiConditionDataProvider.h:22-111; HDF5 layout doc/hdf5.md:1-172).  All
 * tables are copied to the device at pb200_create; the caller may free them
 * afterwards.
 */
typedef struct pb200_problem_desc {
    int32_t abi_version;       /* PB200_ABI_VERSION */
    const char* model_name;    /* generated model, see pb200_model_info */
    int32_t n_conditions;      /* C: number of simulation conditions */
    int32_t n_opt;             /* number of optimisation parameters */
    int32_t np, nk, ny;        /* must equal the model's dimensions (checked) */
    /* measurements, /measurements/{t,y,ysigma}/<c> of doc/hdf5.md:60-83 */
    const int64_t* t_offsets;  /* [C+1] first output time of condition c in t[] */
    const double* t;           /* [t_offsets[C]] output times (may repeat) */
    const double* y;           /* [t_offsets[C]*ny] row-major obs + it*ny; NaN = no measurement */
    const double* sigma;       /* same layout; NaN = use the model's sigma */
    /* /fixedParameters/k column of the simulation condition, doc/hdf5.md:36-58 */
    const double* k;           /* [C*nk] row-major per condition */
    /* /parameters/optimizationSimulationMapping (transposed: per condition
     * contiguous), -1 = not mapped (multiConditionDataProvider.cpp:84-102) */
    const int32_t* mapping;    /* [C*np] */
    const int32_t* pscale_sim; /* [C*np]  /parameters/pscaleSimulation */
    const int32_t* pscale_opt; /* [n_opt] /parameters/pscaleOptimization */
    const double* overrides;   /* [C*np] /parameters/parameterOverrides, or NULL */
    /* /amiciOptions attributes (deps/AMICI/src/hdf5.cpp) */
    double t0;
    double atol, rtol;         /* also used for sensitivities (rtol_fsa/atol_fsa unset) */
    int32_t maxsteps;          /* total internal steps per simulation (AMICI default 10000) */
    int32_t stldet;            /* BDF stability-limit detection, AMICI default 1 */
    int32_t sensi_errcon;      /* sensitivities in the error test; AMICI forces 1 (solver.cpp:646) */
    /* runAndLogSimulation retry policy (multiConditionProblem.cpp:352-365) */
    int32_t num_trials;        /* PARPE_NUM_SIMULATION_TRIALS, default 6 */
    double relaxation_factor;  /* PARPE_INTEGRATION_TOLERANCE_RELAXATION_FACTOR, default 100 */
    /* kernel mapping hint: lanes cooperating on one condition (0 = automatic) */
    int32_t lanes_per_condition;
    /* steady states (pre-/post-equilibration): > 0 = Newton's method first, with this many steps at most, the
     * simulation to steady state if it fails, Newton once more if that fails too -- what the reference does with
     * Solver::setNewtonMaxSteps(n) AND the model's steadyStateComputationMode = steadyStateSensitivityMode =
     * integrateIfNewtonFails (deps/AMICI/src/steadystateproblem.cpp:120-166,628-715; newton_solver.cpp:43-125).
     * 0 = steady state by simulation only: AMICI's defaults (both modes integrationOnly, with which the reference
     * ignores newton_maxsteps).  Small-model kernels only. */
    int32_t newton_maxsteps;
    /* /amiciOptions linsol (amici::LinearSolver, deps/AMICI/include/amici/defines.h; read by hdf5.cpp
     * readSolverSettingsFromHDF5) as far as this path distinguishes linear solvers: 0 = automatic, 1 = dense
     * (LinearSolver::dense / LAPACKDense: the dense inverse), 2 = sparse (LinearSolver::KLU / SuperLUMT: the
     * fixed-pattern sparse LU with a dense Schur block).  The initial value of PB200_OPT_LINSOL; only the
     * large-model kernels have a choice (the small-model kernels always invert densely). */
    int32_t linsol;
    /* 1 = steadyStateComputationMode newtonOnly (steadystateproblem.cpp:143-146: turnOffSimulation): Newton's
     * method with newton_maxsteps steps at most and NO simulation to steady state if it fails -- the simulation
     * then ends with AMICI_ERROR; sensitivities of the steady state by the linear solve (steadyStateSensitivityMode
     * newtonOnly or integrateIfNewtonFails).  Needs newton_maxsteps > 0 (with 0 the reference runs neither method and
     * fails every simulation: refused at pb200_create).  Small-model kernels only. */
    int32_t newton_only;
    /* 1 = steadyStateSensitivityMode newtonOnly together with a SIMULATION to steady state (computation mode
     * integrationOnly, or integrateIfNewtonFails with newton_maxsteps > 0 for `Newton first`): the simulation runs
     * without the sensitivity columns (Solver::switchForwardSensisOff, steadystateproblem.cpp:728-736) and the
     * sensitivities of the steady state come from the linear solve sx = -J^-1 dxdotdp (newton_solver.cpp:52-84;
     * a singular Jacobian is AMICI_ERROR).  0 = sensitivities integrated along (integrationOnly /
     * integrateIfNewtonFails).  Small-model kernels only. */
    int32_t steady_sens_newton;
    /* /amiciOptions ism (amici::InternalSensitivityMethod, CVodeSensInit1's ism): 0 or 2 = staggered -- the
     * path's corrector (north star: forward, staggered); 1 = simultaneous (CV_SIMULTANEOUS: state and
     * sensitivities in one Newton iteration and one error test, cvodes_nls_sim.c) -- run by the full-feature
     * (steady-state) instantiation of the model's kernels, small- and large-model alike.  AMICI's own default is
     * simultaneous; the variants agree to ~1e-9 relative. */
    int32_t ism;
    int32_t reserved[2];
} pb200_problem_desc;

typedef struct pb200_handle pb200_handle;

/* library / model registry ------------------------------------------------ */
int pb200_abi_version(void);
/* number of CUDA devices visible; <= 0 if none (no CPU fallback exists) */
int pb200_device_count(void);
int pb200_model_count(void);
/* name of model i, or NULL */
const char* pb200_model_name(int i);
/* dimensions of a generated model; returns 0 or PB200_ERR_MODEL */
int pb200_model_info(const char* model_name, int* nx, int* np, int* nk, int* ny);
/* last error message of the calling thread ("" if none) */
const char* pb200_last_error(void);

/* measured FP64 DFMA peak of `device` in TFLOP/s (roofline denominator; no reference
 * counterpart) */
int pb200_measure_fp64_peak(int device, double* tflops);

/* lifetime ---------------------------------------------------------------- */
/*
 * Replaces: construction of MultiConditionDataProviderHDF5 +
 * AmiciSummedGradientFunction (multiConditionProblem.cpp:612-635).
 * device: CUDA ordinal.  Returns 0 and *out, or a negative PB200_ERR_*.
 */
int pb200_create(const pb200_problem_desc* desc, int device, pb200_handle** out);
void pb200_destroy(pb200_handle* h);

/*
 * Pre-equilibration.  Replaces: the pre-equilibration branch of
 * MultiConditionDataProviderHDF5::updateFixedSimulationParameters
 * (src/parpeamici/multiConditionDataProvider.cpp:224-250: column 0 of
 * /fixedParameters/simulationConditions != -1 selects the pre-equilibration condition,
 * /fixedParameters/reinitializationIndices the states re-initialised afterwards; doc/hdf5.md:36-58)
 * and what AMICI then does with ExpData::fixedParametersPreequilibration /
 * reinitialization_state_idxs_sim (deps/AMICI/src/amici.cpp:95-108,
 * steadystateproblem.cpp:82-106,711-794): the model is simulated with the pre-equilibration
 * condition's constants from x0 until the weighted RMS norm of xdot (and of the sensitivity
 * right-hand sides) falls below 1 at 100 x the integration tolerances, the main simulation starts
 * from that steady state (and its sensitivities), and the listed states whose initial value depends
 * on the constants are reset to their initial value.  AMICI's defaults are implemented -- steady
 * state by simulation, sensitivities integrated along, convergence check on the sensitivities on --
 * and, with pb200_problem_desc::newton_maxsteps > 0, Newton's method first; newton_only and
 * steady_sens_newton select the model's other steady-state modes (small-model kernels).
 *   has_preeq       [C]     1 = condition c is pre-equilibrated; NULL switches pre-equilibration off
 *   k_preeq         [C*nk]  constants of the pre-equilibration condition (rows of conditions
 *                           without one are ignored)
 *   reinit_offsets  [C+1], reinit_idx: ragged list of state indices per condition; NULL = none
 * Output times t = +inf in the problem table (post-equilibration) need no call: the data point is
 * evaluated at the steady state the simulation is carried on to (rdata.cpp:234-258).
 * Both kernel families (PB200_ERR_ARG for a model whose kernels were built without the steady-state
 * instantiation); Newton's method for the small-model kernels only.
 */
int pb200_set_preequilibration(pb200_handle* h, const int32_t* has_preeq, const double* k_preeq,
                               const int64_t* reinit_offsets, const int32_t* reinit_idx);

/*
 * Shard the data indices of every evaluation over `world` handles (one per GPU /
 * process): this handle then simulates positions [floor(rank*n/world),
 * floor((rank+1)*n/world)) of the index list only.
 * Replaces: LoadBalancerMaster's dynamic assignment of packages to MPI workers
 * (src/parpeloadbalancer/loadBalancerMaster.cpp:104-134) by a static contiguous
 * partition in index order (amiciSimulationRunner.cpp:49-80 chunk order).
 * pb200_evaluate on a partitioned handle returns the GLOBAL fval / gradient /
 * status on every rank (the reference master sums all workers' results before
 * evaluate returns: multiConditionProblem.cpp:738-748,775-824) and therefore
 * needs a collective: pb200_nccl_init (in-library NCCL all-reduce of 2 + n_opt
 * doubles on the device) or pb200_set_allreduce (the caller's host transport,
 * e.g. MPI_Allreduce).  Without one it fails with PB200_ERR_ARG, unless
 * rank-local partial sums were requested with PB200_OPT_PARTIAL_SUMS.
 */
int pb200_set_partition(pb200_handle* h, int rank, int world);

/*
 * In-library collective: one NCCL communicator over the `world` handles of a
 * partitioned problem (one handle per GPU; the handles may live in different
 * processes or in different threads of one process).  Rank 0 obtains a 128-byte
 * id with pb200_nccl_unique_id and distributes it by any means (MPI_Bcast in
 * parPE; a thread-shared buffer in a single process); every rank then calls
 * pb200_nccl_init, which also sets the partition (rank, world).  From then on
 * pb200_evaluate, pb200_evaluate_device and pb200_hier_evaluate combine the
 * ranks' (nllh, #errors, gradient[n_opt]) with ONE ncclAllReduce on the
 * handle's stream before the +inf / failure rule is applied.  libnccl.so.2 is
 * resolved at run time (dlopen); PB200_ERR_CUDA if it is absent.
 */
int pb200_nccl_unique_id(char* id128);
int pb200_nccl_init(pb200_handle* h, const char* id128, int rank, int world);

/*
 * Tuning / experiment switches.  The library never reads the environment; these
 * are set per handle by the caller (tools/ scripts set them for experiments).
 */
enum {
    PB200_OPT_REORDER = 1,        /* 1 (default): conditions are handed out costliest first, by the
                                     step counts of an earlier evaluation of the same index list;
                                     0: index order.  Results do not depend on it. */
    PB200_OPT_ORDER_REFRESH = 2,  /* refresh that order every this many evaluations (default 16) */
    PB200_OPT_KERNEL_VARIANT = 3, /* small-model kernel CTA shape: 2 (default) one CTA per SM, 0 two-warp CTAs */
    PB200_OPT_CTAS_PER_SM = 4,    /* cap on resident CTAs per SM (0 = occupancy limit) */
    PB200_OPT_SYNC_PERIOD = 5,    /* CTA barrier every this many step attempts (default 1) */
    PB200_OPT_STAGE_SYNC = 6,     /* CTA barrier between the stages of a step attempt (default 0) */
    PB200_OPT_PARTIAL_SUMS = 7,   /* 1: pb200_evaluate on a partitioned handle without a collective
                                     returns rank-local partial sums instead of failing (tests) */
    PB200_OPT_FUSE_OFF = 9,       /* large-model path, experiments: bit 0 switches the fused history pass off, bit 1
                                     the fused Newton iteration of the sparse solver, bit 2 the lazy zeroing of the
                                     correction (default 0: all on; results agree to rounding) */
    PB200_OPT_LINSOL = 8          /* large-model path: 0 automatic (by fill of the symbolic factorisation),
                                     1 dense inverse (amici LinearSolver::dense), 2 sparse fixed-pattern LU
                                     with a dense Schur block (LinearSolver::KLU's role) */
};
int pb200_set_option(pb200_handle* h, int key, int value);
/* [lo,hi) of an n-long index list owned by (rank, world): the assignment rule */
void pb200_partition_bounds(int n, int rank, int world, int* lo, int* hi);

/* the hot path ------------------------------------------------------------ */
/*
 * Replaces: AmiciSummedGradientFunction::evaluate(parameters, datasets, fval,
 * gradient, logger, cpuTime)  (multiConditionProblem.cpp:647-675).
 *   theta[n_opt]   optimisation parameters (optimiser scale), host memory
 *   idx[n_idx]     data (condition) indices, host memory; duplicates allowed
 *   grad           NULL => cost only (reference: gradient.empty())
 *   fval, grad     fully overwritten; grad has n_opt entries
 *   seconds        nullable; device time of the evaluation
 * Returns PB200_EVAL_SUCCESS / PB200_EVAL_FAILURE (then *fval = +inf), or a
 * negative PB200_ERR_* for API misuse / CUDA errors.
 */
int pb200_evaluate(pb200_handle* h, const double* theta, int n_opt, const int* idx,
                   int n_idx, double* fval, double* grad, double* seconds);

/*
 * Same computation, enqueued on `stream` (a cudaStream_t; NULL = default stream)
 * without host synchronisation, results left on the device:
 * d_out[0] = partial nllh, d_out[1] = number of failed conditions,
 * d_out[2..2+n_opt) = partial gradient (zeros if !want_grad).  theta and idx are
 * host pointers (the optimiser lives on the host; they are staged through pinned
 * memory).  With pb200_nccl_init the all-reduce over the ranks is enqueued on the
 * same stream and d_out holds the GLOBAL sums; otherwise the partial sums of this
 * rank.  Used for device-timed throughput.  The +inf / failure mapping is the
 * caller's job.
 */
int pb200_evaluate_device(pb200_handle* h, const double* theta, int n_opt, const int* idx,
                          int n_idx, int want_grad, double* d_out, void* stream);

/*
 * Large-model path: what the one-time symbolic analysis of the Newton matrix's pattern produced
 * (the klu_analyze role, sunlinsol_klu.c:259-330).  out[0..n): solver the automatic choice takes
 * (1 dense inverse, 2 sparse LU + dense Schur block), nx, sparse pivots n1, dense block m, nnz(J),
 * nnz(L), nnz(U) of the sparse part, multiply-adds of one numeric factorisation of the sparse part
 * incl. the Schur assembly, levels of the forward / backward substitution, flops of one numeric
 * factorisation (without the dense block), flops of the two triangular solves for one right-hand
 * side, bytes of the factors (sparse values + dense block) written per setup / read per solve.
 * Returns the number of entries written (0 for the small-model kernels).
 */
int pb200_linsol_info(pb200_handle* h, int64_t* out, int n);

/* host<->device bytes moved by the most recent pb200_evaluate call on this handle (parameter table,
 * index list if it changed, results, processing-order refresh): what bench.py reports as e2e bytes */
int pb200_last_transfer_bytes(pb200_handle* h, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* launch geometry and device time of the most recent simulate+reduce launch pair */
int pb200_last_launch(pb200_handle* h, int* lanes_per_condition, int* grid, int* block,
                      double* seconds);

/*
 * Per-condition results of the most recent evaluation on this handle, in
 * index-list order restricted to this handle's partition (what the reference
 * ships back per condition as AmiciResultPackageSimple,
 * include/parpeamici/amiciSimulationRunner.h:157-182).  Any pointer may be NULL.
 *   llh[n]  sllh[n*np] (plist-compacted, simulation scale)  status[n]
 *   stats[n*PB200_NSTATS]
 * Returns the number of conditions n, or negative on error.
 */
int pb200_get_condition_results(pb200_handle* h, double* llh, double* sllh,
                                int* status, int64_t* stats);

/*
 * Replaces: AmiciSummedGradientFunction::getModelOutputsAndSigmas
 * (multiConditionProblem.cpp:506-563): simulate ALL conditions without
 * sensitivities; y_out / sigmay_out use the layout of desc->y.
 * Returns PB200_EVAL_SUCCESS / PB200_EVAL_FAILURE.
 * On a partitioned handle only the conditions of this rank's share are
 * simulated: their outputs are filled, all others are NaN (the status is the
 * global one when a collective is set).  pb200_hier_evaluate works on these
 * per-rank outputs and reduces what it needs.
 */
int pb200_get_model_outputs(pb200_handle* h, const double* theta, int n_opt,
                            double* y_out, double* sigmay_out);

/* parPE's HDF5 input file ------------------------------------------------ */
/*
 * Replaces: the HDF5 reads of MultiConditionDataProviderHDF5
 * (src/parpeamici/multiConditionDataProvider.cpp:42-60,84-102,194-206,224-250,
 * 302-321) and AnalyticalParameterHdf5Reader, for the input layout of
 * doc/hdf5.md:1-172 -- /parameters/{optimizationSimulationMapping,
 * parameterOverrides,pscaleSimulation,pscaleOptimization}, /fixedParameters/{k,
 * simulationConditions}, /measurements/{t,y,ysigma}/<i>, /amiciOptions
 * attributes, /{scaling,offset,sigma}Parameter{Indices,sMapToObservables}.
 * The file is parsed by a built-in reader of the HDF5 subset such files use
 * (old-style groups, contiguous / compact / chunked datasets, gzip + shuffle
 * filters, variable-length integer sequences in global heap collections;
 * parpe_b200/csrc/h5min.h); no HDF5 library is needed.  Pre-equilibration
 * conditions (column 0 of /fixedParameters/simulationConditions) and
 * /fixedParameters/reinitializationIndices are read (multiConditionDataProvider.cpp:224-250,
 * 497-524) and applied by pb200_create_from_hdf5 through pb200_set_preequilibration.
 */
typedef struct pb200_h5 pb200_h5;
int pb200_h5_open(const char* path, pb200_h5** out);
void pb200_h5_close(pb200_h5* f);
int pb200_h5_exists(pb200_h5* f, const char* path);
/* rank, dims[<= 8], class (0 integer, 1 floating point, 3 string) and element size of a dataset */
int pb200_h5_dataset_info(pb200_h5* f, const char* path, int* rank, int64_t* dims, int* type_class, int* type_size);
/* whole dataset as doubles, row-major; returns the element count (out == NULL: count only) or < 0 */
int64_t pb200_h5_read_double(pb200_h5* f, const char* path, double* out, int64_t capacity);
/* the elements' bytes as stored (e.g. fixed-length strings); returns the byte count (out == NULL: count only) or < 0 */
int64_t pb200_h5_read_raw(pb200_h5* f, const char* path, char* out, int64_t capacity);
/* dataset of strings (fixed-length, or variable-length as h5py writes str arrays), newline separated; returns
 * their number or < 0 */
int pb200_h5_read_strings(pb200_h5* f, const char* path, char* buf, int capacity);
/* numeric attribute of a group / dataset as doubles; returns the element count or < 0 */
int pb200_h5_attribute_double(pb200_h5* f, const char* object, const char* name, double* out, int capacity);
/* newline-separated names of a group's children; returns their number or < 0 */
int pb200_h5_list(pb200_h5* f, const char* group, char* buf, int capacity);

/* the assembled tables of an input file (owned by the returned object; the desc points into it) */
typedef struct pb200_h5_problem pb200_h5_problem;
int pb200_h5_load_problem(const char* path, const char* model_name, pb200_h5_problem** out);
const pb200_problem_desc* pb200_h5_problem_desc(pb200_h5_problem* p);
const struct pb200_hier_desc* pb200_h5_problem_hier(pb200_h5_problem* p); /* NULL: no hierarchical tables */
/* /parameters/{lowerBound, upperBound, nominalValues, parameterNames} of the file (what
 * MultiConditionProblem::fillParametersMin / fillParametersMax and
 * MultiConditionDataProviderHDF5::getProblemParameterIds read, multiConditionDataProvider.cpp:49-58,426-432):
 * pointers into the object, NULL for a dataset the file does not have; names newline separated.  Returns n_opt. */
int pb200_h5_problem_parameters(pb200_h5_problem* p, const double** lower, const double** upper,
                                const double** nominal, const char** names);
/* the arguments of pb200_set_preequilibration as read from the file; returns 1 if any condition is
 * pre-equilibrated (pointers into the object), else 0 */
int pb200_h5_problem_preeq(pb200_h5_problem* p, const int32_t** has_preeq, const double** k_preeq,
                           const int64_t** reinit_offsets, const int32_t** reinit_idx);
void pb200_h5_problem_free(pb200_h5_problem* p);
/* pb200_h5_load_problem + pb200_create (+ pb200_hier_attach / pb200_set_preequilibration when the file has the tables) */
int pb200_create_from_hdf5(const char* path, const char* model_name, int device, pb200_handle** out);

/* per-simulation result logging ----------------------------------------- */
/*
 * SURVEY 8(f) rank 4.  Replaces: logSimulation / saveSimulation
 * (src/parpeamici/multiConditionProblem.cpp:254-325), which appends one record per simulation job to
 * extendable datasets of the result file's group /multistarts/<n> (doc/hdf5.md:190-226):
 *   jobId [int 1 x n], simulationLabel [string n] ("o<opt>i<iter>c<condition>"),
 *   simulationLogLikelihood [1 x n], simulationLogLikelihoodGradient [np x n] (sllh in simulation
 *   scale, plist-compacted, NaN padded -- or all NaN when no gradient was requested),
 *   simulationParameters [np x n] (the condition's model parameters in simulation scale),
 *   simulationWallTimeInSec [1 x n], simulationStatus [int 1 x n] (AMICI status),
 * plus, beyond the reference's datasets, the tolerance-relaxation retry statistics:
 *   simulationTrials [int 1 x n], simulationIntegratorStatistics [int64 14 x n] (PB200_STAT_NST ..
 *   PB200_STAT_QSUM).
 * The records are collected in host memory and written ONCE by pb200_result_log_write as a classic-format
 * HDF5 file (parpe_b200/csrc/h5wmin.h; no HDF5 library needed; contiguous datasets of the same names,
 * types and shapes instead of extendable chunked ones).  Conditions run side by side on the GPU, so
 * simulationWallTimeInSec is the evaluation's device time shared equally among its simulations.
 */
typedef struct pb200_result_log pb200_result_log;
int pb200_result_log_create(pb200_result_log** out);
void pb200_result_log_destroy(pb200_result_log* log);
/* the per-condition results of the handle's last pb200_evaluate, as records of `group`; returns their number */
int pb200_result_log_append(pb200_result_log* log, pb200_handle* h, const char* group, int optimization_idx,
                            int iteration_idx);
/* the same from host arrays (sllh / parameters: [n*np] row-major per record; seconds, sllh, parameters, stats
 * ([n*PB200_NSTATS]) may be NULL) */
int pb200_result_log_append_records(pb200_result_log* log, const char* group, int n, int np,
                                    const int32_t* condition_idx, int optimization_idx, int iteration_idx,
                                    const double* llh, const double* sllh, const double* parameters,
                                    const double* seconds, const int32_t* status, const int64_t* stats);
int pb200_result_log_num_records(pb200_result_log* log, const char* group);
int pb200_result_log_write(pb200_result_log* log, const char* path);
/*
 * Replaces: the result datasets of the standalone simulator (src/parpeamici/standaloneSimulator.cpp:57-270;
 * doc/hdf5.md:228-252): simulates all conditions at `theta` and writes <group>/llh [C], <group>/parameters
 * [n_opt], <group>/t/<m> [nt], <group>/yMes/<m> [nt x ny], <group>/ySim/<m> [nt x ny].  Returns the
 * evaluation status (PB200_EVAL_*) or a negative error.
 */
int pb200_write_simulation_results(pb200_handle* h, const double* theta, int n_opt, const char* path,
                                   const char* group);

/* hierarchical optimisation -------------------------------------------- */
/*
 * Analytic inner problem of parpe::HierarchicalOptimizationWrapper
 * (src/parpeamici/hierarchicalOptimization.cpp:142-238,702-930,989-1134).
 * The inner parameters are identified by their optimisation-parameter index;
 * each comes with (condition, observable) pairs it applies to
 * (AnalyticalParameterProvider, hierarchicalOptimizationAnalyticalParameterProvider.h).
 */
typedef struct pb200_hier_desc {
    /* /scalingParameterIndices, /offsetParameterIndices, /sigmaParameterIndices
     * (doc/hdf5.md:118-172): optimisation-parameter indices, ascending */
    int32_t n_scaling, n_offset, n_sigma;
    const int32_t* scaling_idx;     /* [n_scaling] */
    const int32_t* offset_idx;      /* [n_offset] */
    const int32_t* sigma_idx;       /* [n_sigma] */
    /* /{scaling,offset,sigma}ParametersMapToObservables: rows of
     * (inner parameter number, condition index, observable index), row-major
     * int[n_rows*3], as read by AnalyticalParameterHdf5Reader::
     * readParameterConditionObservableMappingFromFile
     * (hierarchicalOptimizationAnalyticalParameterProvider.cpp) */
    int32_t n_scaling_rows, n_offset_rows, n_sigma_rows;
    const int32_t* scaling_map;
    const int32_t* offset_map;
    const int32_t* sigma_map;
    int32_t reserved[8];
} pb200_hier_desc;

int pb200_hier_attach(pb200_handle* h, const pb200_hier_desc* hd);
/* number of outer (numerically optimised) parameters = n_opt - n_inner */
int pb200_hier_num_outer(pb200_handle* h);
/*
 * Collective hook for partitioned (multi-GPU / multi-process) HIERARCHICAL evaluation: with
 * pb200_set_partition(rank, world > 1) every rank simulates its share of the conditions, and the
 * analytic inner parameters need sums over ALL conditions (hierarchicalOptimization.cpp:311-441 run
 * on the master over all results).  The library calls fn(user, buf, n, op) on HOST buffers at the
 * reduction points -- error count of pass 1, scaling/offset sums, sigma sums and maxima, analytic
 * nllh, and (nllh, #errors, gradient) of pass 2 -- and expects the element-wise all-reduce in place
 * (op: PB200_REDUCE_SUM / PB200_REDUCE_MAX); every rank must call pb200_hier_evaluate with the same
 * arguments.  The caller implements fn with its own transport (MPI_Allreduce in parPE, NCCL or gloo
 * through torch.distributed in this repository).  NULL (default) = single rank.
 */
enum { PB200_REDUCE_SUM = 0, PB200_REDUCE_MAX = 1 };
typedef void (*pb200_allreduce_fn)(void* user, double* buf, int n, int op);
int pb200_set_allreduce(pb200_handle* h, pb200_allreduce_fn fn, void* user);

/*
 * Replaces: HierarchicalOptimizationWrapper::evaluate (hierarchicalOptimization.cpp:
 * 142-238).  theta_outer[n_outer] -> fval, grad_outer[n_outer] (NULL => cost only);
 * full_parameters[n_opt] (nullable) receives the spliced vector incl. the optimal
 * inner parameters.
 */
int pb200_hier_evaluate(pb200_handle* h, const double* theta_outer, int n_outer,
                        double* fval, double* grad_outer, double* full_parameters,
                        double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* PARPE_B200_H */
