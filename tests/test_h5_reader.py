"""The built-in HDF5 reader (parpe_b200/csrc/h5min.h behind pb200_h5_*) and the input-file -> tables
assembly (pb200_h5_load_problem: the MultiConditionDataProviderHDF5 role).

Pinned two ways: (1) on a file written by the REAL HDF5 library -- the reference tree's own
deps/AMICI/tests/cpp/testOptions.h5 (read in place; values from deps/AMICI/tests/cpp/testfunctions.h /
the model tests that consume it) -- and (2) on parPE-layout input files written by the committed minimal writer
tools/h5write_min.py (contiguous, chunked and gzip-compressed datasets), round-tripped against the tables
they were written from; on a GPU the handle created from the file gives the same objective as the handle
created from the in-memory tables."""
import os
import sys

import numpy as np
import pytest

from parpe_b200 import capi, synth
from parpe_b200.problem import B200SummedGradientFunction, tables_from_hdf5

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_FILE = "/root/reference/deps/AMICI/tests/cpp/testOptions.h5"


def test_reads_a_file_written_by_the_hdf5_library():
    if not os.path.exists(REF_FILE):
        pytest.skip("/root/reference not present")
    f = capi.H5File(REF_FILE)
    top = f.list("/")
    assert top == ["model_calvetti", "model_dirac", "model_events", "model_jakstat_adjoint", "model_nested_events",
                   "model_neuron", "model_robertson", "model_steadystate"]
    cases = f.list("/model_steadystate")
    assert "sensiforward" in cases and "nosensi" in cases and len(cases) == 14   # (a multi-entry group B-tree)
    g = "/model_steadystate/sensiforward/options"
    assert f.list(g) == ["kappa", "sens_ind", "theta", "ts"]
    # the steady-state example's nominal parameters (log10 of 1, 0.5, 0.4, 2, 0.1) and constants -- the same
    # numbers examples/parpeamici/steadystate/exampleSteadystateScaledTest.h:40-48 uses
    np.testing.assert_allclose(f.read(g + "/theta"), np.log10([1.0, 0.5, 0.4, 2.0, 0.1]), atol=1e-8)
    np.testing.assert_allclose(f.read(g + "/kappa"), [0.1, 0.4, 0.7, 1.0])
    assert np.array_equal(f.read(g + "/sens_ind"), [0, 1, 2, 3, 4])
    assert f.info(g + "/theta") == dict(shape=(5,), type_class=1, type_size=8)
    assert f.info(g + "/sens_ind") == dict(shape=(5,), type_class=0, type_size=4)
    # solver settings as attributes (deps/AMICI/src/hdf5.cpp readSolverSettingsFromHDF5): AMICI's defaults,
    # linsol = 9 is LinearSolver::KLU
    assert f.attribute(g, "atol")[0] == 1e-16 and f.attribute(g, "rtol")[0] == 1e-8
    assert f.attribute(g, "maxsteps")[0] == 10000 and f.attribute(g, "sensi")[0] == 1 and f.attribute(g, "sensi_meth")[0] == 1
    assert f.attribute(g, "linsol")[0] == 9 and f.attribute(g, "stldet")[0] == 1 and f.attribute(g, "newton_maxsteps")[0] == 40
    ts = f.read(g + "/ts")
    assert ts.shape == (51,) and ts[0] == 0.0 and ts[1] == pytest.approx(100.0 / 49.0) and np.all(np.diff(ts) > 0)
    assert not f.exists(g + "/no_such_dataset")
    with pytest.raises(capi.B200LibraryError):
        f.read(g + "/no_such_dataset")


def test_rejects_non_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(capi.B200LibraryError, match="HDF5"):
        capi.H5File(p)


CASES = [("steadystate_scaled", 5, dict(scale="mixed", condition_specific=2, sigma_mode="mixed", missing_fraction=0.2, ragged=True)),
         ("boehm", 4, dict(scale="log10", holes=2)),
         ("net32", 3, dict(scale="log10"))]


@pytest.mark.parametrize("compress", [True, False], ids=["gzip-chunked", "contiguous"])
@pytest.mark.parametrize("model,C,kw", CASES, ids=[c[0] for c in CASES])
def test_input_file_round_trip(tmp_path, model, C, kw, compress):
    """ProblemTables -> parPE input file (tools/h5write_min.py) -> pb200_h5_load_problem: every table identical
    (bit for bit: NaN markers of missing data / model sigmas included), transposed layouts resolved as the
    reference's hyperslab reads do."""
    import h5write_min
    tab = synth.make_problem(model, C, seed=3, **kw)
    path = tmp_path / "input.h5"
    h5write_min.write_parpe_input(str(path), tab, compress=compress)
    f = capi.H5File(path)
    assert f.info("/parameters/optimizationSimulationMapping")["shape"] == (tab.np_, C)
    assert f.info("/fixedParameters/k")["shape"] == (tab.nk, C)
    assert f.list("/measurements/y") == sorted(str(c) for c in range(C))
    got = tables_from_hdf5(path, model)
    assert got.n_conditions == C and got.n_opt == tab.n_opt
    for name in ("t_offsets", "t", "k", "mapping", "pscale_sim", "pscale_opt"):
        assert np.array_equal(getattr(got, name), getattr(tab, name)), name
    for name in ("y", "sigma"):
        a, b = getattr(got, name), getattr(tab, name)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), name
    if tab.overrides is None:
        assert got.overrides is None
    else:
        assert np.array_equal(np.isnan(got.overrides), np.isnan(tab.overrides))
        assert np.array_equal(got.overrides[~np.isnan(got.overrides)], tab.overrides[~np.isnan(tab.overrides)])
    assert got.atol == tab.atol and got.rtol == tab.rtol and got.maxsteps == tab.maxsteps and got.stldet == tab.stldet


def test_hierarchical_tables_and_preequilibration(tmp_path):
    import h5write_min
    tab = synth.steadystate_hierarchical(6, seed=1)
    path = tmp_path / "hier.h5"
    h5write_min.write_parpe_input(str(path), tab)
    got = tables_from_hdf5(path, "steadystate_scaled")
    assert got.hier is not None
    for name in ("scaling_idx", "offset_idx", "sigma_idx", "scaling_map", "offset_map", "sigma_map"):
        assert np.array_equal(getattr(got.hier, name), getattr(tab.hier, name)), name
    # pre-equilibration: column 0 of simulationConditions selects further columns of k, the ragged
    # /fixedParameters/reinitializationIndices (variable-length int32 sequences in a global heap collection, as
    # h5py.vlen_dtype writes them for parPE: python/parpe/hdf5_pe_input.py:430-437) the states to re-initialise.
    # (No file with such a dataset written by the real HDF5 library exists in the reference tree: this part of
    # the reader is pinned on the committed writer's rendering of the format specification only.)
    ptab = synth.make_problem("boehm", 7, seed=2, preeq=True, posteq=1)
    p2 = tmp_path / "preeq.h5"
    h5write_min.write_parpe_input(str(p2), ptab)
    f = capi.H5File(p2)
    assert f.info("/fixedParameters/k")["shape"] == (ptab.nk, 7 + int(np.sum(ptab.has_preeq)))
    assert f.info("/fixedParameters/reinitializationIndices")["type_class"] == 9
    got = tables_from_hdf5(p2, "boehm")
    assert np.array_equal(got.has_preeq, ptab.has_preeq)
    assert np.array_equal(got.k, ptab.k) and np.array_equal(got.t, ptab.t) and np.isinf(got.t).sum() == 7
    for c in range(7):
        if ptab.has_preeq[c]:
            assert np.array_equal(got.k_preeq[c], ptab.k_preeq[c])
            assert got.reinit_idx[c] == list(ptab.reinit_idx[c])
        else:
            assert got.reinit_idx[c] == []
    # Newton's method for steady states: the reference uses it only if the MODEL's modes allow it (AMICI's default
    # integrationOnly ignores newton_maxsteps); (2, 2) = integrateIfNewtonFails for both is built, others are refused
    p3 = tmp_path / "newton.h5"
    h5write_min.write_parpe_input(str(p3), ptab, attrs=dict(newton_maxsteps=np.int32(40)))
    assert tables_from_hdf5(p3, "boehm").newton_maxsteps == 0
    h5write_min.write_parpe_input(str(p3), ptab, attrs=dict(newton_maxsteps=np.int32(40), steadyStateComputationMode=np.int32(2),
                                                           steadyStateSensitivityMode=np.int32(2)))
    assert tables_from_hdf5(p3, "boehm").newton_maxsteps == 40
    h5write_min.write_parpe_input(str(p3), ptab, attrs=dict(newton_maxsteps=np.int32(40), steadyStateComputationMode=np.int32(0)))
    with pytest.raises(capi.B200LibraryError, match="steadyStateComputationMode"):
        tables_from_hdf5(p3, "boehm")
    # newtonOnly computation mode (0, 0|2) with newton_maxsteps > 0: Newton's method without the simulation fallback
    for sens in (0, 2):
        h5write_min.write_parpe_input(str(p3), ptab, attrs=dict(newton_maxsteps=np.int32(40), steadyStateComputationMode=np.int32(0),
                                                               steadyStateSensitivityMode=np.int32(sens)))
        got = tables_from_hdf5(p3, "boehm")
        assert got.newton_maxsteps == 40 and got.newton_only == 1
    # the whole table of (computation mode, sensitivity mode, newton_maxsteps) -> what runs (csrc/h5input.cu;
    # SteadystateProblem::findSteadyState, deps/AMICI/src/steadystateproblem.cpp:120-166)
    table = {  # (comp, sens, nmax): (newton_maxsteps, newton_only, steady_sens_newton) or None = refused
        (1, 1, 0): (0, 0, 0), (1, 2, 30): (0, 0, 0), (1, 0, 30): (0, 0, 1), (2, 0, 0): (0, 0, 1), (2, 1, 0): (0, 0, 0),
        (2, 2, 30): (30, 0, 0), (2, 0, 30): (30, 0, 1), (2, 1, 30): None, (0, 0, 30): (30, 1, 0), (0, 2, 30): (30, 1, 0),
        (0, 1, 30): None, (0, 0, 0): None, (3, 1, 0): None, (1, 3, 0): None}
    for (comp, sens, nmax), want in table.items():
        h5write_min.write_parpe_input(str(p3), ptab, attrs=dict(newton_maxsteps=np.int32(nmax), steadyStateComputationMode=np.int32(comp),
                                                               steadyStateSensitivityMode=np.int32(sens)))
        if want is None:
            with pytest.raises(capi.B200LibraryError, match="steadyStateComputationMode"):
                tables_from_hdf5(p3, "boehm")
        else:
            got = tables_from_hdf5(p3, "boehm")
            assert (got.newton_maxsteps, got.newton_only, got.steady_sens_newton) == want, (comp, sens, nmax)
    # ... which has no role in a file without steady states
    plain = synth.make_problem("boehm", 3, seed=2)
    h5write_min.write_parpe_input(str(p3), plain, attrs=dict(newton_maxsteps=np.int32(30), steadyStateComputationMode=np.int32(0),
                                                             steadyStateSensitivityMode=np.int32(1)))
    got = tables_from_hdf5(p3, "boehm")
    assert (got.newton_maxsteps, got.newton_only, got.steady_sens_newton) == (0, 0, 0)
    with pytest.raises(capi.B200LibraryError):
        tables_from_hdf5(path, "no_such_model")


def test_algorithm_selecting_options_are_honoured_or_refused(tmp_path):
    """/amiciOptions attributes that select an algorithm (deps/AMICI/src/hdf5.cpp readSolverSettingsFromHDF5): what the
    path implements is accepted -- parPE's own generator writes sensi_meth = adjoint and linsol = KLU
    (python/parpe/hdf5_pe_input.py:859-879) --, the linear solver is mapped, anything that would change the results
    beyond the path's tolerance is refused with the attribute's name in the message."""
    import h5write_min
    i32, f64 = np.int32, np.float64
    tab = synth.make_problem("net32", 2, seed=3, scale="log10")
    path = tmp_path / "opts.h5"

    def load(t=tab, model="net32", **attrs):
        h5write_min.write_parpe_input(str(path), t, attrs=attrs)
        return tables_from_hdf5(path, model)

    # what parPE's generator writes
    got = load(linsol=i32(9), sensi=i32(1), sensi_meth=i32(2), sensi_meth_preeq=i32(2), steadyStateSensitivityMode=i32(1),
               nmaxevent=i32(10), maxstepsB=i32(5000))
    assert got.linsol == 0 and got.newton_maxsteps == 0 and got.newton_only == 0
    # linear solvers: dense variants -> dense inverse, sparse direct -> automatic, the rest refused
    assert load(linsol=i32(1)).linsol == 1 and load(linsol=i32(3)).linsol == 1 and load(linsol=i32(10)).linsol == 0
    for v in (2, 4, 5, 6, 7, 8, 0, 11):
        with pytest.raises(capi.B200LibraryError, match="linsol"):
            load(linsol=i32(v))
    # the sensitivity corrector variants and AMICI's defaults spelled out are accepted
    # ... the simultaneous one is honoured (pb200_problem_desc::ism: the full-feature instantiation of the kernels),
    # staggered1 runs the staggered corrector, an absent attribute means staggered (the north-star configuration)
    assert [load(ism=i32(ism)).ism for ism in (1, 2, 3)] == [1, 0, 0] and load().ism == 0
    load(lmm=i32(2), iter=i32(2), max_nonlin_iters=i32(3), max_conv_fails=i32(10), max_step_size=f64(0.0),
         rtol_fsa=f64(tab.rtol), atol_fsa=f64(tab.atol), interpType=i32(1), ordering=i32(1))
    for name, v in (("lmm", i32(1)), ("iter", i32(1)), ("ism", i32(4)), ("sensi_meth", i32(3)), ("max_nonlin_iters", i32(5)),
                    ("max_conv_fails", i32(20)), ("max_step_size", f64(0.5)), ("rtol_fsa", f64(1e-4)), ("atol_fsa", f64(1e-6))):
        with pytest.raises(capi.B200LibraryError, match=name):
            load(**{name: v})
    # steady-state options matter only for files with pre-/post-equilibration
    assert load(ss_tol_factor=f64(10.0), newton_step_steadystate_conv=i32(1)).n_conditions == 2
    ptab = synth.make_problem("boehm", 3, seed=2, preeq=True, posteq=1)
    load(ptab, "boehm", ss_tol_factor=f64(100.0), ss_atol=f64(100.0 * ptab.atol), ss_rtol=f64(100.0 * ptab.rtol),
         check_sensi_steadystate_conv=i32(1), newton_step_steadystate_conv=i32(0))
    for name, v in (("ss_tol_factor", f64(10.0)), ("ss_tol_sensi_factor", f64(1.0)), ("ss_atol", f64(1e-9)), ("ss_rtol_sensi", f64(1e-3)),
                    ("check_sensi_steadystate_conv", i32(0)), ("newton_step_steadystate_conv", i32(1))):
        with pytest.raises(capi.B200LibraryError, match=name):
            load(ptab, "boehm", **{name: v})
    # the damping mode matters only when Newton's method is used
    load(ptab, "boehm", newton_damping_factor_mode=i32(0))
    with pytest.raises(capi.B200LibraryError, match="newton_damping_factor_mode"):
        load(ptab, "boehm", newton_damping_factor_mode=i32(0), newton_maxsteps=i32(20), steadyStateComputationMode=i32(2),
             steadyStateSensitivityMode=i32(2))


def test_parameter_names_bounds_and_nominal_values(tmp_path):
    """/parameters/{parameterNames, lowerBound, upperBound, nominalValues}: what MultiConditionProblem's
    fillParametersMin / fillParametersMax and getProblemParameterIds read (multiConditionDataProvider.cpp:49-58,
    426-432); names as h5py writes a list of str (variable-length strings) and as fixed-length strings."""
    import h5write_min
    tab = synth.make_problem("boehm", 3, seed=1, condition_specific=2)
    tab.lower_bound = np.linspace(-5, -3, tab.n_opt)
    tab.upper_bound = np.linspace(3, 5, tab.n_opt)
    tab.nominal_values = np.linspace(-1, 1, tab.n_opt)
    tab.parameter_names = ["p%d_%s" % (j, "x" * (j % 4)) for j in range(tab.n_opt)]
    path = tmp_path / "names.h5"
    h5write_min.write_parpe_input(str(path), tab)
    got = tables_from_hdf5(path, "boehm")
    assert got.parameter_names == tab.parameter_names
    assert np.array_equal(got.lower_bound, tab.lower_bound) and np.array_equal(got.upper_bound, tab.upper_bound)
    assert np.array_equal(got.nominal_values, tab.nominal_values)
    f = capi.H5File(path)
    assert f.read_strings("/parameters/parameterNames") == tab.parameter_names
    assert f.info("/parameters/parameterNames")["type_class"] == 9
    # a file without them: None, not an error
    plain = synth.make_problem("boehm", 2, seed=1)
    p2 = tmp_path / "plain.h5"
    h5write_min.write_parpe_input(str(p2), plain)
    g2 = tables_from_hdf5(p2, "boehm")
    assert g2.parameter_names is None and g2.nominal_values is None and g2.lower_bound is not None
    # fixed-length strings (the result log's labels) through the same call
    log = capi.ResultLog()
    log.append_records("/multistarts/0", [3, 11], [1.0, 2.0], [0, 0], optimization_idx=2, iteration_idx=5)
    p3 = tmp_path / "log.h5"
    log.write(p3)
    assert capi.H5File(p3).read_strings("/multistarts/0/simulationLabel") == ["o2i5c3", "o2i5c11"]
    with pytest.raises(capi.B200LibraryError, match="strings"):
        capi.H5File(p3).read_strings("/multistarts/0/jobId")


@pytest.mark.gpu
def test_handle_from_input_file_matches_handle_from_tables(tmp_path, oracle_available):
    import h5write_min
    from oracle.binding import OracleProblem
    from tests.util import assert_grad_close, assert_llh_close
    tab = synth.make_problem("boehm", 6, seed=3, scale="mixed", holes=1, sigma_mode="mixed")
    path = tmp_path / "boehm.h5"
    h5write_min.write_parpe_input(str(path), tab)
    theta = synth.random_theta(tab, seed=1, width=0.3)
    a = B200SummedGradientFunction.from_hdf5(path, "boehm").evaluate(theta)
    b = B200SummedGradientFunction(tab).evaluate(theta)
    assert a[0] == b[0] == 0 and a[1] == b[1] and np.array_equal(a[2], b[2])
    ro = OracleProblem(tab).evaluate(theta)
    assert_llh_close(a[1], ro["fval"])
    assert_grad_close(a[2], ro["grad"])
    # a file with pre-equilibration conditions and t = inf output times: the handle from the file applies them
    ptab = synth.make_problem("boehm", 6, seed=5, preeq=True, posteq=1)
    pp = tmp_path / "boehm_preeq.h5"
    h5write_min.write_parpe_input(str(pp), ptab)
    a = B200SummedGradientFunction.from_hdf5(pp, "boehm").evaluate(theta)
    b = B200SummedGradientFunction(ptab).evaluate(theta)
    assert a[0] == b[0] == 0 and a[1] == b[1] and np.array_equal(a[2], b[2])
    ro = OracleProblem(ptab).evaluate(theta)
    assert_llh_close(a[1], ro["fval"])
    assert_grad_close(a[2], ro["grad"])
    # hierarchical tables ride along: pb200_create_from_hdf5 attaches them
    import ctypes as C
    htab = synth.fujita_hierarchical(12, seed=2)
    hp = tmp_path / "fujita.h5"
    h5write_min.write_parpe_input(str(hp), htab)
    h = C.c_void_p()
    capi.check(capi.lib().pb200_create_from_hdf5(str(hp).encode(), b"fujita", 0, C.byref(h)), "pb200_create_from_hdf5")
    assert capi.lib().pb200_hier_num_outer(h) == htab.n_opt - len(htab.hier.inner_idx)
    capi.lib().pb200_destroy(h)


def test_damaged_files_end_in_errors_not_in_out_of_bounds_reads(tmp_path):
    """The reader parses files it did not write: truncated and bit-flipped copies of a parPE input file (chunked +
    gzip datasets, variable-length sequences and strings, attributes) and of a result file go through every accessor
    of h5min.h in an ASan/UBSan build; each must end in an H5Error or in data, never in a memory error."""
    import subprocess
    import h5write_min
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("libasan not available")
    exe = str(tmp_path / "h5_robust")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-o", exe, os.path.join(ROOT, "tests", "host_harness", "h5_robust.cpp"), "-ldl"])
    # the writer (h5wmin.h) under the same sanitizers: every kind of object it emits, read back
    w = subprocess.run([exe, "--write", str(tmp_path / "written.h5")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1"))
    assert w.returncode == 0 and "WRITE-OK" in w.stdout, w.stdout + w.stderr[-2000:]
    tab = synth.make_problem("boehm", 5, seed=2, preeq=True, posteq=1, holes=1)
    tab.parameter_names = ["p%d" % j for j in range(tab.n_opt)]
    good = tmp_path / "good.h5"
    h5write_min.write_parpe_input(str(good), tab)
    log = capi.ResultLog()
    log.append_records("/multistarts/0", [1, 2, 3], [1.0, 2.0, 3.0], [0, 0, -4], sllh=np.ones((3, 4)))
    good2 = tmp_path / "good2.h5"
    log.write(good2)
    files = [str(good), str(good2)]
    rng = np.random.default_rng(int(os.environ.get("PB200_FUZZ_SEED", "0")))
    for src in (good, good2):
        raw = bytearray(src.read_bytes())
        for i in range(150):
            b = bytearray(raw)
            kind = i % 3
            if kind == 0:   # truncation
                b = b[:int(rng.integers(8, len(b)))]
            elif kind == 1:  # a few flipped bytes anywhere
                for pos in rng.integers(0, len(b), 4):
                    b[pos] ^= int(rng.integers(1, 256))
            else:           # an 8-byte field overwritten by a large number (addresses, sizes, counts)
                pos = int(rng.integers(0, len(b) - 8))
                b[pos:pos + 8] = int(rng.integers(0, 2 ** 62)).to_bytes(8, "little")
            p = tmp_path / ("m_%s_%d.h5" % (src.stem, i))
            p.write_bytes(bytes(b))
            files.append(str(p))
    r = subprocess.run([exe] + files, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:allocator_may_return_null=1"))
    assert r.returncode == 0 and "ROBUST-OK" in r.stdout, r.stdout[-500:] + r.stderr[-3000:]
    assert "opened=" in r.stdout and int(r.stdout.split("datasets=")[1]) > 100
