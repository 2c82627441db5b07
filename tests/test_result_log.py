"""Per-simulation result logging: the records parPE's logSimulation / saveSimulation writes per
simulation job (src/parpeamici/multiConditionProblem.cpp:254-325) and the standalone simulator's result datasets
(doc/hdf5.md:190-252), written without an HDF5 library (parpe_b200/csrc/h5wmin.h) and read back through the
independent reader (h5min.h, itself pinned on a file written by the real HDF5 library: tests/test_h5_reader.py).
No file written by the reference exists in its tree and h5py / libhdf5 are absent from this image: that a real HDF5
library opens these files is NOT verified here ("parity unpinned"); names, types and shapes follow doc/hdf5.md."""
import numpy as np
import pytest

from parpe_b200 import capi, synth


def test_records_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    log = capi.ResultLog()
    np_, n1, n2 = 5, 4, 3
    rec = []
    for it, n in enumerate((n1, n2)):
        cidx = rng.integers(0, 10, n)
        llh = rng.standard_normal(n)
        sllh = rng.standard_normal((n, np_))
        sllh[:, 3:] = np.nan  # nplist = 3 < np: NaN padded, as saveSimulation pads
        par = rng.standard_normal((n, np_))
        sec = rng.random(n)
        status = np.array([0, -4, 0, 0][:n], dtype=np.int32)
        stats = rng.integers(0, 1000, (n, 20))
        stats[:, 12] = 1 + np.arange(n) % 3
        assert log.append_records("/multistarts/0", cidx, llh, status, sllh, par, sec, stats, optimization_idx=0,
                                  iteration_idx=it) == n
        rec.append((cidx, llh, sllh, par, sec, status, stats))
    # a second optimiser run goes to its own group
    log.append_records("/multistarts/1", [7], [1.5], [0], np.full((1, np_), np.nan), np.ones((1, np_)), [0.25], None,
                       optimization_idx=1)
    assert log.num_records("/multistarts/0") == n1 + n2 and log.num_records("/multistarts/1") == 1
    path = tmp_path / "results.h5"
    log.write(path)
    f = capi.H5File(path)
    assert f.list("/") == ["multistarts"] and f.list("/multistarts") == ["0", "1"]
    g = "/multistarts/0"
    assert f.list(g) == sorted(["jobId", "simulationLabel", "simulationLogLikelihood", "simulationLogLikelihoodGradient",
                                "simulationParameters", "simulationWallTimeInSec", "simulationStatus", "simulationTrials",
                                "simulationIntegratorStatistics"])
    n = n1 + n2
    cat = lambda k: np.concatenate([r[k] for r in rec])
    # shapes of doc/hdf5.md:204-224: one COLUMN per simulation job
    assert f.info(g + "/simulationLogLikelihood") == dict(shape=(1, n), type_class=1, type_size=8)
    assert f.info(g + "/simulationLogLikelihoodGradient")["shape"] == (np_, n)
    assert f.info(g + "/simulationParameters")["shape"] == (np_, n)
    assert f.info(g + "/jobId") == dict(shape=(1, n), type_class=0, type_size=4)
    assert f.info(g + "/simulationStatus") == dict(shape=(1, n), type_class=0, type_size=4)
    assert np.array_equal(f.read(g + "/simulationLogLikelihood")[0], cat(1))
    got = f.read(g + "/simulationLogLikelihoodGradient").T
    ref = cat(2)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])
    assert np.array_equal(f.read(g + "/simulationParameters").T, cat(3))
    assert np.array_equal(f.read(g + "/simulationWallTimeInSec")[0], cat(4))
    assert np.array_equal(f.read(g + "/simulationStatus")[0], cat(5))
    assert np.array_equal(f.read(g + "/jobId")[0], np.arange(n))  # running number over the whole log
    assert np.array_equal(f.read("/multistarts/1/jobId")[0], [n])
    labels = f.read_strings(g + "/simulationLabel")
    assert labels == ["o0i0c%d" % c for c in rec[0][0]] + ["o0i1c%d" % c for c in rec[1][0]]
    assert f.read_strings("/multistarts/1/simulationLabel") == ["o1i0c7"]
    # the tolerance-relaxation retry statistics (beyond the reference's datasets)
    assert np.array_equal(f.read(g + "/simulationTrials")[0], cat(6)[:, 12])
    assert f.info(g + "/simulationIntegratorStatistics") == dict(shape=(14, n), type_class=0, type_size=8)
    assert np.array_equal(f.read(g + "/simulationIntegratorStatistics").T, cat(6)[:, :14])
    assert np.all(f.read("/multistarts/1/simulationTrials") == 0)


def test_many_groups_and_empty_log(tmp_path):
    """more children in one group than a default symbol node holds; an empty log is a valid file"""
    log = capi.ResultLog()
    for s in range(40):
        log.append_records("/multistarts/%d" % s, [s], [float(s)], [0])
    p = tmp_path / "many.h5"
    log.write(p)
    f = capi.H5File(p)
    assert f.list("/multistarts") == sorted(str(s) for s in range(40))
    assert f.read("/multistarts/17/simulationLogLikelihood")[0, 0] == 17.0
    assert f.info("/multistarts/17/simulationParameters")["shape"] == (0, 1)
    empty = capi.ResultLog()
    p2 = tmp_path / "empty.h5"
    empty.write(p2)
    assert capi.H5File(p2).list("/") == []
    with pytest.raises(capi.B200LibraryError):
        log.write(tmp_path / "no_such_dir" / "x.h5")


@pytest.mark.gpu
def test_log_of_an_evaluation_and_simulation_result_file(tmp_path, oracle_available):
    from oracle.binding import OracleProblem
    from parpe_b200.problem import B200SummedGradientFunction
    from tests.util import assert_grad_close, assert_llh_close, assert_outputs_close
    tab = synth.make_problem("boehm", 6, seed=3, scale="mixed", holes=1, sigma_mode="mixed")
    theta = synth.random_theta(tab, seed=1, width=0.3)
    fun = B200SummedGradientFunction(tab)
    ora = OracleProblem(tab)
    log = capi.ResultLog()
    sub = [4, 1, 1, 5]
    st, fval, grad, _ = fun.evaluate(theta, sub)
    assert log.append(fun, "/multistarts/0", 0, 0) == len(sub)
    st, fval0, _, _ = fun.evaluate(theta, want_gradient=False)
    assert log.append(fun, "/multistarts/0", 0, 1) == tab.n_conditions
    path = tmp_path / "log.h5"
    log.write(path)
    f = capi.H5File(path)
    g = "/multistarts/0"
    n = len(sub) + tab.n_conditions
    ro = ora.evaluate(theta, idx=sub)
    llh = f.read(g + "/simulationLogLikelihood")[0]
    assert llh.shape == (n,)
    assert_llh_close(llh[:len(sub)], ro["llh"])
    assert_llh_close(-np.sum(llh[len(sub):]), fval0)
    sl = f.read(g + "/simulationLogLikelihoodGradient").T
    par = f.read(g + "/simulationParameters").T
    for r, c in enumerate(sub):
        k = int(np.sum(tab.mapping[c] >= 0))  # nplist of the condition: sllh is plist-compacted, NaN padded
        assert_grad_close(sl[r, :k], ro["sllh"][r][:k])
        assert np.all(np.isnan(sl[r, k:]))
        p_sim = np.zeros(tab.np_)
        from oracle import binding as ob
        ob.lib().ora_map_parameters(ora._h, c, ob._dp(np.ascontiguousarray(theta)), ob._dp(p_sim))
        assert np.array_equal(par[r], p_sim)  # the condition's model parameters in simulation scale, bit for bit
    assert np.all(np.isnan(sl[len(sub):]))  # no gradient requested
    assert f.read_strings(g + "/simulationLabel") == ["o0i0c%d" % c for c in sub] + ["o0i1c%d" % c for c in range(6)]
    assert np.all(f.read(g + "/simulationStatus") == 0) and np.all(f.read(g + "/simulationTrials") == 1)
    stats = f.read(g + "/simulationIntegratorStatistics")
    assert np.all(stats[0] > 100)  # nst
    assert np.all(f.read(g + "/simulationWallTimeInSec") > 0)
    # simulation result file: llh, parameters, t / yMes / ySim per condition (doc/hdf5.md:228-252)
    sim = tmp_path / "sim.h5"
    rc = capi.check(capi.lib().pb200_write_simulation_results(fun._h, capi.dptr(np.ascontiguousarray(theta)), len(theta),
                                                              str(sim).encode(), b"/multistarts/0"), "write")
    assert rc == 0
    s = capi.H5File(sim)
    assert s.list(g) == ["llh", "parameters", "t", "yMes", "ySim"]
    assert np.array_equal(s.read(g + "/parameters"), theta)
    rall = ora.evaluate(theta, want_grad=False)
    assert_llh_close(s.read(g + "/llh"), rall["llh"])
    rc, yo, _ = ora.get_model_outputs(theta)
    ysim = np.concatenate([s.read(g + "/ySim/%d" % c).reshape(-1) for c in range(6)])
    assert_outputs_close(ysim, yo, tab)
    for c in range(6):
        assert np.array_equal(s.read(g + "/t/%d" % c), tab.t[tab.t_offsets[c]:tab.t_offsets[c + 1]])
        a, b = s.read(g + "/yMes/%d" % c), tab.y_of(c)
        assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
