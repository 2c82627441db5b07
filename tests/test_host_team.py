"""The large-model integrator core (parpe_b200/csrc/bdf_team.cuh), compiled for the host with
the one-thread team, against the CVODES oracle and numpy on the same inputs: the algorithm the
CTA-per-condition kernel runs (blocked Gauss-Jordan inversion, solve GEMM through the pivot
order, sparse right-hand sides, staggered sensitivities) is parity-checked on every CPU run; the
GPU tests then check the CUDA build of the same source through the C ABI."""
import numpy as np
import pytest

from parpe_b200 import synth
from parpe_b200.models import get_spec
from parpe_b200.tables import SCALE_LN, SCALE_LOG10
from tests.util import assert_grad_close, assert_llh_close


def _dense_jacobian(spec, Jv):
    J = np.zeros((spec.nx, spec.nx))
    rows = np.repeat(np.arange(spec.nx), np.diff(spec.j_ptr))
    J[rows, spec.j_col] = Jv
    return J


@pytest.mark.parametrize("model,ncols,panel_widths", [("net32", 7, (0, 8, 4, 1)), ("net31", 6, (0, 8)),
                                                      ("chen", 51, (0, 16))])
def test_blocked_inversion_and_solve_gemm(model, ncols, panel_widths):
    """(I - gamma J)^-1 B for random values on the model's pattern vs numpy; every panel width
    must give the same answer (NB = 16 is what nx = 1200 uses on the device)."""
    from tests.host_harness.binding import team_solve
    spec = get_spec(model)
    rng = np.random.default_rng(0)
    Jv = 10.0 * rng.standard_normal(spec.nnz)
    gamma = 0.37
    A = np.eye(spec.nx) - gamma * _dense_jacobian(spec, Jv)
    B = rng.standard_normal((ncols, spec.nx))
    Xref = np.linalg.solve(A, B.T).T
    for nb in panel_widths:
        X = team_solve(model, Jv, gamma, B, panel_width=nb)
        assert np.max(np.abs(X - Xref)) <= 1e-11 * np.max(np.abs(Xref)), nb


def test_singular_newton_matrix_is_reported():
    from tests.host_harness.binding import team_solve
    spec = get_spec("net32")
    with pytest.raises(RuntimeError):
        # gamma = 0 and J = 0 would be the identity; a pattern-wide zero with gamma*J = I on the
        # diagonal makes the matrix exactly singular
        Jv = np.zeros(spec.nnz)
        rows = np.repeat(np.arange(spec.nx), np.diff(spec.j_ptr))
        Jv[rows == spec.j_col] = 1.0
        team_solve("net32", Jv, 1.0, np.ones((1, spec.nx)))


def _unscaled_p(tab, c, theta, oracle):
    p_sim = oracle.map_parameters(c, theta)
    out = np.empty_like(p_sim)
    for i, v in enumerate(p_sim):
        s = tab.pscale_sim[c, i]
        out[i] = 10.0 ** v if s == SCALE_LOG10 else (np.exp(v) if s == SCALE_LN else v)
    return out


@pytest.mark.parametrize("kw", [dict(scale="log10"), dict(scale="mixed", sigma_mode="mixed", missing_fraction=0.1)],
                         ids=["log10", "mixed"])
def test_host_team_core_matches_cvodes(oracle_available, kw):
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_team
    model = "net32"
    tab = synth.make_problem(model, 3, seed=7, **kw)
    ora = OracleProblem(tab, 2, 1)  # staggered, dense
    theta = synth.random_theta(tab, seed=3, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p = _unscaled_p(tab, c, theta, ora)
        plist = ora.plist(c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate_team(model, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), rtol=tab.rtol,
                          atol=tab.atol, panel_width=8 if c == 1 else 0)
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        fac = np.array([p[i] * np.log(10) if tab.pscale_sim[c, i] == SCALE_LOG10 else
                        (p[i] if tab.pscale_sim[c, i] == SCALE_LN else 1.0) for i in plist])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])
        nst_ref = ro["stats"][c, 0]
        assert abs(r["stats"][0] - nst_ref) <= 0.1 * nst_ref + 5
        # cost only
        r0 = simulate_team(model, p, tab.k[c], [], tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), rtol=tab.rtol,
                           atol=tab.atol)
        assert r0["status"] == 0
        assert_llh_close(r0["llh"], ora.evaluate(theta, idx=[c], want_grad=False)["llh"][0])


def test_network_tables_consistent():
    """The emitted tables describe what the generator says: J is the derivative of xdot (finite
    differences through the oracle-independent numpy restatement of the evaluator)."""
    spec = get_spec("net32")
    rng = np.random.default_rng(1)
    x = rng.uniform(0.1, 2.0, spec.nx)
    p = rng.uniform(0.5, 2.0, spec.np)

    def xdot(x, p):
        w = spec.kappa * p[spec.pidx]
        w = w * np.where(spec.r_a >= 0, x[np.maximum(spec.r_a, 0)], 1.0) * np.where(spec.r_b >= 0, x[np.maximum(spec.r_b, 0)], 1.0)
        out = np.zeros(spec.nx)
        for i in range(spec.nx):
            e = slice(spec.s_ptr[i], spec.s_ptr[i + 1])
            out[i] = np.sum(spec.s_coef[e] * w[spec.s_reac[e]])
        return out

    Jv = np.zeros(spec.nnz)
    for e in range(spec.nnz):
        for t in range(spec.jt_ptr[e], spec.jt_ptr[e + 1]):
            r = spec.jt_reac[t]
            Jv[e] += spec.jt_coef[t] * spec.kappa[r] * p[spec.pidx[r]] * (x[spec.jt_other[t]] if spec.jt_other[t] >= 0 else 1.0)
    J = _dense_jacobian(spec, Jv)
    Jfd = np.zeros_like(J)
    for j in range(spec.nx):
        d = np.zeros(spec.nx)
        d[j] = 1e-6
        Jfd[:, j] = (xdot(x + d, p) - xdot(x - d, p)) / 2e-6
    assert np.max(np.abs(J - Jfd)) <= 1e-6 * np.max(np.abs(J))
    # CSC view addresses the same entries
    rows = np.repeat(np.arange(spec.nx), np.diff(spec.j_ptr))
    cols = np.repeat(np.arange(spec.nx), np.diff(spec.csc_ptr))
    assert np.array_equal(spec.csc_row, rows[spec.csc_from_csr])
    assert np.array_equal(cols, spec.j_col[spec.csc_from_csr])


def test_network_evaluator_against_independent_integration(oracle_available):
    """The oracle's network model and the device share one evaluator (net_model.cuh), so it is pinned
    separately: a numpy restatement of the mass-action right-hand side, integrated by scipy's BDF,
    must reproduce the observables the oracle (AMICI/CVODES over the C evaluator) returns."""
    from scipy.integrate import solve_ivp
    from oracle.binding import OracleProblem
    model = "net32"
    spec = get_spec(model)
    tab = synth.make_problem(model, 2, seed=3, timepoints=synth.DEFAULT_TIMEPOINTS[model][:4])
    theta = synth.random_theta(tab, seed=2, width=0.3)
    ora = OracleProblem(tab)
    rc, y_ora, _ = ora.get_model_outputs(theta)
    assert rc == 0
    ra, rb = spec.r_a, spec.r_b

    def rates(x, p):
        w = spec.kappa * p[spec.pidx]
        return w * np.where(ra >= 0, x[np.maximum(ra, 0)], 1.0) * np.where(rb >= 0, x[np.maximum(rb, 0)], 1.0)

    S = np.zeros((spec.nx, spec.nr))
    for i in range(spec.nx):
        for e in range(spec.s_ptr[i], spec.s_ptr[i + 1]):
            S[i, spec.s_reac[e]] += spec.s_coef[e]
    O = np.zeros((spec.ny, spec.nx))
    for j in range(spec.ny):
        for e in range(spec.o_ptr[j], spec.o_ptr[j + 1]):
            O[j, spec.o_idx[e]] += spec.o_w[e]
    for c in range(tab.n_conditions):
        p = _unscaled_p(tab, c, theta, ora)
        k = tab.k[c]
        x0 = spec.x0_base * np.where(spec.x0_kidx >= 0, k[np.maximum(spec.x0_kidx, 0)], 1.0)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        t = tab.t[lo:hi]
        sol = solve_ivp(lambda _t, x: S @ rates(x, p), (0.0, t[-1]), x0, method="BDF", t_eval=t, rtol=1e-10, atol=1e-13)
        assert sol.success
        y_ref = (O @ sol.y).T
        np.testing.assert_allclose(y_ora[lo * spec.ny:hi * spec.ny].reshape(-1, spec.ny), y_ref, rtol=2e-6, atol=1e-9)


def test_host_team_output_time_edge_cases(oracle_available):
    """Output times at t0 (no integration, data point on the initial state) and repeated output times
    (second one within round-off of the current time: no further step), as ForwardProblem handles them
    (forwardproblem.cpp:23-35,152-159)."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_team
    tab = synth.make_problem("net32", 2, seed=7, timepoints=np.array([0.0, 0.5, 0.5, 2.0]))
    theta = synth.random_theta(tab, seed=3, width=0.3)
    ora = OracleProblem(tab, 2, 1)
    ro = ora.evaluate(theta)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p = _unscaled_p(tab, c, theta, ora)
        plist = ora.plist(c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate_team("net32", p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c))
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        fac = np.array([p[i] * np.log(10) for i in plist])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])


# ---- sparse fixed-pattern LU + dense Schur block (sparse_lu.h; the reference's KLU role) ----------
def _diag_dominant_values(spec, rng):
    """Random values on the pattern whose I - gamma*J keeps usable diagonal pivots (the sparse part pivots
    on the diagonal, as KLU does unless |diag| < 0.001 max|column|)."""
    Jv = rng.standard_normal(spec.nnz)
    rows = np.repeat(np.arange(spec.nx), np.diff(spec.j_ptr))
    Jv[rows == spec.j_col] = -np.abs(Jv[rows == spec.j_col]) - 3.0   # degradation-like diagonal
    return Jv


@pytest.mark.parametrize("model,ncols,deg_max,stage", [("net32", 7, 0, 0), ("net32", 7, 3, 0), ("net32", 7, 100, 0),
                                                       ("net31", 6, 4, 0), ("net32", 7, 0, 70),
                                                       ("chen", 51, 0, 0), ("chen", 51, 0, 6000), ("froehlich", 4, 0, 0)])
def test_sparse_lu_with_dense_schur_block(model, ncols, deg_max, stage):
    """(I - gamma J)^-1 B through the sparse factors (entry-wise numeric phase over dependency levels,
    level-scheduled triangular solves) and the dense inverse of the Schur block vs numpy, for several
    splits (deg_max = 100 on net32: everything sparse, m = 0) and column batch sizes."""
    from tests.host_harness.binding import team_solve
    spec = get_spec(model)
    rng = np.random.default_rng(0)
    Jv = _diag_dominant_values(spec, rng)
    gamma = 0.37
    A = np.eye(spec.nx) - gamma * _dense_jacobian(spec, Jv)
    B = rng.standard_normal((ncols, spec.nx))
    Xref = np.linalg.solve(A, B.T).T
    X, info = team_solve(model, Jv, gamma, B, sparse=True, deg_max=deg_max, stage_doubles=stage, want_info=True)
    assert info["n1"] + info["m"] == spec.nx and info["n1"] > 0
    if deg_max == 100:
        assert info["m"] == 0
    assert np.max(np.abs(X - Xref)) <= 1e-10 * np.max(np.abs(Xref)), info
    # the dense path gives the same answer
    Xd = team_solve(model, Jv, gamma, B)
    assert np.max(np.abs(X - Xd)) <= 1e-10 * np.max(np.abs(Xref))


def test_sparse_lu_symbolic_matches_superlu_fill():
    """The symbolic phase on the Chen-shape pattern: minimum degree on J + J^T leaves a dense block of
    about half the matrix (no small separators in a random network) and a sparse part whose factors are
    far smaller than the matrix; counts are what DESIGN.md's roofline uses."""
    from tests.host_harness.binding import team_solve
    spec = get_spec("chen")
    rng = np.random.default_rng(1)
    _, info = team_solve("chen", _diag_dominant_values(spec, rng), 0.1, rng.standard_normal((1, spec.nx)), sparse=True,
                         want_info=True)
    assert 200 <= info["m"] <= 300 and info["m"] % 2 == 0
    assert info["nnzL"] + info["nnzU"] < 0.05 * spec.nx ** 2
    assert info["fwd_levels"] < info["n1"] and info["bwd_levels"] < info["n1"]


def test_sparse_lu_weak_pivot_falls_back_to_dense():
    """A diagonal pivot below 0.001 of an entry of its column (KLU would leave the diagonal there) is
    reported by the sparse numeric phase (return code 2 -> 12 through the harness), so that lsetup redoes
    the matrix by the dense path with partial pivoting."""
    from tests.host_harness.binding import team_solve
    spec = get_spec("net32")
    rng = np.random.default_rng(2)
    Jv = _diag_dominant_values(spec, rng)
    rows = np.repeat(np.arange(spec.nx), np.diff(spec.j_ptr))
    gamma = 0.5
    Jv[rows == spec.j_col] = 1.0 / gamma - 1e-9   # I - gamma*J has a (nearly) zero diagonal
    with pytest.raises(RuntimeError, match="12"):
        team_solve("net32", Jv, gamma, np.ones((1, spec.nx)), sparse=True)
    X = team_solve("net32", Jv, gamma, np.ones((1, spec.nx)))  # dense path: partial pivoting copes
    A = np.eye(spec.nx) - gamma * _dense_jacobian(spec, Jv)
    np.testing.assert_allclose(A @ X[0], np.ones(spec.nx), rtol=0, atol=1e-8)


@pytest.mark.parametrize("model,deg_max,stage", [("net32", 0, 0), ("net32", 3, 150), ("net32", 100, 0)])
def test_host_team_sparse_path_matches_cvodes(oracle_available, model, deg_max, stage):
    """Full simulations with the sparse linear solver against the CVODES oracle (KLU), and against
    the dense path of the same core: same step counts within a few steps, results within tolerance."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_team
    tab = synth.make_problem(model, 2, seed=7, scale="mixed", sigma_mode="mixed", missing_fraction=0.1)
    ora = OracleProblem(tab, 2, 9)  # staggered, KLU (the reference default)
    theta = synth.random_theta(tab, seed=3, width=0.3)
    ro = ora.evaluate(theta, collect_stats=True)
    assert ro["status"] == 0
    for c in range(tab.n_conditions):
        p = _unscaled_p(tab, c, theta, ora)
        plist = ora.plist(c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        args = (model, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c))
        r = simulate_team(*args, rtol=tab.rtol, atol=tab.atol, sparse=True, deg_max=deg_max, stage_doubles=stage)
        rd = simulate_team(*args, rtol=tab.rtol, atol=tab.atol)
        assert r["status"] == 0
        assert r["stats"][16] == 0  # no fallback to the dense path
        assert_llh_close(r["llh"], ro["llh"][c])
        fac = np.array([p[i] * np.log(10) if tab.pscale_sim[c, i] == SCALE_LOG10 else
                        (p[i] if tab.pscale_sim[c, i] == SCALE_LN else 1.0) for i in plist])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])
        assert abs(int(r["stats"][0]) - int(rd["stats"][0])) <= 0.05 * rd["stats"][0] + 3


@pytest.mark.parametrize("ns_mode", ["sens", "cost"])
def test_fused_history_pass_is_bit_identical(ns_mode):
    """advance_fused (cvCompleteStep + error weights + cvAdjustOrder + cvRescale + cvPredict in ONE pass over
    the history between the steps of a CVode call) performs CVODES' operations per element in CVODES' order:
    the simulation is bit-identical to the one with separate sweeps -- same steps, orders, llh, sllh."""
    from tests.host_harness.binding import simulate_team
    tab = synth.make_problem("net32", 2, seed=7, scale="log10")
    spec = get_spec("net32")
    rng = np.random.default_rng(5)
    for c in range(2):
        p = np.asarray(spec.p_nominal) * 10.0 ** rng.uniform(-0.3, 0.3, spec.np)
        plist = list(range(spec.np)) if ns_mode == "sens" else []
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        args = ("net32", p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c))
        a = simulate_team(*args, fuse_history=True)
        b = simulate_team(*args, fuse_history=False)
        assert a["status"] == b["status"] == 0
        assert np.array_equal(a["stats"][:14], b["stats"][:14])
        assert np.array_equal(a["numsteps"], b["numsteps"]) and np.array_equal(a["order"], b["order"])
        assert a["llh"] == b["llh"] and np.array_equal(a["sllh"], b["sllh"])
        assert a["stats"][0] > 500 and a["stats"][11] >= 0  # a long run with order changes up and down


@pytest.mark.parametrize("stage", [0, 150, 40])
def test_fused_newton_iteration_of_the_sparse_path(stage):
    """Sparse solver: right-hand sides + forward substitution in one staging pass (rhs_sens_fwd) and backward
    substitution + acor update + WRMS norm in another (sparse_backward(update)) -- same arithmetic per element
    as the separate sweeps, so the one-thread host build must reproduce them bit for bit, for any batch size
    (stage = 40 doubles < 2 nx: the fused right-hand side is not possible and the plain one runs)."""
    from tests.host_harness.binding import simulate_team
    tab = synth.make_problem("net32", 2, seed=7, scale="log10")
    spec = get_spec("net32")
    rng = np.random.default_rng(6)
    for c in range(2):
        p = np.asarray(spec.p_nominal) * 10.0 ** rng.uniform(-0.3, 0.3, spec.np)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        args = ("net32", p, tab.k[c], list(range(spec.np)), tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c))
        a = simulate_team(*args, sparse=True, stage_doubles=stage, fuse_newton=True)
        b = simulate_team(*args, sparse=True, stage_doubles=stage, fuse_newton=False)
        assert a["status"] == b["status"] == 0
        assert np.array_equal(a["stats"][:14], b["stats"][:14])
        assert a["llh"] == b["llh"] and np.array_equal(a["sllh"], b["sllh"])


@pytest.mark.parametrize("model", ["net32", "net31"])
@pytest.mark.parametrize("sparse", [False, True], ids=["dense", "sparse"])
def test_simultaneous_corrector_team(oracle_available, model, sparse):
    """pb200_problem_desc::ism = 1 for the large-model core (BdfTeam::nls_simultaneous, steady-state instantiation,
    one-thread team): state and sensitivities in one Newton iteration with the joint norm and one error test
    (cvodes_nls_sim.c) -- against the reference run with InternalSensitivityMethod::simultaneous, both linear
    solvers; CVODES' counters of that corrector (no nniS / netfS), step counts within a few per cent."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_team
    from tests.util import assert_grad_close, assert_llh_close, condition_inputs
    tab = synth.make_problem(model, 3, seed=0, scale="log10")
    theta = synth.random_theta(tab, seed=1, width=0.3)
    tab.ism = 1
    ro = OracleProblem(tab, 2, 9).evaluate(theta, collect_stats=True)  # (the descriptor's ism wins)
    stag = OracleProblem(synth.make_problem(model, 3, seed=0, scale="log10"), 2, 9).evaluate(theta, collect_stats=True)
    assert ro["status"] == 0 and np.any(ro["stats"][:, 0] != stag["stats"][:, 0])
    for c in range(tab.n_conditions):
        p, plist, fac = condition_inputs(tab, theta, c)
        lo, hi = tab.t_offsets[c], tab.t_offsets[c + 1]
        r = simulate_team(model, p, tab.k[c], plist, tab.t[lo:hi], tab.y_of(c), tab.sigma_of(c), errconS=5, steady=True,
                          sparse=sparse)
        assert r["status"] == 0
        assert_llh_close(r["llh"], ro["llh"][c])
        assert_grad_close(r["sllh"] * fac, ro["sllh"][c][:len(plist)])
        assert abs(r["numsteps"][-1] - ro["stats"][c, 0]) <= 0.1 * ro["stats"][c, 0] + 2
        assert r["stats"][5] == 0 and r["stats"][8] == 0 and r["stats"][4] > 0


def test_simultaneous_corrector_team_chen_shape(oracle_available):
    """The same on the Chen shape (nx = 500, 50 sensitivity columns, sparse LU + dense Schur block) over the first two
    output times: the reference's step count exactly."""
    from oracle.binding import OracleProblem
    from tests.host_harness.binding import simulate_team
    from tests.util import assert_grad_close, assert_llh_close, condition_inputs
    tab = synth.make_problem("chen", 1, seed=16, scale="log10", timepoints=[0.1, 0.3])
    theta = synth.random_theta(tab, seed=5, width=0.3)
    tab.ism = 1
    ro = OracleProblem(tab, 2, 9).evaluate(theta, collect_stats=True)
    p, plist, fac = condition_inputs(tab, theta, 0)
    r = simulate_team("chen", p, tab.k[0], plist, tab.t, tab.y_of(0), tab.sigma_of(0), errconS=5, steady=True, sparse=True)
    assert r["status"] == 0 and ro["status"] == 0
    assert_llh_close(r["llh"], ro["llh"][0])
    assert_grad_close(r["sllh"] * fac, ro["sllh"][0][:len(plist)])
    assert abs(r["numsteps"][-1] - ro["stats"][0, 0]) <= 2
