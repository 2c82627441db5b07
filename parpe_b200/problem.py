"""Host-side mirror of the reference's interface for the hot path.

``B200SummedGradientFunction`` plays the role of ``parpe::AmiciSummedGradientFunction``
(``include/parpeamici/multiConditionProblem.h:95-237``): same method names and argument
meaning (``evaluate(parameters, datasets) -> (status, fval, gradient)``,
``numParameters``, ``getModelOutputsAndSigmas``, ``getAllMeasurements``,
``getParameterScaling``), same error behaviour (simulation failure => ``fval = +inf`` and
``functionEvaluationFailure``; wrong parameter-vector size raises).  It forwards to the
C ABI; the numerics run in the CUDA library only.  The C++ adapter with the reference's
exact signatures is ``parpe_b200/include/parpe_b200/b200SummedGradientFunction.h``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .tables import PB200_NSTATS, STAT_NAMES, ProblemTables

functionEvaluationSuccess = 0
functionEvaluationFailure = 1


class B200SummedGradientFunction:
    """``parpe::SummedGradientFunction<int>`` over a CUDA-resident multi-condition problem."""

    def __init__(self, tables: ProblemTables, device: int = 0):
        self.tables = tables
        self._desc = tables.to_desc()
        self._h = C.c_void_p()
        L = capi.lib()
        capi.check(L.pb200_create(C.byref(self._desc), device, C.byref(self._h)), "pb200_create")
        self.device = device
        self.rank, self.world = 0, 1
        if tables.hier is not None:
            self._hd = tables.hier.to_desc()
            capi.check(L.pb200_hier_attach(self._h, C.byref(self._hd)), "pb200_hier_attach")
        pre = tables.preeq_args()
        if pre is not None:
            flag, kp, off, flat = pre
            capi.check(L.pb200_set_preequilibration(self._h, flag.ctypes.data_as(C.POINTER(C.c_int32)), capi.dptr(kp),
                                                    off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                    flat.ctypes.data_as(C.POINTER(C.c_int32))),
                       "pb200_set_preequilibration")

    @classmethod
    def from_hdf5(cls, path, model_name: str, device: int = 0):
        """The data-provider role: a handle straight from a parPE parameter-estimation input file
        (doc/hdf5.md; pb200_create_from_hdf5).  ``tables`` is rebuilt from the file for the Python side."""
        tab = tables_from_hdf5(path, model_name)
        return cls(tab, device=device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            capi.lib().pb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parpe::SummedGradientFunction<int> ------------------------------------
    def numParameters(self) -> int:
        return self.tables.n_opt

    def getParameterScaling(self, parameterIndex: int) -> int:
        return int(self.tables.pscale_opt[parameterIndex])

    def getAllMeasurements(self):
        T = self.tables
        return [T.y_of(c).reshape(-1).copy() for c in range(T.n_conditions)]

    def evaluate(self, parameters: Sequence[float], datasets=None, want_gradient: bool = True):
        """-> (status, fval, gradient or None, seconds).  ``datasets``: one index or a list;
        default = all conditions (SummedGradientFunctionGradientFunctionAdapter,
        include/parpecommon/functions.h:144-153)."""
        T = self.tables
        theta = np.ascontiguousarray(parameters, dtype=np.float64)
        if datasets is None:
            idx = np.arange(T.n_conditions, dtype=np.int32)
        else:
            idx = np.ascontiguousarray(np.atleast_1d(datasets), dtype=np.int32)
        fval = C.c_double()
        sec = C.c_double()
        grad = np.zeros(T.n_opt) if want_gradient else None
        rc = capi.lib().pb200_evaluate(self._h, capi.dptr(theta), len(theta), capi.iptr(idx), len(idx),
                                       C.byref(fval), capi.dptr(grad), C.byref(sec))
        capi.check(rc, "pb200_evaluate")
        return rc, fval.value, grad, sec.value

    def evaluate_device(self, parameters, datasets, want_gradient: bool, d_out_ptr: int, stream_ptr: int = 0):
        """Enqueue on ``stream``; partial (nllh, nerr, grad) lands in the device buffer."""
        theta = np.ascontiguousarray(parameters, dtype=np.float64)
        idx = np.ascontiguousarray(np.atleast_1d(datasets), dtype=np.int32)
        rc = capi.lib().pb200_evaluate_device(self._h, capi.dptr(theta), len(theta), capi.iptr(idx), len(idx),
                                              int(want_gradient), C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr))
        capi.check(rc, "pb200_evaluate_device")

    def getModelOutputsAndSigmas(self, parameters):
        T = self.tables
        theta = np.ascontiguousarray(parameters, dtype=np.float64)
        y = np.zeros_like(T.y)
        s = np.zeros_like(T.y)
        rc = capi.check(capi.lib().pb200_get_model_outputs(self._h, capi.dptr(theta), len(theta), capi.dptr(y),
                                                           capi.dptr(s)), "pb200_get_model_outputs")
        return rc, y, s

    # ---- beyond the reference interface -------------------------------------------
    def set_partition(self, rank: int, world: int, partial_sums: bool = False):
        """Shard every index list over ``world`` handles.  ``evaluate`` then returns the GLOBAL result and
        needs a collective (``nccl_init`` or ``set_allreduce``); ``partial_sums=True`` makes it return the
        rank-local sums instead (tests)."""
        capi.check(capi.lib().pb200_set_partition(self._h, rank, world), "pb200_set_partition")
        self.set_option(capi.OPT_PARTIAL_SUMS, int(partial_sums))
        self.rank, self.world = rank, world

    def set_option(self, key: int, value: int):
        capi.check(capi.lib().pb200_set_option(self._h, key, value), "pb200_set_option")

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().pb200_nccl_unique_id(buf), "pb200_nccl_unique_id")
        return buf.raw

    def nccl_init(self, unique_id: bytes, rank: int, world: int):
        """In-library NCCL communicator over the ``world`` handles of a partitioned problem (also sets
        the partition): evaluate / evaluate_device / hier_evaluate then return global sums."""
        assert len(unique_id) == 128
        capi.check(capi.lib().pb200_nccl_init(self._h, unique_id, rank, world), "pb200_nccl_init")
        self.rank, self.world = rank, world

    def last_transfer_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        capi.check(capi.lib().pb200_last_transfer_bytes(self._h, C.byref(a), C.byref(b)), "pb200_last_transfer_bytes")
        return a.value, b.value

    def linsol_info(self):
        """Linear-solver description of the large-model path (symbolic analysis done at create)."""
        v = (C.c_int64 * 16)()
        n = capi.check(capi.lib().pb200_linsol_info(self._h, v, 16), "pb200_linsol_info")
        names = ["kind", "nx", "n_sparse", "m_dense", "nnz_J", "nnz_L", "nnz_U", "updates", "levels_fwd",
                 "levels_bwd", "factor_flops", "solve_flops", "factor_bytes", "solve_bytes"]
        return {k: int(v[i]) for i, k in enumerate(names[:n])}

    def set_allreduce(self, fn):
        """Collective hook of partitioned hierarchical evaluations: ``fn(array, op)`` all-reduces the
        numpy view ``array`` in place over the ranks (op: ``capi.REDUCE_SUM`` / ``capi.REDUCE_MAX``);
        ``None`` removes it.  With torch.distributed: ``lambda a, op: dist.all_reduce(torch.from_numpy(a),
        op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)`` (gloo) or through a device buffer (NCCL)."""
        if fn is None:
            self._allreduce_cb = None
            capi.check(capi.lib().pb200_set_allreduce(self._h, C.cast(None, capi.ALLREDUCE_FN), None),
                       "pb200_set_allreduce")
            return

        def _cb(_user, buf, n, op):
            fn(np.ctypeslib.as_array(buf, shape=(n,)), int(op))
        self._allreduce_cb = capi.ALLREDUCE_FN(_cb)  # keep alive
        capi.check(capi.lib().pb200_set_allreduce(self._h, self._allreduce_cb, None), "pb200_set_allreduce")

    def condition_results(self):
        """Per-condition result packages of the last evaluation (A7)."""
        T = self.tables
        cap = T.n_conditions if T.n_conditions > 0 else 1
        L = capi.lib()
        n = capi.check(L.pb200_get_condition_results(self._h, None, None, None, None), "pb200_get_condition_results")
        n = max(n, 1)
        llh = np.zeros(n)
        sllh = np.zeros((n, T.np_))
        status = np.zeros(n, dtype=np.int32)
        stats = np.zeros((n, PB200_NSTATS), dtype=np.int64)
        n2 = capi.check(L.pb200_get_condition_results(self._h, capi.dptr(llh), capi.dptr(sllh), capi.iptr(status),
                                                      stats.ctypes.data_as(C.POINTER(C.c_int64))),
                        "pb200_get_condition_results")
        return dict(llh=llh[:n2], sllh=sllh[:n2], status=status[:n2], stats=stats[:n2],
                    stat_names=STAT_NAMES)

    def last_launch(self):
        lanes, grid, block, sec = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        capi.lib().pb200_last_launch(self._h, C.byref(lanes), C.byref(grid), C.byref(block), C.byref(sec))
        return dict(lanes=lanes.value, grid=grid.value, block=block.value, seconds=sec.value)

    # ---- hierarchical wrapper (parpe::HierarchicalOptimizationWrapper) -------------
    def hier_num_outer(self) -> int:
        return capi.check(capi.lib().pb200_hier_num_outer(self._h), "pb200_hier_num_outer")

    def hier_evaluate(self, reducedParameters, want_gradient: bool = True):
        theta = np.ascontiguousarray(reducedParameters, dtype=np.float64)
        fval, sec = C.c_double(), C.c_double()
        grad = np.zeros(len(theta)) if want_gradient else None
        full = np.zeros(self.tables.n_opt)
        rc = capi.check(capi.lib().pb200_hier_evaluate(self._h, capi.dptr(theta), len(theta), C.byref(fval),
                                                       capi.dptr(grad), capi.dptr(full), C.byref(sec)),
                        "pb200_hier_evaluate")
        return dict(status=rc, fval=fval.value, grad=grad, full_parameters=full, seconds=sec.value)


def tables_from_hdf5(path, model_name: str) -> ProblemTables:
    """ProblemTables of a parPE input file, assembled by the library (pb200_h5_load_problem): what
    MultiConditionDataProviderHDF5 serves per condition, read once."""
    from .tables import HierTables
    L = capi.lib()
    P = C.c_void_p()
    capi.check(L.pb200_h5_load_problem(str(path).encode(), model_name.encode(), C.byref(P)), "pb200_h5_load_problem")
    try:
        d = L.pb200_h5_problem_desc(P).contents
        Cn, np_, nk, ny, n_opt = d.n_conditions, d.np, d.nk, d.ny, d.n_opt
        arr = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dt).copy()
        t_off = arr(d.t_offsets, Cn + 1, np.int64)
        nt = int(t_off[-1])
        tab = ProblemTables(
            model_name=model_name, np_=np_, nk=nk, ny=ny, n_opt=n_opt, t_offsets=t_off, t=arr(d.t, nt, np.float64),
            y=arr(d.y, nt * ny, np.float64), sigma=arr(d.sigma, nt * ny, np.float64), k=arr(d.k, Cn * nk, np.float64),
            mapping=arr(d.mapping, Cn * np_, np.int32), pscale_sim=arr(d.pscale_sim, Cn * np_, np.int32),
            pscale_opt=arr(d.pscale_opt, n_opt, np.int32),
            overrides=arr(d.overrides, Cn * np_, np.float64) if d.overrides else None, t0=d.t0, atol=d.atol, rtol=d.rtol,
            maxsteps=d.maxsteps, stldet=d.stldet, sensi_errcon=d.sensi_errcon, num_trials=d.num_trials,
            relaxation_factor=d.relaxation_factor, newton_maxsteps=d.newton_maxsteps, newton_only=d.newton_only,
            linsol=d.linsol, steady_sens_newton=d.steady_sens_newton, ism=d.ism)
        hp = L.pb200_h5_problem_hier(P)
        if hp:
            h = hp.contents
            tab.hier = HierTables(
                scaling_idx=arr(h.scaling_idx, h.n_scaling, np.int32), offset_idx=arr(h.offset_idx, h.n_offset, np.int32),
                sigma_idx=arr(h.sigma_idx, h.n_sigma, np.int32),
                scaling_map=arr(h.scaling_map, 3 * h.n_scaling_rows, np.int32),
                offset_map=arr(h.offset_map, 3 * h.n_offset_rows, np.int32),
                sigma_map=arr(h.sigma_map, 3 * h.n_sigma_rows, np.int32))
        lo, up, nom, names = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.c_char_p()
        L.pb200_h5_problem_parameters(P, C.byref(lo), C.byref(up), C.byref(nom), C.byref(names))
        tab.lower_bound = arr(lo, n_opt, np.float64) if lo else None
        tab.upper_bound = arr(up, n_opt, np.float64) if up else None
        tab.nominal_values = arr(nom, n_opt, np.float64) if nom else None
        tab.parameter_names = names.value.decode().split("\n")[:n_opt] if names.value else None
        flag, kp = C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
        roff, ridx = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        if L.pb200_h5_problem_preeq(P, C.byref(flag), C.byref(kp), C.byref(roff), C.byref(ridx)):
            tab.has_preeq = arr(flag, Cn, np.int32)
            tab.k_preeq = arr(kp, Cn * nk, np.float64).reshape(Cn, nk)
            off = arr(roff, Cn + 1, np.int64)
            flat = arr(ridx, max(int(off[-1]), 1), np.int32)
            tab.reinit_idx = [flat[off[c]:off[c + 1]].tolist() for c in range(Cn)]
        return tab
    finally:
        L.pb200_h5_problem_free(P)
