"""Problem tables: the in-memory form of what parPE's data provider serves.

``ProblemTables`` carries exactly the arrays of ``pb200_problem_desc``
(``include/parpe_b200.h``), i.e. the content of the reference's HDF5 input
(``doc/hdf5.md:1-172``) that ``AmiciSummedGradientFunction::evaluate`` consumes
through ``MultiConditionDataProviderHDF5``
(``src/parpeamici/multiConditionDataProvider.cpp:84-169,302-339``).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np

PB200_ABI_VERSION = 2
PB200_NSTATS = 20
STAT_NAMES = ["nst", "nfe", "nsetups", "nje", "nni", "nniS", "nfSe", "netf", "netfS", "ncfn", "ncfnS", "nor",
              "trials", "qsum"]
SCALE_NONE, SCALE_LN, SCALE_LOG10 = 0, 1, 2


class ProblemDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("model_name", C.c_char_p),
        ("n_conditions", C.c_int32),
        ("n_opt", C.c_int32),
        ("np", C.c_int32), ("nk", C.c_int32), ("ny", C.c_int32),
        ("t_offsets", C.POINTER(C.c_int64)),
        ("t", C.POINTER(C.c_double)),
        ("y", C.POINTER(C.c_double)),
        ("sigma", C.POINTER(C.c_double)),
        ("k", C.POINTER(C.c_double)),
        ("mapping", C.POINTER(C.c_int32)),
        ("pscale_sim", C.POINTER(C.c_int32)),
        ("pscale_opt", C.POINTER(C.c_int32)),
        ("overrides", C.POINTER(C.c_double)),
        ("t0", C.c_double),
        ("atol", C.c_double), ("rtol", C.c_double),
        ("maxsteps", C.c_int32),
        ("stldet", C.c_int32),
        ("sensi_errcon", C.c_int32),
        ("num_trials", C.c_int32),
        ("relaxation_factor", C.c_double),
        ("lanes_per_condition", C.c_int32),
        ("newton_maxsteps", C.c_int32),
        ("linsol", C.c_int32),
        ("newton_only", C.c_int32),
        ("steady_sens_newton", C.c_int32),
        ("ism", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


class HierDesc(C.Structure):
    _fields_ = [
        ("n_scaling", C.c_int32), ("n_offset", C.c_int32), ("n_sigma", C.c_int32),
        ("scaling_idx", C.POINTER(C.c_int32)),
        ("offset_idx", C.POINTER(C.c_int32)),
        ("sigma_idx", C.POINTER(C.c_int32)),
        ("n_scaling_rows", C.c_int32), ("n_offset_rows", C.c_int32), ("n_sigma_rows", C.c_int32),
        ("scaling_map", C.POINTER(C.c_int32)),
        ("offset_map", C.POINTER(C.c_int32)),
        ("sigma_map", C.POINTER(C.c_int32)),
        ("reserved", C.c_int32 * 8),
    ]


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


@dataclasses.dataclass
class HierTables:
    """(inner parameter, condition, observable) row tables, the content of
    ``/{scaling,offset,sigma}ParameterIndices`` and ``...MapToObservables``."""
    scaling_idx: np.ndarray
    offset_idx: np.ndarray
    sigma_idx: np.ndarray
    scaling_map: np.ndarray  # [rows,3]
    offset_map: np.ndarray
    sigma_map: np.ndarray

    def __post_init__(self):
        for f in ("scaling_idx", "offset_idx", "sigma_idx"):
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=np.int32).reshape(-1))
        for f in ("scaling_map", "offset_map", "sigma_map"):
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=np.int32).reshape(-1, 3))

    def to_desc(self) -> HierDesc:
        d = HierDesc()
        d.n_scaling, d.n_offset, d.n_sigma = len(self.scaling_idx), len(self.offset_idx), len(self.sigma_idx)
        d.scaling_idx, d.offset_idx, d.sigma_idx = (_ptr(self.scaling_idx, C.c_int32), _ptr(self.offset_idx, C.c_int32),
                                                    _ptr(self.sigma_idx, C.c_int32))
        d.n_scaling_rows, d.n_offset_rows, d.n_sigma_rows = (len(self.scaling_map), len(self.offset_map),
                                                             len(self.sigma_map))
        d.scaling_map, d.offset_map, d.sigma_map = (_ptr(self.scaling_map, C.c_int32), _ptr(self.offset_map, C.c_int32),
                                                    _ptr(self.sigma_map, C.c_int32))
        return d

    @property
    def inner_idx(self):
        return np.sort(np.concatenate([self.scaling_idx, self.offset_idx, self.sigma_idx]))


@dataclasses.dataclass
class ProblemTables:
    model_name: str
    np_: int
    nk: int
    ny: int
    n_opt: int
    t_offsets: np.ndarray      # int64 [C+1]
    t: np.ndarray              # float64 [sum nt]
    y: np.ndarray              # float64 [sum nt * ny], NaN = missing
    sigma: np.ndarray          # float64 [sum nt * ny], NaN = model sigma
    k: np.ndarray              # float64 [C, nk]
    mapping: np.ndarray        # int32 [C, np]
    pscale_sim: np.ndarray     # int32 [C, np]
    pscale_opt: np.ndarray     # int32 [n_opt]
    overrides: Optional[np.ndarray] = None  # float64 [C, np]
    t0: float = 0.0
    atol: float = 1e-12
    rtol: float = 1e-8
    maxsteps: int = 10000
    stldet: int = 1
    sensi_errcon: int = 1
    num_trials: int = 6
    relaxation_factor: float = 100.0
    lanes_per_condition: int = 0
    newton_maxsteps: int = 0   # steady states: Newton iterations before / after the simulation (AMICI default 0)
    newton_only: int = 0       # 1 = steadyStateComputationMode newtonOnly: no simulation if Newton's method fails
    steady_sens_newton: int = 0  # 1 = steadyStateSensitivityMode newtonOnly with a simulation to steady state
    ism: int = 0               # 0 / 2 staggered corrector, 1 simultaneous
    linsol: int = 0            # large models: 0 automatic, 1 dense, 2 sparse (initial value of PB200_OPT_LINSOL)
    hier: Optional[HierTables] = None
    # pre-equilibration (pb200_set_preequilibration): per simulation condition a flag, the constants of
    # the pre-equilibration condition and the states re-initialised afterwards
    has_preeq: Optional[np.ndarray] = None   # int32 [C]
    k_preeq: Optional[np.ndarray] = None     # float64 [C, nk]
    reinit_idx: Optional[list] = None        # C lists of state indices
    # /parameters/{lowerBound, upperBound, nominalValues, parameterNames} (the optimisation problem's side of the file)
    lower_bound: Optional[np.ndarray] = None
    upper_bound: Optional[np.ndarray] = None
    nominal_values: Optional[np.ndarray] = None
    parameter_names: Optional[list] = None

    def preeq_args(self):
        """(has_preeq, k_preeq, reinit_offsets, reinit_idx) arrays for pb200_set_preequilibration, or None."""
        if self.has_preeq is None:
            return None
        Cn = self.n_conditions
        flag = np.ascontiguousarray(self.has_preeq, dtype=np.int32).reshape(Cn)
        kp = np.ascontiguousarray(self.k_preeq if self.k_preeq is not None else np.zeros((Cn, self.nk)),
                                  dtype=np.float64).reshape(Cn, self.nk)
        lists = self.reinit_idx if self.reinit_idx is not None else [[] for _ in range(Cn)]
        off = np.zeros(Cn + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(l) for l in lists])
        flat = np.ascontiguousarray([i for l in lists for i in l] or [0], dtype=np.int32)
        return flag, kp, off, flat

    def __post_init__(self):
        self.t_offsets = np.ascontiguousarray(self.t_offsets, dtype=np.int64)
        self.t = np.ascontiguousarray(self.t, dtype=np.float64)
        self.y = np.ascontiguousarray(self.y, dtype=np.float64).reshape(-1)
        self.sigma = np.ascontiguousarray(self.sigma, dtype=np.float64).reshape(-1)
        Cn = len(self.t_offsets) - 1
        self.k = np.ascontiguousarray(self.k, dtype=np.float64).reshape(Cn, self.nk)
        self.mapping = np.ascontiguousarray(self.mapping, dtype=np.int32).reshape(Cn, self.np_)
        self.pscale_sim = np.ascontiguousarray(self.pscale_sim, dtype=np.int32).reshape(Cn, self.np_)
        self.pscale_opt = np.ascontiguousarray(self.pscale_opt, dtype=np.int32).reshape(self.n_opt)
        if self.overrides is not None:
            self.overrides = np.ascontiguousarray(self.overrides, dtype=np.float64).reshape(Cn, self.np_)
        assert self.y.size == self.t.size * self.ny == self.sigma.size
        assert self.t_offsets[-1] == self.t.size

    @property
    def n_conditions(self) -> int:
        return len(self.t_offsets) - 1

    def nt(self, c: int) -> int:
        return int(self.t_offsets[c + 1] - self.t_offsets[c])

    def to_desc(self) -> ProblemDesc:
        d = ProblemDesc()
        d.abi_version = PB200_ABI_VERSION
        self._name_bytes = self.model_name.encode()
        d.model_name = self._name_bytes
        d.n_conditions, d.n_opt = self.n_conditions, self.n_opt
        d.np, d.nk, d.ny = self.np_, self.nk, self.ny
        d.t_offsets = _ptr(self.t_offsets, C.c_int64)
        d.t, d.y, d.sigma, d.k = (_ptr(self.t, C.c_double), _ptr(self.y, C.c_double), _ptr(self.sigma, C.c_double),
                                  _ptr(self.k, C.c_double))
        d.mapping, d.pscale_sim, d.pscale_opt = (_ptr(self.mapping, C.c_int32), _ptr(self.pscale_sim, C.c_int32),
                                                 _ptr(self.pscale_opt, C.c_int32))
        d.overrides = _ptr(self.overrides, C.c_double) if self.overrides is not None else None
        d.t0, d.atol, d.rtol = self.t0, self.atol, self.rtol
        d.maxsteps, d.stldet, d.sensi_errcon = self.maxsteps, self.stldet, self.sensi_errcon
        d.num_trials, d.relaxation_factor = self.num_trials, self.relaxation_factor
        d.lanes_per_condition = self.lanes_per_condition
        d.newton_maxsteps = self.newton_maxsteps
        d.newton_only, d.linsol, d.steady_sens_newton = self.newton_only, self.linsol, self.steady_sens_newton
        d.ism = self.ism
        return d

    # per-condition views in the reference's layout (row-major obs + it*ny)
    def y_of(self, c):
        return self.y[self.t_offsets[c] * self.ny:self.t_offsets[c + 1] * self.ny].reshape(-1, self.ny)

    def sigma_of(self, c):
        return self.sigma[self.t_offsets[c] * self.ny:self.t_offsets[c + 1] * self.ny].reshape(-1, self.ny)
