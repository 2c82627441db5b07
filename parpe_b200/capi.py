"""ctypes binding of the C ABI (``include/parpe_b200.h``) -- the library is
``parpe_b200/csrc/libparpe_b200.so`` (CUDA, sm_100a).  There is no CPU fallback: if the
library or a CUDA device is missing, the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .tables import PB200_NSTATS, HierDesc, ProblemDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libparpe_b200.so")

EXPORTED_SYMBOLS = [
    "pb200_abi_version", "pb200_device_count", "pb200_model_count", "pb200_model_name", "pb200_model_info",
    "pb200_last_error", "pb200_create", "pb200_destroy", "pb200_set_partition", "pb200_partition_bounds",
    "pb200_evaluate", "pb200_evaluate_device", "pb200_last_launch", "pb200_get_condition_results",
    "pb200_get_model_outputs", "pb200_measure_fp64_peak", "pb200_set_allreduce", "pb200_hier_attach", "pb200_hier_num_outer", "pb200_hier_evaluate",
    "pb200_set_option", "pb200_nccl_unique_id", "pb200_nccl_init", "pb200_last_transfer_bytes", "pb200_linsol_info",
    "pb200_h5_open", "pb200_h5_close", "pb200_h5_exists", "pb200_h5_dataset_info", "pb200_h5_read_double",
    "pb200_h5_attribute_double", "pb200_h5_list", "pb200_h5_read_raw", "pb200_h5_load_problem", "pb200_h5_problem_desc",
    "pb200_h5_problem_hier", "pb200_h5_problem_free", "pb200_create_from_hdf5", "pb200_set_preequilibration",
    "pb200_h5_problem_preeq", "pb200_h5_problem_parameters", "pb200_h5_read_strings", "pb200_result_log_create", "pb200_result_log_destroy", "pb200_result_log_append",
    "pb200_result_log_append_records", "pb200_result_log_num_records", "pb200_result_log_write",
    "pb200_write_simulation_results",
]
# pb200_set_option keys (include/parpe_b200.h)
OPT_REORDER, OPT_ORDER_REFRESH, OPT_KERNEL_VARIANT, OPT_CTAS_PER_SM, OPT_SYNC_PERIOD, OPT_STAGE_SYNC = 1, 2, 3, 4, 5, 6
OPT_PARTIAL_SUMS, OPT_LINSOL, OPT_FUSE_OFF = 7, 8, 9
LINSOL_AUTO, LINSOL_DENSE, LINSOL_SPARSE = 0, 1, 2

_lib = None
# void (*pb200_allreduce_fn)(void* user, double* buf, int n, int op)
ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int)
REDUCE_SUM, REDUCE_MAX = 0, 1


class B200LibraryError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            "CUDA extension missing: %s (run `python -m parpe_b200.build`); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.pb200_abi_version.restype = C.c_int
    L.pb200_device_count.restype = C.c_int
    L.pb200_model_count.restype = C.c_int
    L.pb200_model_name.restype = C.c_char_p
    L.pb200_model_name.argtypes = [C.c_int]
    L.pb200_model_info.argtypes = [C.c_char_p, ip, ip, ip, ip]
    L.pb200_last_error.restype = C.c_char_p
    L.pb200_create.argtypes = [C.POINTER(ProblemDesc), C.c_int, C.POINTER(C.c_void_p)]
    L.pb200_destroy.argtypes = [C.c_void_p]
    L.pb200_destroy.restype = None
    L.pb200_set_partition.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pb200_partition_bounds.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    L.pb200_partition_bounds.restype = None
    L.pb200_evaluate.argtypes = [C.c_void_p, dp, C.c_int, ip, C.c_int, dp, dp, dp]
    L.pb200_evaluate_device.argtypes = [C.c_void_p, dp, C.c_int, ip, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.pb200_last_launch.argtypes = [C.c_void_p, ip, ip, ip, dp]
    L.pb200_get_condition_results.argtypes = [C.c_void_p, dp, dp, ip, C.POINTER(C.c_int64)]
    L.pb200_get_model_outputs.argtypes = [C.c_void_p, dp, C.c_int, dp, dp]
    L.pb200_measure_fp64_peak.argtypes = [C.c_int, dp]
    L.pb200_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p]
    L.pb200_hier_attach.argtypes = [C.c_void_p, C.POINTER(HierDesc)]
    L.pb200_hier_num_outer.argtypes = [C.c_void_p]
    L.pb200_hier_evaluate.argtypes = [C.c_void_p, dp, C.c_int, dp, dp, dp, dp]
    L.pb200_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pb200_nccl_unique_id.argtypes = [C.c_char_p]
    L.pb200_nccl_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    L.pb200_last_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.pb200_linsol_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    L.pb200_h5_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.pb200_h5_close.argtypes = [C.c_void_p]
    L.pb200_h5_close.restype = None
    L.pb200_h5_exists.argtypes = [C.c_void_p, C.c_char_p]
    L.pb200_h5_dataset_info.argtypes = [C.c_void_p, C.c_char_p, ip, C.POINTER(C.c_int64), ip, ip]
    L.pb200_h5_read_double.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_int64]
    L.pb200_h5_read_double.restype = C.c_int64
    L.pb200_h5_read_raw.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int64]
    L.pb200_h5_read_raw.restype = C.c_int64
    L.pb200_h5_attribute_double.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, dp, C.c_int]
    L.pb200_h5_list.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    L.pb200_h5_load_problem.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.pb200_h5_problem_desc.argtypes = [C.c_void_p]
    L.pb200_h5_problem_desc.restype = C.POINTER(ProblemDesc)
    L.pb200_h5_problem_hier.argtypes = [C.c_void_p]
    L.pb200_h5_problem_hier.restype = C.POINTER(HierDesc)
    L.pb200_h5_problem_preeq.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(dp),
                                         C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32))]
    i32p = C.POINTER(C.c_int32)
    L.pb200_result_log_create.argtypes = [C.POINTER(C.c_void_p)]
    L.pb200_result_log_destroy.argtypes = [C.c_void_p]
    L.pb200_result_log_destroy.restype = None
    L.pb200_result_log_append.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    L.pb200_result_log_append_records.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, dp, dp,
                                                  dp, dp, i32p, C.POINTER(C.c_int64)]
    L.pb200_result_log_num_records.argtypes = [C.c_void_p, C.c_char_p]
    L.pb200_result_log_write.argtypes = [C.c_void_p, C.c_char_p]
    L.pb200_write_simulation_results.argtypes = [C.c_void_p, dp, C.c_int, C.c_char_p, C.c_char_p]
    L.pb200_h5_problem_parameters.argtypes = [C.c_void_p, C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), C.POINTER(C.c_char_p)]
    L.pb200_h5_read_strings.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    L.pb200_h5_problem_free.argtypes = [C.c_void_p]
    L.pb200_h5_problem_free.restype = None
    L.pb200_create_from_hdf5.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.pb200_set_preequilibration.argtypes = [C.c_void_p, C.POINTER(C.c_int32), dp, C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int32)]
    _lib = L
    return L


def last_error() -> str:
    return lib().pb200_last_error().decode()


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise B200LibraryError("%s failed (%d): %s" % (what, rc, last_error()))
    return rc


def device_count() -> int:
    return lib().pb200_device_count()


def models():
    L = lib()
    out = {}
    for i in range(L.pb200_model_count()):
        name = L.pb200_model_name(i)
        nx, np_, nk, ny = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.pb200_model_info(name, C.byref(nx), C.byref(np_), C.byref(nk), C.byref(ny))
        out[name.decode()] = dict(nx=nx.value, np=np_.value, nk=nk.value, ny=ny.value)
    return out


def measure_fp64_peak(device: int = 0) -> float:
    v = C.c_double()
    check(lib().pb200_measure_fp64_peak(device, C.byref(v)), "pb200_measure_fp64_peak")
    return v.value


def partition_bounds(n: int, rank: int, world: int):
    lo, hi = C.c_int(), C.c_int()
    lib().pb200_partition_bounds(n, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class ResultLog:
    """Per-simulation result logging (pb200_result_log_*): what parPE's logSimulation writes per simulation job
    (multiConditionProblem.cpp:254-325; doc/hdf5.md:190-226), collected in memory and written once."""

    def __init__(self):
        self._h = C.c_void_p()
        check(lib().pb200_result_log_create(C.byref(self._h)), "pb200_result_log_create")

    def append(self, fun, group="/multistarts/0", optimization_idx=0, iteration_idx=0):
        """the per-condition results of ``fun``'s last evaluate (a B200SummedGradientFunction)"""
        return check(lib().pb200_result_log_append(self._h, fun._h, group.encode(), optimization_idx, iteration_idx),
                     "pb200_result_log_append")

    def append_records(self, group, condition_idx, llh, status, sllh=None, parameters=None, seconds=None, stats=None,
                       optimization_idx=0, iteration_idx=0, np_=0):
        cidx = np.ascontiguousarray(condition_idx, dtype=np.int32)
        llh = np.ascontiguousarray(llh, dtype=np.float64)
        status = np.ascontiguousarray(status, dtype=np.int32)
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        sllh, parameters, seconds = f(sllh), f(parameters), f(seconds)
        if stats is not None:
            stats = np.ascontiguousarray(stats, dtype=np.int64)
        if sllh is not None:
            np_ = sllh.reshape(len(llh), -1).shape[1]
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        return check(lib().pb200_result_log_append_records(
            self._h, group.encode(), len(llh), np_, i32(cidx), optimization_idx, iteration_idx, dptr(llh), dptr(sllh),
            dptr(parameters), dptr(seconds), i32(status),
            stats.ctypes.data_as(C.POINTER(C.c_int64)) if stats is not None else None), "pb200_result_log_append_records")

    def num_records(self, group="/multistarts/0"):
        return lib().pb200_result_log_num_records(self._h, group.encode())

    def write(self, path):
        check(lib().pb200_result_log_write(self._h, str(path).encode()), "pb200_result_log_write")

    def close(self):
        if self._h:
            lib().pb200_result_log_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class H5File:
    """parPE input files through the library's built-in HDF5 reader (pb200_h5_*; no HDF5 library needed)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        check(lib().pb200_h5_open(str(path).encode(), C.byref(self._h)), "pb200_h5_open")

    def close(self):
        if self._h:
            lib().pb200_h5_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exists(self, path):
        return bool(lib().pb200_h5_exists(self._h, path.encode()))

    def info(self, path):
        rank, cls, size = C.c_int(), C.c_int(), C.c_int()
        dims = (C.c_int64 * 8)()
        check(lib().pb200_h5_dataset_info(self._h, path.encode(), C.byref(rank), dims, C.byref(cls), C.byref(size)),
              "pb200_h5_dataset_info")
        return dict(shape=tuple(int(dims[i]) for i in range(rank.value)), type_class=cls.value, type_size=size.value)

    def read_strings(self, path):
        """dataset of fixed- or variable-length strings -> list of str"""
        buf = C.create_string_buffer(1 << 20)
        n = check(lib().pb200_h5_read_strings(self._h, path.encode(), buf, len(buf)), "pb200_h5_read_strings")
        return buf.value.decode().split("\n")[:n]

    def read(self, path):
        n = check(lib().pb200_h5_read_double(self._h, path.encode(), None, 0), "pb200_h5_read_double")
        out = np.zeros(max(int(n), 1))
        check(lib().pb200_h5_read_double(self._h, path.encode(), dptr(out), len(out)), "pb200_h5_read_double")
        return out[:n].reshape(self.info(path)["shape"])

    def attribute(self, obj, name):
        out = np.zeros(64)
        n = check(lib().pb200_h5_attribute_double(self._h, obj.encode(), name.encode(), dptr(out), len(out)),
                  "pb200_h5_attribute_double")
        return out[:n]

    def list(self, group):
        buf = C.create_string_buffer(1 << 20)
        check(lib().pb200_h5_list(self._h, group.encode(), buf, len(buf)), "pb200_h5_list")
        return [s for s in buf.value.decode().split("\n") if s]
