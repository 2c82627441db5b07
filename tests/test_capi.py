"""The C-ABI library loads and exports every symbol include/parpe_b200.h declares; without a CUDA
device the compute entry points fail loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from parpe_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "parpe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTED_SYMBOLS) == names


def test_abi_version_and_model_registry():
    L = capi.lib()
    assert L.pb200_abi_version() == 2
    m = capi.models()
    assert m["boehm"] == dict(nx=8, np=9, nk=2, ny=3)
    assert m["steadystate"] == dict(nx=3, np=5, nk=4, ny=3)
    assert m["steadystate_scaled"] == dict(nx=3, np=13, nk=1, ny=6)
    assert m["fujita"]["nx"] == 9


def test_partition_bounds_rule():
    """rank r of R owns [floor(r n / R), floor((r+1) n / R)): contiguous, in order, exhaustive."""
    for n in (0, 1, 7, 16, 10000, 16385):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = capi.partition_bounds(n, r, world)
                assert lo == (r * n) // world and hi == ((r + 1) * n) // world
                assert lo == prev
                prev = hi
            assert prev == n


def test_no_cpu_fallback():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from parpe_b200.problem import B200SummedGradientFunction
    tab = synth.make_problem("steadystate", 2)
    with pytest.raises(capi.B200LibraryError) as e:
        B200SummedGradientFunction(tab)
    assert "no CUDA device" in str(e.value)


def test_argument_validation_before_device_use():
    L = capi.lib()
    tab = synth.make_problem("steadystate", 2)
    d = tab.to_desc()
    h = C.c_void_p()
    d.model_name = b"no_such_model"
    assert L.pb200_create(C.byref(d), 0, C.byref(h)) == -2  # PB200_ERR_MODEL
    d = tab.to_desc()
    d.abi_version = 99
    assert L.pb200_create(C.byref(d), 0, C.byref(h)) == -1  # PB200_ERR_ARG
    d = tab.to_desc()
    d.np = 4
    assert L.pb200_create(C.byref(d), 0, C.byref(h)) == -2
