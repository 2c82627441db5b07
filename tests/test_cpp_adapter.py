"""The header-only C++ adapter (parpe_b200/include/parpe_b200) compiles against (a) the standalone
interface mirror and (b) -- where /root/reference exists -- the reference's own
parpecommon/functions.h, and on a GPU drives the CUDA path with the reference's call patterns."""
import json
import os
import subprocess

import numpy as np
import pytest

from parpe_b200 import capi
from tests.util import assert_grad_close, assert_llh_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapter")
REF = "/root/reference"


def _compile(extra, out, link=True, src=SRC):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", out, src] + extra
    if link:
        cmd += [capi.LIB_PATH, "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH)]
    subprocess.check_call(cmd)


def test_adapter_compiles_standalone():
    _compile([], EXE)
    assert os.path.exists(EXE)


def test_adapter_compiles_against_reference_headers(tmp_path):
    """Drop-in check: B200SummedGradientFunction derives from the reference's own
    parpe::SummedGradientFunction<int> when built with PB200_USE_PARPE_HEADERS."""
    if not os.path.isdir(os.path.join(REF, "include", "parpecommon")):
        pytest.skip("/root/reference not present")
    inc = ["-DPB200_USE_PARPE_HEADERS", "-I" + os.path.join(REF, "include"),
           "-I" + os.path.join(REF, "deps", "AMICI", "ThirdParty", "gsl"), "-c"]
    # parpecommon/logging.h includes the generated parpeConfig.h: provide an empty one
    cfg = tmp_path / "parpecommon"
    cfg.mkdir()
    (cfg / "parpeConfig.h").write_text("#pragma once\n")
    _compile(inc + ["-I" + str(tmp_path)], str(tmp_path / "adapter.o"), link=False)


@pytest.mark.gpu
def test_adapter_runs_and_matches_python_path():
    from parpe_b200.problem import B200SummedGradientFunction
    from parpe_b200.tables import ProblemTables
    if not os.path.exists(EXE):
        _compile([], EXE)
    out = subprocess.run([EXE], stdout=subprocess.PIPE, check=True, text=True).stdout
    doc = json.loads(out)
    assert doc["status"] == [0, 0, 0, 0] and doc["numParameters"] == 5 and doc["throws_on_size_mismatch"]
    # the same tables through the Python mirror
    C, nt = 4, 20
    t = np.tile(10.0 ** (-5.0 + 6.0 * np.arange(nt) / (nt - 1)), C)
    y = np.array([[0.4 + 0.01 * ((i + 7 * c) % 5) for i in range(nt * 3)] for c in range(C)]).reshape(-1)
    k = np.array([[0.1, 0.4, 0.7, 1.0 + 0.1 * c] for c in range(C)])
    tab = ProblemTables("steadystate", 5, 4, 3, 5, np.arange(C + 1) * nt, t, y, np.full(y.size, 0.1), k,
                        np.tile(np.arange(5), (C, 1)), np.full((C, 5), 2), np.full(5, 2))
    theta = np.array([0.0, np.log10(0.5), np.log10(0.4), np.log10(2.0), -1.0])
    gpu = B200SummedGradientFunction(tab)
    st, f, g, _ = gpu.evaluate(theta)
    # same library; the time grid is computed by libm here and numpy there (1-ulp differences)
    assert_llh_close(doc["fval"], f)
    assert_grad_close(doc["grad"], g)
    assert_llh_close(doc["fval_costonly"], f)
    st, f1, g1, _ = gpu.evaluate(theta, 2)
    assert_llh_close(doc["fval_single"], f1)
    assert_grad_close(doc["grad_single"], g1)
    st, fb, gb, _ = gpu.evaluate(theta, [3, 1])
    assert_llh_close(doc["fval_batch"], fb)
    assert_grad_close(doc["grad_batch"], gb)
    assert doc["n_outputs"] == nt * 3
    from oracle import binding as ob
    if ob.available():
        ro = ob.OracleProblem(tab).evaluate(theta)
        assert_llh_close(doc["fval"], ro["fval"])
        assert_grad_close(doc["grad"], ro["grad"])


SRC_MULTI = os.path.join(ROOT, "tests", "cpp", "test_multi_gpu.cpp")
EXE_MULTI = os.path.join(ROOT, "tests", "cpp", "test_multi_gpu")


def test_multi_gpu_adapter_compiles():
    """parpe::B200MultiGpuSummedGradientFunction (one handle per device + the library's NCCL all-reduce)
    builds without Python or torch; without GPUs the program reports that it was skipped."""
    _compile([], EXE_MULTI, src=SRC_MULTI)
    out = subprocess.run([EXE_MULTI, "2"], stdout=subprocess.PIPE, check=True, text=True).stdout
    doc = json.loads(out.strip().splitlines()[-1])
    assert "skipped" in doc or doc["ok"]


@pytest.mark.gpu
def test_multi_gpu_adapter_matches_single_gpu():
    """Two GPUs driven from C++ through the reference's adapter: same fval / gradient as one GPU."""
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not os.path.exists(EXE_MULTI):
        _compile([], EXE_MULTI, src=SRC_MULTI)
    r = subprocess.run([EXE_MULTI, "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    doc = json.loads(r.stdout.strip().splitlines()[-1])  # (NCCL may print its version banner first)
    assert doc["ok"] and doc["devices"] == 2 and doc["rel_err"] < 1e-11 and doc["grad_rel_err"] < 1e-11


EXE_MB = os.path.join(ROOT, "tests", "cpp", "test_minibatch")


def test_reference_minibatch_optimizer_builds_against_the_adapter():
    """The reference's UNMODIFIED MinibatchOptimizer (minibatchOptimization.{h,cpp}, logging.cpp, misc.cpp,
    functions.cpp, parpeException.cpp compiled in place from /root/reference, HDF5 type stubs only) links with
    parpe::B200SummedGradientFunction derived from the reference's own SummedGradientFunction<int>."""
    if not os.path.isdir(os.path.join(REF, "src", "parpeoptimization")):
        pytest.skip("/root/reference not present")
    subprocess.check_call([os.path.join(ROOT, "tests", "cpp", "build_minibatch.sh")])
    out = subprocess.run([EXE_MB], stdout=subprocess.PIPE, check=True, text=True).stdout
    doc = json.loads(out.strip().splitlines()[-1])
    assert "skipped" in doc or doc["ok"]


@pytest.mark.gpu
def test_reference_minibatch_optimizer_drives_the_cuda_path():
    """Three epochs of the reference's own mini-batch optimizer (Adam updater, batches of 4 of 12 conditions,
    std::shuffle per epoch) on the steady-state example through the adapter on the GPU: the cost decreases.
    The binary is built where /root/reference exists (build() / the CPU test) and travels with the snapshot."""
    if not os.path.exists(EXE_MB):
        pytest.skip("tests/cpp/test_minibatch not built (needs /root/reference at build time)")
    r = subprocess.run([EXE_MB], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    # (the reference's Logger writes its per-batch debug lines to stdout; the result is the last line)
    doc = json.loads(r.stdout.strip().splitlines()[-1])
    assert doc["ok"] and doc["epochs"] == 3 and doc["cost_final"] < doc["cost_start"]
    assert "Cost:" in r.stdout and "Batch:" in r.stdout  # the reference optimizer's own log of the batches
