"""Build the CUDA library in-tree: ``parpe_b200/csrc/libparpe_b200.so`` for sm_100a.

``nvcc`` cross-compiles here without a GPU; the built ``.so`` is git-ignored but travels
to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libparpe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-Werror", "cross-execution-space-call",  # a __host__ function called from device code must not build
         "-ccbin", "/usr/bin/g++"] + os.environ.get("PB200_NVCC_EXTRA", "").split()


def _sources():
    return [os.path.join(CSRC, "parpe_b200.cu"), os.path.join(CSRC, "microbench.cu"), os.path.join(CSRC, "h5input.cu"),
            os.path.join(CSRC, "team_kernel_plain.cu"), os.path.join(CSRC, "team_kernel_steady.cu")] + \
        sorted(glob.glob(os.path.join(CSRC, "models", "inst_*.cu"))) + \
        sorted(glob.glob(os.path.join(CSRC, "models", "imported", "inst_*.cu")))  # (codegen/amici_import.py, build time)


_INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src, seen=None):
    """The files `src` includes, transitively (quoted includes, resolved relative to the including file): an
    object is rebuilt only when one of ITS headers changed -- the large-model kernels take minutes to compile."""
    seen = set() if seen is None else seen
    try:
        text = open(src).read()
    except OSError:
        return seen
    for inc in _INCLUDE.findall(text):
        path = os.path.normpath(os.path.join(os.path.dirname(src), inc))
        if path not in seen and os.path.exists(path):
            seen.add(path)
            _deps(path, seen)
    return seen


def _compile(src, obj, log):
    cmd = [NVCC] + ARCH + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout[-4000:]))
    return obj


# AMICI-generated models of the reference tree that the back-end codegen/amici_import.py re-homes as device code at
# build time (never committed: the output contains the reference's generated expressions)
REFERENCE_MODELS = ["/root/reference/deps/AMICI/models/model_steadystate",
                    "/root/reference/deps/AMICI/models/model_jakstat_adjoint"]  # (a spline input: amici::spline_pos)
PYGEN_MODELS = [("steadystate", False), ("boehm", True)]


def import_reference_models(verbose: bool = False):
    from .codegen import amici_import
    for d in REFERENCE_MODELS:
        if not os.path.isdir(d):
            continue
        name = "amici_" + os.path.basename(d)[len("model_"):]
        out = os.path.join(amici_import.OUTDIR, name + ".cuh")
        srcs = glob.glob(os.path.join(d, "*.cpp")) + glob.glob(os.path.join(d, "*.h")) + [amici_import.__file__]
        if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(p) for p in srcs):
            continue
        amici_import.import_model(d, name)
        if verbose:
            print("imported", d, "->", out)
    # models written by AMICI's own python code generator (run in place from the reference tree) from the repo's
    # symbolic descriptions: the python-generated layout, without and with repeating elements (w)
    from .codegen import amici_pygen
    if amici_pygen.available():
        for model, with_w in PYGEN_MODELS:
            name = "pygen_" + model + ("_w" if with_w else "")
            src = os.path.join(amici_pygen.SRC_ROOT, name)
            deps = [amici_pygen.__file__, os.path.join(HERE, "models", model + ".py")]
            if not os.path.exists(os.path.join(src, name + ".h")) or \
                    os.path.getmtime(os.path.join(src, name + ".h")) < max(os.path.getmtime(p) for p in deps):
                amici_pygen.generate(model, with_w)
                if verbose:
                    print("generated", src)
            out = os.path.join(amici_import.OUTDIR, name + ".cuh")
            if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(os.path.join(src, name + ".h")),
                                                                      os.path.getmtime(amici_import.__file__)):
                amici_import.import_model(src, name)
        amici_pygen.write_registry()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    import_reference_models(verbose)
    srcs = _sources()
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        newest = max([os.path.getmtime(s)] + [os.path.getmtime(p) for p in _deps(s)])
        stale = force or not os.path.exists(o) or os.path.getmtime(o) < newest
        jobs.append((s, o, stale))
    todo = [(s, o) for s, o, st in jobs if st]
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            futs = [ex.submit(_compile, s, o, o[:-2] + ".log") for s, o in todo]
            for f in futs:
                f.result()
                if verbose:
                    print("compiled", f.result())
    objs = [o for _, o, _ in jobs]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
