#!/bin/bash
# Compiles tests/cpp/test_minibatch.cpp against the reference's OWN optimiser sources, in place
# (needs /root/reference; outputs only tests/cpp/test_minibatch, which is git-ignored but travels to
# the GPU box with the snapshot).  HDF5 headers: type stubs in tests/cpp/stubs.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
LIBDIR="$HERE/../../parpe_b200/csrc"
g++ -std=c++17 -O1 -pthread -DPB200_USE_PARPE_HEADERS -Dgsl_CONFIG_CONTRACT_VIOLATION_THROWS \
    -I"$HERE/stubs" -I"$REF/include" -I"$REF/deps/AMICI/ThirdParty/gsl" \
    -o "$HERE/test_minibatch" "$HERE/test_minibatch.cpp" \
    "$REF/src/parpeoptimization/minibatchOptimization.cpp" "$REF/src/parpecommon/logging.cpp" \
    "$REF/src/parpecommon/misc.cpp" "$REF/src/parpecommon/functions.cpp" "$REF/src/parpecommon/parpeException.cpp" \
    "$LIBDIR/libparpe_b200.so" -Wl,-rpath,"$LIBDIR" -ldl
