#!/bin/bash
# compute-sanitizer memcheck + racecheck of both kernels (and the sparse / fused paths of the large-model
# kernel); the logs are kept under gpurun_out/ and copied to profiles/ (VERDICT round 1, item 12).
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, '.')
from parpe_b200 import capi, synth
from parpe_b200.problem import B200SummedGradientFunction
model, C, ntp, linsol = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
tab = synth.make_problem(model, C, seed=0, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp])
theta = synth.random_theta(tab, seed=1, width=0.3)
gpu = B200SummedGradientFunction(tab)
gpu.set_option(capi.OPT_LINSOL, linsol)
print(model, "linsol", linsol, gpu.evaluate(theta)[:2], gpu.evaluate(theta, want_gradient=False)[:2])
PY
out=gpurun_out/r02_sanitizer.log
: > $out
run() { echo "== $1: $2 $3 conditions, $4 output times, linsol $5" | tee -a $out; timeout ${6:-900} compute-sanitizer --tool $1 --print-limit 5 python /tmp/san_case.py $2 $3 $4 $5 2>&1 | tail -8 | tee -a $out; }
run memcheck boehm 40 6 0
run racecheck boehm 40 6 0
run memcheck net32 3 2 2
run racecheck net32 3 2 2
run memcheck net32 3 2 1
run racecheck net32 3 2 1
run memcheck chen 2 1 0 1500
run racecheck chen 1 1 0 1500
