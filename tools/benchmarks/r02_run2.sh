#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_team.py tests/test_gpu_round2.py tests/test_golden_fixtures.py tests/test_cpp_adapter.py -m gpu -x -q > gpurun_out/r02b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_gputests.log
tail -3 gpurun_out/r02b_gputests.log
timeout 300 python tools/benchmarks/team_bench.py chen 148 8 0 2 > gpurun_out/r02b_chen.log 2>&1; tail -2 gpurun_out/r02b_chen.log
timeout 900 python tools/benchmarks/team_bench.py froehlich 148 8 0 1 > gpurun_out/r02b_froehlich.log 2>&1; tail -2 gpurun_out/r02b_froehlich.log
timeout 900 python tools/benchmarks/team_bench.py froehlich 37 8 1 1 > gpurun_out/r02b_froehlich_dense.log 2>&1; tail -2 gpurun_out/r02b_froehlich_dense.log
