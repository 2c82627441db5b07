#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_team.py tests/test_gpu_round2.py tests/test_golden_fixtures.py -m gpu -x -q > gpurun_out/r02d_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_gputests.log
tail -3 gpurun_out/r02d_gputests.log
timeout 300 python tools/benchmarks/team_bench.py chen 148 8 0 2 > gpurun_out/r02d_chen.log 2>&1; tail -1 gpurun_out/r02d_chen.log
timeout 300 python tools/benchmarks/team_bench.py chen 37 8 0 1 > gpurun_out/r02d_chen37.log 2>&1; tail -1 gpurun_out/r02d_chen37.log
timeout 900 python tools/benchmarks/team_bench.py froehlich 148 8 0 1 > gpurun_out/r02d_froehlich.log 2>&1; tail -1 gpurun_out/r02d_froehlich.log
