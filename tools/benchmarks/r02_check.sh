#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gputests.log; tail -3 gpurun_out/r02e_gputests.log
timeout 300 python tools/benchmarks/team_bench.py chen 148 8 0 2 > gpurun_out/r02e_chen.log 2>&1; tail -1 gpurun_out/r02e_chen.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_chen.json 2> gpurun_out/r02e_bench_chen.err; echo "bench rc=$?"; head -c 300 gpurun_out/r02e_bench_chen.json; echo
