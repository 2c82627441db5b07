#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/benchmarks/team_bench.py chen 148 8 0 2 0 2>&1 | tail -1 | tee -a gpurun_out/r02f_ab.log
