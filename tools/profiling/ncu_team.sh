#!/bin/bash
# one `ncu --set full` capture of the large-model kernel (Chen shape, 148 conditions, first 3 output times) with
# source attribution, plus the launch list of a bench run; reports land in gpurun_out/
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
MODEL=${1:-chen}; C=${2:-148}; NTP=${3:-3}
timeout 300 python tools/benchmarks/team_bench.py $MODEL $C $NTP 0 1 2>&1 | tail -1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:simulate_team_kernel -c 1 -o gpurun_out/r02_team_${MODEL} -f python tools/benchmarks/team_bench.py $MODEL $C $NTP 0 1 > gpurun_out/r02_ncu_team_${MODEL}.log 2>&1
tail -3 gpurun_out/r02_ncu_team_${MODEL}.log
ls -la gpurun_out/r02_team_${MODEL}.ncu-rep
