#!/bin/bash
# final 1-GPU pass of round 2: scaling records at N = 1, the whole GPU test suite, smoke(), the driver's bench line
# and the reference arm, the Froehlich-shape line, ncu launch list + full capture of the dominant kernel, sanitizer
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
bash tools/benchmarks/scaling.sh 1 2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_final_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_gputests.log; tail -3 gpurun_out/r02_final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_chen.json 2> gpurun_out/r02_bench_chen.err; echo "bench rc=$?"; head -c 500 gpurun_out/r02_bench_chen.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02_bench_chen_reference.json 2> gpurun_out/r02_bench_chen_reference.err; echo "reference rc=$?"; head -c 400 gpurun_out/r02_bench_chen_reference.json; echo
timeout 900 python bench.py --workload boehm --steps 10 --warmup 3 > gpurun_out/r02_bench_boehm.json 2> gpurun_out/r02_bench_boehm.err; echo "boehm rc=$?"
timeout 900 python bench.py --workload froehlich --steps 1 --warmup 3 --no-extra --no-cost-only --no-cpu-baseline > gpurun_out/r02_bench_froehlich.json 2> gpurun_out/r02_bench_froehlich.err; echo "froehlich rc=$?"; head -c 300 gpurun_out/r02_bench_froehlich.json; echo
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_chen.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cost-only --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
# one full capture of the dominant kernel over the whole horizon (DRAM traffic per launch for roofline.traffic)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:simulate_team_kernel -c 1 -o gpurun_out/r02_team_full -f python tools/benchmarks/team_bench.py chen 148 8 0 1 > gpurun_out/r02_ncu_team_full.log 2>&1; echo "full capture rc=$?"; ls -la gpurun_out/r02_team_full.ncu-rep
bash tools/profiling/sanitize.sh > /dev/null 2>&1; tail -40 gpurun_out/r02_sanitizer.log | grep -E "==|ERROR SUMMARY|RACECHECK SUMMARY" | head -30
