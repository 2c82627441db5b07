#!/bin/bash
# first GPU pass of round 2: parity tests, then short bench runs of the large-model path (sparse vs dense)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
tail -5 gpurun_out/r02_gputests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_chen_sparse.json 2> gpurun_out/r02_chen_sparse.err; echo "chen sparse rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --linsol 1 > gpurun_out/r02_chen_dense.json 2> gpurun_out/r02_chen_dense.err; echo "chen dense rc=$?"
timeout 600 python bench.py --workload froehlich --conditions 148 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_froehlich_sparse.json 2> gpurun_out/r02_froehlich_sparse.err; echo "froehlich rc=$?"
timeout 300 python bench.py --workload boehm --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_boehm.json 2> gpurun_out/r02_boehm.err; echo "boehm rc=$?"
timeout 300 python bench.py --workload boehm --scaling strong --conditions 1250 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_boehm_1250.json 2> gpurun_out/r02_boehm_1250.err; echo "boehm1250 rc=$?"
head -c 600 gpurun_out/r02_chen_sparse.json; echo; head -c 600 gpurun_out/r02_chen_dense.json; echo; head -c 600 gpurun_out/r02_froehlich_sparse.json; echo
