#!/bin/bash
# the north-star configuration at 8 GPUs, one quick point (a single evaluation takes ~36 s):
# Froehlich shape (nx = 1200, nplist = 100), 1184 conditions = 148 per GPU (a 10^3-minibatch-sized index list)
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload froehlich --scaling strong --conditions $((148 * N)) --quick > gpurun_out/r02_froehlich_${N}gpu.json 2> gpurun_out/r02_froehlich_${N}gpu.err
echo "rc=$?"; head -c 600 gpurun_out/r02_froehlich_${N}gpu.json; echo
