#!/bin/bash
# two GPUs of one box: in-library NCCL (Python threads and the C++ multi-GPU adapter), bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_2gpu_gpus.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_cpp_adapter.py tests/test_gpu_parity.py -m gpu -q -k "nccl or multi_gpu or two_models or partition" > gpurun_out/r02_2gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_2gpu_tests.log
tail -4 gpurun_out/r02_2gpu_tests.log
./tests/cpp/test_multi_gpu 2 > gpurun_out/r02_cpp_multi_gpu.json 2>&1; echo "cpp multi gpu rc=$?"; cat gpurun_out/r02_cpp_multi_gpu.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT ./tests/cpp/test_multi_gpu 2 2>&1 | grep -iE "NCCL version|NVLS|P2P|via" | head -12 > gpurun_out/r02_cpp_multi_gpu_nccl.txt
for W in chen boehm; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload $W --no-extra --no-cpu-baseline > gpurun_out/r02_2gpu_${W}.json 2> gpurun_out/r02_2gpu_${W}.err; echo "$W N=2 rc=$?"; tail -c 400 gpurun_out/r02_2gpu_${W}.json | head -c 400; echo
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --workload boehm > gpurun_out/r02_2gpu_reference_boehm.json 2> gpurun_out/r02_2gpu_reference_boehm.err; echo "ref rc=$?"; head -c 300 gpurun_out/r02_2gpu_reference_boehm.json; echo
