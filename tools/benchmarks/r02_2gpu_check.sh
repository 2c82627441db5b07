#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_cpp_adapter.py -m gpu -q -k "nccl or multi_gpu or two_models" > gpurun_out/r02_2gpu_tests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_2gpu_tests_final.log
tail -3 gpurun_out/r02_2gpu_tests_final.log
./tests/cpp/test_multi_gpu 2 2>&1 | tail -1 > gpurun_out/r02_cpp_multi_gpu_final.json; cat gpurun_out/r02_cpp_multi_gpu_final.json
