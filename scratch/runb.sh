timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
