timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size" > gpurun_out/s10_tests.log 2>&1; tail -15 gpurun_out/s10_tests.log
