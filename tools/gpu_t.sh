timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugins.py -m gpu -q -x 2>&1 | tail -2
