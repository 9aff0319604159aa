timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugins.py -m gpu -x -q -k "l2" 2>&1 | tail -3
timeout 120 python tools/l2_chunks.py 12.5e6 4096 100 2>&1 | grep filter
timeout 300 python tools/l2_bench.py 12.5e6 128 100 2>&1 | tail -6 | cut -c1-130
