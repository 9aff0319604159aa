timeout 300 python tools/profile_hash.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "itq_hash" 2>&1 | tail -5
