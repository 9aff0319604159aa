timeout 600 python -m pytest tests/test_gpu_plugins.py -m gpu -x -q -k "simple_rp or lsh" 2>&1 | tail -4
