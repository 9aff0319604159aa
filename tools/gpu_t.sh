timeout 900 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_fullsize.py -m gpu -x -q -k "itq or fit or c5 or c1" 2>&1 | tail -4
timeout 600 python tools/fit_bench.py 1e6 256 256 5 2>&1 | tail -7
timeout 600 python tools/fit_bench.py 1e6 256 64 5 2>&1 | head -4
