timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/fit_bench.py 1e7 256 256 1 2>&1 | tail -3
