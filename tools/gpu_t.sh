timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/l2_bench.py 12.5e6 128 100 2>&1 | tail -6 | cut -c1-250
