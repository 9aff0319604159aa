timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugins.py -m gpu -q -x -k "l2 or flat" 2>&1 | tail -3
timeout 300 python tools/l2_bench.py 2>&1 | tail -8 | tee gpurun_out/r1_flat_l2_bench.log
