timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugins.py -m gpu -x -q -k "l2" 2>&1 | tail -3
for cfg in "12.5e6 256 100" "12.5e6 4096 100" "12.5e6 32 100"; do timeout 120 python tools/l2_scale.py $cfg 2>&1 | tail -1; done
timeout 300 python tools/l2_debug.py 2>&1 | grep -v "^ " | tail -5
