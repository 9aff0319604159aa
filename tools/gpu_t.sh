timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tensor_core_scan or dispatch" 2>&1 | grep -E "^E|passed|failed|Error" | head -20
timeout 200 python tools/scan_tc_bench.py 10e6 4096 10 2>&1 | grep "^tc\|identical"
timeout 200 python tools/scan_tc_bench.py 10e6 4096 10 4 2>&1 | grep "^tc\|identical"
