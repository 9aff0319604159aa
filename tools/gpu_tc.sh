timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?"; tail -20 gpurun_out/tc_debug.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "itq_hash" > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/tc_pytest.log
