timeout 600 python -m pytest tests/test_gpu_plugins.py -m gpu -x -q -k "streaming or itq" 2>&1 | tail -3
timeout 600 python tools/fit_bench.py 5e7 256 256 10 > gpurun_out/r1_c5_fit_build_b256.log 2>&1; tail -7 gpurun_out/r1_c5_fit_build_b256.log
timeout 600 python tools/fit_bench.py 5e7 256 64 50 > gpurun_out/r1_c5_fit_build_b64.log 2>&1; tail -7 gpurun_out/r1_c5_fit_build_b64.log
