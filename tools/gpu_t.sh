timeout 600 python -m pytest tests/test_gpu_plugins.py -m gpu -q -k "gram or itq_fit or fit" 2>&1 | grep -E "^E|passed|failed" | head -20
timeout 600 python tools/fit_bench.py 50e6 256 256 10 2>&1 | tee gpurun_out/c5_fit_b256_tcgram.log
timeout 600 python tools/fit_bench.py 50e6 256 64 50 2>&1 | tee gpurun_out/c5_fit_b64_tcgram.log
