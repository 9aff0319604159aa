timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plugins.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python tools/scan_tc_bench.py 10e6 4096 10 2>&1 | grep "^tc"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
