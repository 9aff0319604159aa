timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; echo "bench rc=$?"; tail -3 gpurun_out/b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['single_query_scan'], d['kernel_ms_per_step'])
PY
