mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1b_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
PY
