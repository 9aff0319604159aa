mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1b_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['cpu_baseline'])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1b_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1b_ncu_launches.log 2>&1; echo "ncu1 rc=$?"
