mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['kernel_ms_per_step'], d['clocks'], d['single_query_scan'])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1b_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1b_ncu_launches.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ham_filter_tc -s 7 -c 1 -o gpurun_out/r1_ham_tc_full -f python tools/scan_tc_bench.py 10e6 4096 10 > gpurun_out/ham_ncu.log 2>&1; echo "ncu2 rc=$?"
