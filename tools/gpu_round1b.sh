mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches')}, d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['kernel_ms_per_step'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ham_ -s 26 -c 26 --csv --log-file gpurun_out/r1b_scan_tc_launches.csv python tools/scan_tc_bench.py 10e6 4096 10 > gpurun_out/r1b_ncu_tc.log 2>&1; echo "ncu rc=$?"
