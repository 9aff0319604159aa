mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1_ncu_launches.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hamming_scan_kernel -s 3 -c 1 -o gpurun_out/r1_scan_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu2 rc=$?"
