# Round-2 closing single-GPU evidence in one gpurun call (default scan = packed FP4, three queries per column):
# GPU test suite, smoke, bench line + reference arm, ncu launch list of the bench command, per-launch DRAM bytes of
# one scan batch, ncu --set full of the FP4 filter's largest chunk, the C4 bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 30 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 --no-graph > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 --no-graph > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
ONLY=fp4 python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/plain2.log 2>&1 &&
ONLY=fp4 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ham -s 32 -c 32 \
    --csv --log-file gpurun_out/r2_scan_tc_launches.csv python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/ncu2.log 2>&1; echo "scan launches rc=$?"
ONLY=fp4 ncu --set full --clock-control none --import-source on -k regex:ham_filter_fp4 --launch-skip 7 --launch-count 1 -f -o gpurun_out/r2_ham_fp4x3 \
    python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/ncu3.log 2>&1; echo "scan full rc=$?"
python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"
cat gpurun_out/plain2.log | tail -2
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
