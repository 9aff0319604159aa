# Round-2 single-GPU evidence in one gpurun call: bench line, ncu launch list of the bench command, ncu --set full of
# the scan kernel (FP8) and of the flat-L2 filter, DRAM traffic of one scan batch.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 30 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"
# launch list of the bench command (kernel by kernel: graph replay would hide the launches from the host-side count)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 --no-graph > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fit-iters 5 --no-graph > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
# one scan batch: time + DRAM bytes of every launch (second batch of the script: -s skips the warm-up batch)
ONLY=fp8 python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/plain2.log 2>&1 &&
ONLY=fp8 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ham_ -s 30 -c 30 \
    --csv --log-file gpurun_out/r2_scan_tc_launches.csv python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/ncu2.log 2>&1; echo "scan launches rc=$?"
# the scan kernel's largest chunk, full set
ONLY=fp8 ncu --set full --clock-control none --import-source on -k regex:ham_filter_tc -s 17 -c 1 -o gpurun_out/r2_ham_tc \
    python tools/scan_tc_bench.py 10e6 4096 10 8 > gpurun_out/ncu3.log 2>&1; echo "scan full rc=$?"
# flat L2 filter (current kernel), one 12.5M x 128 shard, Q = 4096: the largest chunk
python tools/l2_bench.py 12.5e6 128 100 > gpurun_out/r2_flat_l2_bench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l2_filter_tma -s 4 -c 1 -o gpurun_out/r2_l2_filter \
    python tools/l2_bench.py 12.5e6 128 100 > gpurun_out/ncu4.log 2>&1; echo "l2 full rc=$?"
tail -4 gpurun_out/r2_flat_l2_bench.log
python tools/build_bench.py > gpurun_out/r2_build_bench.log 2>&1; cat gpurun_out/r2_build_bench.log
python tools/rerank_bench.py all > gpurun_out/r2_rerank_bench.log 2>&1; cat gpurun_out/r2_rerank_bench.log
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
