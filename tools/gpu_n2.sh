python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r1_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1_bench_n2.json 2> gpurun_out/r1_bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r1_bench_n2.json; tail -5 gpurun_out/r1_bench_n2.err
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 30 > gpurun_out/r1_bench_ref.json 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/r1_bench_ref.json
