python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1_bench_n8.json 2> gpurun_out/r1_bench_n8.err; echo "bench n8 rc=$?"
tail -c 1800 gpurun_out/r1_bench_n8.json; tail -3 gpurun_out/r1_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r1_bench_n4.json 2> gpurun_out/r1_bench_n4.err; echo "bench n4 rc=$?"
tail -c 600 gpurun_out/r1_bench_n4.json
