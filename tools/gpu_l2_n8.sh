mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tools/l2_sharded_bench.py 12.5e6 100 2>gpurun_out/l2s.err | tee gpurun_out/r1_flat_l2_sharded_n8.log
tail -3 gpurun_out/l2s.err
