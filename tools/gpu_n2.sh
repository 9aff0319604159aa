mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1b_bench_n2.json 2> gpurun_out/r1b_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1b_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'], d['kernel_ms_per_step'], d['clocks'])
PY
tail -3 gpurun_out/r1b_bench_n2.err
