mkdir -p gpurun_out
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r1b_bench_n$n.json 2> gpurun_out/r1b_bench_n$n.err; echo "n=$n rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r1b_bench_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'], d['kernel_ms_per_step'], d['clocks'])
PY
done
