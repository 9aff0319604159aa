for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r1_bench_n$n.json 2> gpurun_out/r1_bench_n$n.err; echo "bench n$n rc=$?"
done
python - <<'PY'
import json
for n in (2,4,8):
    d=json.loads(open('gpurun_out/r1_bench_n%d.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['clocks'])
PY
