# BASELINE configs[1] under torchrun on N GPUs of one box (default partition).  gpurun --gpus N -- bash tools/gpu_scale.sh N [tests]
N=$1
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n${N}_auto.json 2> gpurun_out/r2_bench_n${N}_auto.err
echo "n=$N rc=$?"
grep -v "^frame\|^$\|OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_n${N}_auto.err | tail -3 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}_auto.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','eager','parity_checked','n_gpus')}, d['e2e']['value'])
    print(d['config']['parallelism']); print({k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}); print(d['index'], d['clocks'])
except Exception as e:
    print('no line', e)
PY
if [ "$2" = "tests" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4; fi
