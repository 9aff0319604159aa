# BASELINE configs[1] under torchrun on 8 (and 4) GPUs of one box: default partition (queries), then the
# row-partitioned scan for comparison.  Run through: gpurun --gpus 8 -- bash tools/gpu_n8.sh
mkdir -p gpurun_out
run() {  # n partition tag
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2961$1 \
    bench.py --gpus $1 --steps 20 --warmup 3 --scan-partition $2 > gpurun_out/r2_bench_n$1_$2.json 2> gpurun_out/r2_bench_n$1_$2.err
  echo "n=$1 $2 rc=$?"
  grep -v "^frame\|^$\|OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_n$1_$2.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$1_$2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','eager','parity_checked','n_gpus')}, d['e2e']['value'])
    print(d['config']['parallelism']); print({k:round(v,4) for k,v in d['kernel_ms_per_step'].items()}); print(d['index'], d['clocks'])
except Exception as e:
    print('no line', e)
PY
}
run 8 auto
run 8 rows
run 4 auto
