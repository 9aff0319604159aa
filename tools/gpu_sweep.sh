for U in 1.25e6 2.5e6 1e7; do for w in 1 2 3 4 6 8 12 16; do echo -n "U=$U waves=$w: "; SB_SCAN_WAVES=$w timeout 120 python tools/scan_bench.py $U 4096 2>&1 | grep "variant 0" | sed 's/.*scan kernel//'; done; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b2.json 2> gpurun_out/b2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['kernel_ms_per_step'])
PY
tail -3 gpurun_out/b2.err
