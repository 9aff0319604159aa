# BASELINE configs[2..4] through bench.py --config (one GPU): JSON lines under gpurun_out/
for c in c4 c3 c5; do
  timeout 400 python bench.py --config $c --steps 3 --warmup 3 --fit-iters 10 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  echo "$c rc=$?"; tail -2 gpurun_out/r2_bench_$c.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$c.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","parity_checked","fit_s","build_s")}); print(d.get("e2e")); print(d.get("rerank")); print(d.get("kernel_ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
