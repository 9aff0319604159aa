timeout 300 python tools/scan_bench.py 1e7 1 > gpurun_out/few_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hamming_scan_few -s 8 -c 1 -o gpurun_out/r1_scan_few_full python tools/scan_bench.py 1e7 1 > gpurun_out/few_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import torch
x = torch.randint(-2**31, 2**31-1, (10_000_000, 8), dtype=torch.int32, device="cuda")
for _ in range(3): x.sum()
e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
fl=torch.empty(512<<20,dtype=torch.uint8,device="cuda")
ts=[]
for _ in range(10):
    fl.zero_(); e0.record(); x.sum(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("torch sum over 320MB: median %.4f ms -> %.1f GB/s" % (ts[5], 320/ts[5]))
y=torch.empty_like(x)
ts=[]
for _ in range(10):
    fl.zero_(); e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("torch copy 320MB: median %.4f ms -> %.1f GB/s (r+w)" % (ts[5], 640/ts[5]))
PY
