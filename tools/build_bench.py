"""Index build (sb_unique_codes) timing: 10M x 256-bit codes -> sorted unique table + CSR.
Usage: python tools/build_bench.py [rows] [W]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from smqtk_indexing_b200 import _lib, device  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
g = torch.Generator(device="cuda").manual_seed(0)
for label, codes in (
        ("random", torch.randint(-2 ** 31, 2 ** 31 - 1, (n, W), dtype=torch.int32, device="cuda", generator=g)),
        ("20% duplicates", None)):
    if codes is None:
        base = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, W), dtype=torch.int32, device="cuda", generator=g)
        pick = torch.randint(0, int(0.8 * n), (n,), device="cuda", generator=g)
        codes = base[pick].contiguous()
    device.unique_codes(codes)
    torch.cuda.synchronize()
    _lib.profile_fetch()
    _lib.profile_enable(True)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        table, rc, off, rows, mx = device.unique_codes(codes)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    _lib.profile_enable(False)
    per = {}
    for name, ms in _lib.profile_fetch():
        per[name] = per.get(name, 0.0) + ms / reps
    print("%s: n=%d W=%d U=%d max/code=%d  %.2f ms per build (%.0f M rows/s)  kernels: %s" % (
        label, n, W, table.shape[0], mx, dt * 1e3, n / dt / 1e6,
        ", ".join("%s %.2f" % kv for kv in per.items())))
