"""Scratch: tensor-core Hamming scan vs the XOR/POPC scan (U x 256-bit codes, Q queries, k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import _lib, device as dev

U = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
W = int(sys.argv[4]) if len(sys.argv) > 4 else 8
g = torch.Generator(device="cuda"); g.manual_seed(0)
db = torch.randint(-2**31, 2**31 - 1, (U, W), dtype=torch.int32, device="cuda", generator=g)
q = torch.randint(-2**31, 2**31 - 1, (Q, W), dtype=torch.int32, device="cuda", generator=g)
WITH_POPC = os.environ.get("WITH_POPC", "1") != "0"
cases = [("fp4", lambda: dev.hamming_scan_keys_tc(db, q, k, fmt="fp4")[0]), ("fp8", lambda: dev.hamming_scan_keys_tc(db, q, k, fmt="fp8")[0])]
if WITH_POPC:
    cases.append(("popc", lambda: dev.hamming_scan_keys(db, q, k, variant=1)))
ONLY = os.environ.get("ONLY")
if ONLY:
    cases = [c for c in cases if c[0] == ONLY]
ref = None
for name, fn in cases:
    out = fn(); torch.cuda.synchronize()
    _lib.profile_fetch(); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(False)
    agg = {}
    for nm, ms in _lib.profile_fetch():
        agg[nm] = agg.get(nm, 0.0) + ms / reps
    ms = e0.elapsed_time(e1) / reps
    print("%-5s U=%d Q=%d k=%d W=%d: %.3f ms/batch  %.1f q/s | %s" % (name, U, Q, k, W, ms, Q / ms * 1e3,
          {a: round(b, 3) for a, b in sorted(agg.items(), key=lambda kv: -kv[1])}))
    if ref is None:
        ref = out
    else:
        print("   identical keys:", bool(torch.equal(ref, out)))
