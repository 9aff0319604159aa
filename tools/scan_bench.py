"""Scratch micro-benchmark of the Hamming scan variants (not the contract bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import device as D, _lib

U = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
k = 10
g = torch.Generator(device="cuda").manual_seed(0)
db = torch.randint(-2**31, 2**31 - 1, (U, W), dtype=torch.int32, device="cuda", generator=g)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for Q in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4096", "1"])]:
    q = torch.randint(-2**31, 2**31 - 1, (Q, W), dtype=torch.int32, device="cuda", generator=g)
    ref = None
    for variant in (0, 2, 1):
        for _ in range(2):
            keys = D.hamming_scan_keys(db, q, k, variant=variant)
        if ref is None:
            ref = keys.clone()
        else:
            assert torch.equal(ref, keys), "variant %d differs" % variant
        torch.cuda.synchronize()
        _lib.profile_fetch(); _lib.profile_enable(True)
        iters = 5 if Q > 64 else 20
        for _ in range(iters):
            flush.zero_(); flush.view(torch.int64).sum()
            D.hamming_scan_keys(db, q, k, variant=variant)
        _lib.profile_enable(False)
        ms = sorted(m for n, m in _lib.profile_fetch() if n == "hamming_scan_kernel")
        med = ms[len(ms) // 2]
        print("variant %d: U=%d Q=%d W=%d  scan kernel %.4f ms (min %.4f)  %.1f q/s  alg %.1f GB/s  %.3e pair/s" % (
            variant, U, Q, W, med, ms[0], Q / med * 1e3, U * W * 4 * Q / med / 1e6, U * Q / med * 1e3), flush=True)
