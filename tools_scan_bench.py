"""Scratch micro-benchmark of the Hamming scan (not the contract bench)."""
import sys, time
import torch
from smqtk_indexing_b200 import device as D

U = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
k = 10
g = torch.Generator(device="cuda").manual_seed(0)
db = torch.randint(-2**31, 2**31 - 1, (U, W), dtype=torch.int32, device="cuda", generator=g)
q = torch.randint(-2**31, 2**31 - 1, (Q, W), dtype=torch.int32, device="cuda", generator=g)
for variant in (0, 1):
    for _ in range(2):
        keys = D.hamming_scan_keys(db, q, k, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    iters = 3
    e0.record()
    for _ in range(iters):
        keys = D.hamming_scan_keys(db, q, k, variant=variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("variant %d: U=%d Q=%d W=%d  %.3f ms  %.1f q/s  alg %.1f GB/s  %.3e pair/s" % (
        variant, U, Q, W, ms, Q / ms * 1e3, U * W * 4 * Q / ms / 1e6, U * Q / ms * 1e3))
