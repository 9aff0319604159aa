"""Scratch: flat L2 kNN (BASELINE config 3 per-GPU shard: 12.5M x 128-d, k=100)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import device as D, _lib

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.rand((N, dim), generator=g, device="cuda")
prep = D.l2_prepare(X)
torch.cuda.synchronize()
for Q in (4096, 256, 32, 1):
    q = torch.rand((Q, dim), generator=g, device="cuda")
    for _ in range(2):
        idx, dist = D.l2_topk(X, q, k, prepared=prep)
    torch.cuda.synchronize()
    _lib.profile_fetch(); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    iters = 3
    e0.record()
    for _ in range(iters):
        idx, dist = D.l2_topk(X, q, k, prepared=prep)
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(False)
    ms = e0.elapsed_time(e1) / iters
    agg = {}
    for n_, m_ in _lib.profile_fetch():
        agg[n_] = agg.get(n_, 0.0) + m_ / iters
    print("N=%d D=%d k=%d Q=%d: %.3f ms/batch  %.1f q/s  useful %.1f TFLOP/s | %s" % (
        N, dim, k, Q, ms, Q / ms * 1e3, 2.0 * N * dim * Q / ms / 1e9,
        {a: round(b, 3) for a, b in sorted(agg.items(), key=lambda kv: -kv[1])}), flush=True)
    if Q == 4096:
        # exact check of three queries against a float64 torch brute force
        for qi in (0, 1777, 4095):
            best = None
            for s in range(0, N, 2_000_000):
                d = (X[s:s + 2_000_000].double() - q[qi].double()).square_().sum(1).sqrt_()
                dd, ii = torch.topk(d, min(k, d.numel()), largest=False)
                cand = torch.stack([dd, (ii + s).double()], 1)
                best = cand if best is None else torch.cat([best, cand])
            order = torch.argsort(best[:, 0], stable=True)[:k]
            ref_d, ref_i = best[order, 0], best[order, 1].long()
            assert torch.allclose(dist[qi], ref_d, rtol=1e-9, atol=0), (dist[qi] - ref_d).abs().max()
            assert torch.equal(idx[qi], ref_i), qi
        print("   verified 3 queries against float64 brute force (identical rows, rtol 1e-9)")
