"""Scratch: ItqFunctor.fit_matrix + build throughput (reduced C5: n x 256-d -> b bits)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import _lib, engine
from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
b = int(sys.argv[3]) if len(sys.argv) > 3 else 256
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
X = torch.empty((n, D), device="cuda")
for s0 in range(0, n, 4_000_000):           # generate in slabs (keeps the RNG temporaries small)
    X[s0:s0 + 4_000_000].uniform_()
torch.cuda.synchronize()
f = ItqFunctor(bit_length=b, itq_iterations=iters, random_seed=0)
_lib.profile_fetch(); _lib.profile_enable(True)
t0 = time.perf_counter()
f.fit_matrix(X, want_codes=False)
torch.cuda.synchronize()
t_fit = time.perf_counter() - t0
_lib.profile_enable(False)
agg = {}
for name, ms in _lib.profile_fetch():
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
print("fit %d x %d -> %d bits, %d iterations: %.3f s (%.1f k rows/s)" % (n, D, b, iters, t_fit, n / t_fit / 1e3))
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("   %-28s %4d launches %10.3f ms" % (k, c, ms))
flops_iter = 4.0 * n * b * b
print("   per ITQ iteration: %.1f GFLOP (FP64)" % (flops_iter / 1e9))
t0 = time.perf_counter()
m = engine.DeviceLshIndex()
m.set_rows(X, f.get_hash_packed(X))
torch.cuda.synchronize()
t_b = time.perf_counter() - t0
print("build (hash + sort-unique + CSR) %d rows: %.3f s (%.2f M rows/s), %d unique codes" % (n, t_b, n / t_b / 1e6, m.num_codes))
