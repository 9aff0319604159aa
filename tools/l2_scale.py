import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import device as D
N = int(float(sys.argv[1])); Q = int(sys.argv[2]); k = int(sys.argv[3])
X = torch.rand((N, 128), device="cuda"); prep = D.l2_prepare(X); q = torch.rand((Q, 128), device="cuda")
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter()
    idx, dist = D.l2_topk(X, q, k, prepared=prep)
    torch.cuda.synchronize()
    ov = getattr(D, "LAST_L2_OVERFLOW", None)
    print("N=%d Q=%d k=%d call %d: %.2f ms  overflowed queries: %s" % (N, Q, k, i, 1e3 * (time.perf_counter() - t0), ov), flush=True)
