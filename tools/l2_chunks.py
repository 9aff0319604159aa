import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import device as D, _lib
N, Q, k = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
X = torch.rand((N, 128), device="cuda"); prep = D.l2_prepare(X); q = torch.rand((Q, 128), device="cuda")
for _ in range(2): D.l2_topk(X, q, k, prepared=prep)
torch.cuda.synchronize(); _lib.profile_fetch(); _lib.profile_enable(True)
D.l2_topk(X, q, k, prepared=prep)
_lib.profile_enable(False)
for n_, ms in _lib.profile_fetch(): print("%-26s %8.3f ms" % (n_, ms))
