import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smqtk_indexing_b200 import device as D
N, dim, k, Q = 4_000_000, 128, 100, 256
X = torch.rand((N, dim), device="cuda"); prep = D.l2_prepare(X)
q = torch.rand((Q, dim), device="cuda")
for _ in range(3):
    D.l2_topk(X, q, k, prepared=prep)
torch.cuda.synchronize()
print("ok")
