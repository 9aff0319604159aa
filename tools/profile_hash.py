"""ITQ-256 hashing of C2-shaped descriptors (n x 512 fp32 -> 256 bits) through the
tensor-core kernel: the command profiled by ncu for profiles/*_hash_tc*.  Prints
CUDA-event timing when run plainly."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smqtk_indexing_b200 import device as dev

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
D, b = 512, 256
X = torch.rand(n, D, device="cuda")
m = torch.full((D,), 0.5, device="cuda")
R = torch.from_numpy(np.linalg.qr(np.random.RandomState(1).randn(D, D))[0][:, :b].astype(np.float32)).cuda()
img = dev.itq_rotation_image(R)
for _ in range(3):
    dev.itq_hash(X, m, R, variant=2, r_image=img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
iters = 10
e0.record()
for _ in range(iters):
    dev.itq_hash(X, m, R, variant=2, r_image=img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("itq_hash_tc %d x %d -> %d bits: %.3f ms  %.1f Mrows/s  useful %.1f TFLOP/s (issued x3 = %.1f)  X %.1f GB/s" % (
    n, D, b, ms, n / ms / 1e3, 2.0 * n * D * b / ms / 1e9, 6.0 * n * D * b / ms / 1e9, n * D * 4 / ms / 1e6))
