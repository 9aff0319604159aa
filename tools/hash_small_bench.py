"""ITQ hash of SMALL query batches: FFMA kernel vs tensor-core kernel (512-d -> 256 bits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smqtk_indexing_b200 import _lib, device as dev
D, b = 512, 256
R = torch.from_numpy(np.linalg.qr(np.random.RandomState(0).randn(D, D))[0][:, :b].astype(np.float32)).cuda()
mean = torch.full((D,), 0.5, device="cuda")
img = dev.itq_rotation_image(R)
for n in (256, 512, 1024, 2048, 4096, 8192):
    X = torch.rand((n, D), device="cuda")
    out = {}
    for name, variant in (("ffma", 1), ("tc", 2)):
        for _ in range(3):
            dev.itq_hash(X, mean, R, variant=variant, r_image=img)
        torch.cuda.synchronize()
        _lib.profile_fetch(); _lib.profile_enable(True)
        for _ in range(10):
            dev.itq_hash(X, mean, R, variant=variant, r_image=img)
        torch.cuda.synchronize(); _lib.profile_enable(False)
        ms = sorted(m for _, m in _lib.profile_fetch())
        out[name] = ms[len(ms) // 2] * 1e3
    print("n=%5d  ffma %.1f us   tc %.1f us" % (n, out["ffma"], out["tc"]))
