"""Scratch: error statistics + timing of the tensor-core hash kernel vs the FP64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
from smqtk_indexing_b200 import device as dev, _lib
import np_oracle as O

for (n, D, b) in [(256, 16, 32), (1000, 512, 256), (4096, 512, 256)]:
    rng = np.random.RandomState(0)
    x = rng.rand(n, D).astype(np.float32)
    mean = x.mean(0).astype(np.float32)
    rot = np.linalg.qr(rng.randn(max(D, b), max(D, b)))[0][:D, :b]
    X = torch.from_numpy(x).cuda(); m = torch.from_numpy(mean).cuda(); R = torch.from_numpy(rot.astype(np.float32)).cuda()
    zr = O.itq_project(x.astype(np.float64), mean.astype(np.float64), rot)
    for v in (1, 2):
        c, z = dev.itq_hash(X, m, R, want_z=True, variant=v)
        torch.cuda.synchronize()
        z = z.cpu().numpy()
        err = np.abs(z - zr)
        bits = O.pack_codes(zr >= 0, c.shape[1])
        flips = int(np.bitwise_count(dev.codes_to_host(c) ^ bits).sum())
        print("n=%d D=%d b=%d variant=%d max|dz|=%.3e mean|dz|=%.3e flips=%d/%d" % (n, D, b, v, err.max(), err.mean(), flips, n * b), flush=True)

# throughput on the C2 shape
n, D, b = 2_000_000, 512, 256
X = torch.rand(n, D, device="cuda"); m = torch.full((D,), 0.5, device="cuda")
R = torch.from_numpy(np.linalg.qr(np.random.RandomState(1).randn(D, D))[0][:, :b].astype(np.float32)).cuda()
img = dev.itq_rotation_image(R)
for v in (1, 2):
    for _ in range(2):
        dev.itq_hash(X, m, R, variant=v, r_image=img)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        dev.itq_hash(X, m, R, variant=v, r_image=img)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("variant %d: %d x %d -> %d bits: %.3f ms, %.1f Mrows/s, useful %.1f TFLOP/s, X read %.1f GB/s" % (
        v, n, D, b, ms, n / ms / 1e3, 2.0 * n * D * b / ms / 1e9, n * D * 4 / ms / 1e6), flush=True)
