import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smqtk_indexing_b200 import device as D, _lib
lib = _lib.load()
for (N, dim, Q, k) in [(20000, 128, 256, 10), (20000, 64, 300, 10), (200000, 128, 64, 100), (20000, 256, 64, 10)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.rand((N, dim), generator=g, device="cuda"); q = torch.rand((Q, dim), generator=g, device="cuda")
    xn, xmax = D.l2_prepare(X)
    ws_bytes = lib.sb_l2_topk_workspace_bytes(dim, Q, k)
    ws = torch.zeros((ws_bytes,), dtype=torch.uint8, device="cuda")
    idx = torch.empty((Q, k), dtype=torch.int64, device="cuda"); dist = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    ov = torch.empty((Q,), dtype=torch.int32, device="cuda")
    _lib.check(lib.sb_l2_topk(X.data_ptr(), N, dim, dim, xn.data_ptr(), xmax.data_ptr(), q.data_ptr(), Q, dim, k,
                              idx.data_ptr(), dist.data_ptr(), ov.data_ptr(), ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    # workspace layout (flat_l2.cu make_l2_plan)
    cols = (Q + 255) // 256 * 256
    cap = 2048
    while cap < 4 * 9 * k: cap *= 2
    a256 = lambda x: (x + 255) // 256 * 256
    o = a256(cols * (dim + 16) * 8)          # img
    off_qn = o; o += a256(cols * 4)
    off_tau = o; o += a256(cols * 4)
    off_margin = o; o += a256(cols * 4)
    off_tq = o; o += a256(cols * 4)
    off_cnt = o; o += a256(cols * 4)
    off_buf = o
    cnt = ws[off_cnt:off_cnt + Q * 4].view(torch.int32).cpu().numpy()
    tau = ws[off_tau:off_tau + Q * 4].view(torch.float32).cpu().numpy()
    margin = ws[off_margin:off_margin + Q * 4].view(torch.float32).cpu().numpy()
    buf = ws[off_buf:off_buf + Q * cap * 8].view(torch.int64).reshape(Q, cap)
    d2_true = ((X.double()[None, :, :] - q.double()[:4, None, :]) ** 2).sum(-1)      # [4, N]
    kth = torch.sort(d2_true, dim=1).values[:, k - 1].cpu().numpy()
    print("N=%d D=%d Q=%d k=%d: overflow=%d  final cnt min/med/max=%d/%d/%d  tau[0:4]=%s true kth d2=%s margin[0]=%.4f" % (
        N, dim, Q, k, int(ov.sum()), cnt.min(), int(np.median(cnt)), cnt.max(), np.round(tau[:4], 3), np.round(kth, 3), margin[0]))
    keys = buf[0, :cnt[0]].cpu().numpy()
    rows = keys & 0xffffffff
    approx = (keys >> 32).astype(np.uint32).view(np.float32)
    true = d2_true[0].cpu().numpy()[rows]
    print("   query 0 survivors: %d, max |approx - true| d2 = %.5f" % (len(rows), np.abs(approx - true).max() if len(rows) else 0))
