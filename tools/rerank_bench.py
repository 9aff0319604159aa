"""Stage 3 (candidate re-rank, csrc/rerank.cu) in isolation: achieved HBM GB/s of `rerank_kernel` at the
BASELINE shapes -- C2 (Q=4096 x 10 candidates x 512-d, euclidean) and C4 (1M x 4096-d histograms, 50 candidates
per query, histogram intersection).  Candidates are random rows (a gather), fixed-pitch layout as the pipeline
produces it.  Usage: python tools/rerank_bench.py [c2|c4|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from smqtk_indexing_b200 import _lib, device  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {"c2": (2_000_000, 512, 4096, 10, "euclidean"), "c4": (1_000_000, 4096, 1024, 50, "hik"),
         "c4e": (1_000_000, 4096, 1024, 50, "euclidean"), "c4c": (1_000_000, 4096, 1024, 50, "cosine")}
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, (N, D, Q, pitch, metric) in cases.items():
    if which not in ("all", name) and not (which == "c4" and name.startswith("c4")):
        continue
    x = torch.rand((N, D), generator=g, device="cuda")
    q = torch.rand((Q, D), generator=g, device="cuda")
    cand = torch.randint(0, N, (Q * pitch,), generator=g, device="cuda", dtype=torch.int64)
    off = torch.arange(Q + 1, device="cuda", dtype=torch.int64) * pitch
    for fixed in (True, False):
        device.rerank(x, q, cand, off, metric, pitch=pitch if fixed else 0)
        torch.cuda.synchronize()
        _lib.profile_fetch()
        _lib.profile_enable(True)
        for _ in range(10):
            flush.zero_()
            device.rerank(x, q, cand, off, metric, pitch=pitch if fixed else 0)
        torch.cuda.synchronize()
        _lib.profile_enable(False)
        ms = sorted(m for n_, m in _lib.profile_fetch() if n_ == "rerank_kernel")
        med = ms[len(ms) // 2]
        nbytes = Q * pitch * D * 4
        print("%s %-9s %s: %d rows x %d B = %.1f MB in %.1f us (median of 10, L2 flushed) = %.0f GB/s" % (
            name, metric, "fixed-pitch" if fixed else "ragged     ", Q * pitch, D * 4, nbytes / 1e6, med * 1e3, nbytes / med / 1e6))
    del x
