"""
Device plumbing: torch owns HBM allocations and streams, the kernels are called
through the C ABI with raw pointers (``tensor.data_ptr()``) on torch's current
stream.  Nothing here computes; every function is a thin typed wrapper around
one ``sb_*`` entry point.

Storage conventions (torch has no unsigned 32/64-bit arithmetic types worth
using, so raw bits are carried in signed tensors of the same width):
  codes  uint32[rows, W]  -> torch.int32
  keys   uint64[...]      -> torch.int64
"""
import os
import threading
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("smqtk_indexing_b200 needs a CUDA device (B200): the LSH "
                           "hot path has no CPU implementation.")
    _lib.load()


def device(dev: Optional[object] = None) -> torch.device:
    require_cuda()
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def codes_to_device(words: np.ndarray, dev: Optional[object] = None) -> torch.Tensor:
    """uint32[n, W] numpy -> int32 CUDA tensor (bit pattern preserved)."""
    w = np.ascontiguousarray(words, dtype=np.uint32)
    return torch.from_numpy(w.view(np.int32)).to(device(dev), non_blocking=False)


def codes_to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint32)


# --------------------------------------------------------------------- stage 1
def norm_spec(normalize) -> Tuple[int, float]:
    """Map ``ItqFunctor(normalize=...)`` (an ``ord`` for numpy.linalg.norm over a
    1-D vector, reference itq.py:184) onto the kernel's (kind, p)."""
    if normalize is None:
        return _lib.NORM_NONE, 0.0
    if isinstance(normalize, str):
        raise ValueError("normalize=%r is not a valid vector norm order" % (normalize,))
    o = float(normalize)
    if o == float("inf"):
        return _lib.NORM_INF, 0.0
    if o == 0.0:
        return _lib.NORM_L0, 0.0
    if o > 0.0:
        return _lib.NORM_LP, o
    raise ValueError("normalize=%r: only ord in {None, 0, p > 0, inf} is supported on the device" % (normalize,))


#: below this the FFMA kernel's launch is as fast as the tensor-core one.  (Both cost ~50 us for a few hundred
#: rows -- the persistent tensor-core kernel streams the whole pre-split rotation, 2 MB at 512 x 256, into every
#: CTA it starts; a dedicated few-rows kernel was tried in round 2 and was slower, 139 us at 512 rows.)
TC_MIN_ROWS = 512


def itq_rotation_image(R: torch.Tensor) -> Optional[torch.Tensor]:
    """Pre-split rotation (hi/lo TF32, UMMA core-matrix order) for ``sb_itq_hash_tc``;
    None when the shape is not supported by the tensor-core kernel."""
    require_cuda()
    _chk(R, torch.float32, "R")
    D, b = R.shape
    lib = _lib.load()
    nbytes = lib.sb_itq_rotation_image_bytes(D, b)
    if nbytes == 0:
        return None
    img = torch.empty((nbytes,), dtype=torch.uint8, device=R.device)
    with torch.cuda.device(R.device):
        _lib.check(lib.sb_itq_rotation_image(_ptr(R), D, b, _ptr(img), _stream()))
    return img


def itq_tc_supported(X: torch.Tensor, b: int) -> bool:
    n, D = X.shape
    ldx = X.stride(0) if n > 1 else max(D, X.stride(0))
    return (n >= 1 and D % 16 == 0 and D >= 16 and b % 32 == 0 and 32 <= b <= 256 and ldx % 4 == 0
            and X.data_ptr() % 16 == 0)


def itq_hash(X: torch.Tensor, mean: Optional[torch.Tensor], R: torch.Tensor, normalize=None,
             words: Optional[int] = None, want_z: bool = False, variant: int = 0,
             r_image: Optional[torch.Tensor] = None):
    """codes = pack(((X / norm(X)) - mean) @ R >= 0).  X f32[n, D] (row stride may
    exceed D), mean f32[D] | None, R f32[D, b].  Returns codes int32[n, W]
    (and z f32[n, b] when ``want_z``).

    variant 0 = auto (tensor-core kernel when ``r_image`` is given, the shape is
    aligned and n >= TC_MIN_ROWS, else the FFMA kernel), 1 = FFMA kernel,
    2 = tensor-core kernel (raises when the shape is not supported)."""
    require_cuda()
    if X.dim() != 2 or X.dtype != torch.float32 or not X.is_cuda or X.stride(1) != 1:
        raise ValueError("X must be a 2-D float32 CUDA tensor with unit column stride")
    _chk(R, torch.float32, "R")
    n, D = X.shape
    if R.shape[0] != D:
        raise ValueError("R has %d rows, X has %d columns" % (R.shape[0], D))
    b = R.shape[1]
    if mean is not None:
        _chk(mean, torch.float32, "mean")
        if mean.numel() != D:
            raise ValueError("mean has %d entries, X has %d columns" % (mean.numel(), D))
    from .utils.bits import words_for_bits
    W = words or words_for_bits(b)
    kind, p = norm_spec(normalize)
    codes = torch.empty((n, W), dtype=torch.int32, device=X.device)
    z = torch.empty((n, b), dtype=torch.float32, device=X.device) if want_z else None
    ldx = X.stride(0) if n > 1 else max(D, X.stride(0))
    lib = _lib.load()
    use_tc = variant == 2 or (variant == 0 and r_image is not None and n >= TC_MIN_ROWS and itq_tc_supported(X, b))
    with torch.cuda.device(X.device):
        if use_tc:
            if r_image is None:
                r_image = itq_rotation_image(R)
            if r_image is None or not itq_tc_supported(X, b):
                raise ValueError("tensor-core hashing needs D % 16 == 0, b % 32 == 0, 32 <= b <= 256 and 16-byte "
                                 "aligned rows (D=%d, b=%d)" % (D, b))
            div = None
            if kind != _lib.NORM_NONE:
                div = torch.empty((n,), dtype=torch.float32, device=X.device)
                _lib.check(lib.sb_itq_row_div(_ptr(X), n, D, ldx, kind, p, _ptr(div), _stream()))
            _lib.check(lib.sb_itq_hash_tc(_ptr(X), n, D, ldx, _ptr(mean), _ptr(r_image), b, _ptr(div),
                                          _ptr(codes), W, _ptr(z), _stream()))
        else:
            _lib.check(lib.sb_itq_hash(_ptr(X), n, D, ldx, _ptr(mean), _ptr(R), b, kind, p,
                                       _ptr(codes), W, _ptr(z), 0, _stream()))
    return (codes, z) if want_z else codes


# --------------------------------------------------------------------- index build
def unique_codes(codes: torch.Tensor, rows: Optional[torch.Tensor] = None):
    """``sb_unique_codes``: (table int32[U, W] sorted unique, row_code int64[n_rows] (-1 = row not in
    ``rows``), csr_off int64[U + 1], csr_rows int64[n], max rows per code) of the codes of ``rows``
    (int64 ascending; None = every row).  One host read (U and the max count) sizes the outputs."""
    require_cuda()
    _chk(codes, torch.int32, "codes")
    n_rows, W = codes.shape
    if rows is not None:
        _chk(rows, torch.int64, "rows")
    n = n_rows if rows is None else int(rows.numel())
    lib = _lib.load()
    ws_bytes = lib.sb_unique_codes_workspace_bytes(n, W)
    if ws_bytes == 0:
        raise ValueError("sb_unique_codes does not support n=%d W=%d" % (n, W))
    dev = codes.device
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    table = torch.empty((n, W), dtype=torch.int32, device=dev)
    row_code = torch.empty((n_rows,), dtype=torch.int64, device=dev)
    csr_off = torch.empty((n + 1,), dtype=torch.int64, device=dev)
    csr_rows = torch.empty((n,), dtype=torch.int64, device=dev)
    stats = torch.empty((2,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sb_unique_codes(_ptr(codes), n_rows, W, _ptr(rows), n, _ptr(table), _ptr(row_code), _ptr(csr_off),
                                       _ptr(csr_rows), _ptr(stats), _ptr(ws), ws_bytes, _stream()))
    U, max_count = (int(v) for v in stats.cpu().tolist())
    # the slices keep the full-capacity storage alive; the table is compacted when it is much smaller
    table = table[:U] if 2 * U > n else table[:U].clone()
    return table, row_code, csr_off[:U + 1], csr_rows, max_count


# --------------------------------------------------------------------- stage 2
#: the tensor-core scan takes over from this many queries / table rows (below, the XOR/POPC scan
#: streams the table once per handful of queries and is bound by HBM, which is the better regime)
TC_SCAN_MIN_QUERIES = 64
TC_SCAN_MIN_ROWS = 1 << 16
#: ... and this many (query, row) pairs: the tensor-core scan runs ~5.7e12 pairs/s plus ~0.25 ms of
#: per-batch bookkeeping (8 chunks x 4 launches), the XOR/POPC scan ~0.9e12 pairs/s with none
TC_SCAN_MIN_PAIRS = 1 << 28


def _tc_scan_pays(U: int, W: int, Q: int, k: int) -> bool:
    return (Q >= TC_SCAN_MIN_QUERIES and U >= TC_SCAN_MIN_ROWS and Q * U >= TC_SCAN_MIN_PAIRS
            and hamming_scan_tc_supported(U, W, Q, k))
SCAN_VARIANT_TC = 3
#: per-device int32[1] accumulators of the tensor-core scan's overflow flags (device memory: the
#: dispatcher never reads a flag on the host -- the predicated XOR/POPC scan redoes an overflowed batch)
_overflow_acc = {}


def tc_scan_overflows(dev: Optional[object] = None) -> int:
    """How many tensor-core scan batches overflowed a candidate buffer and were redone by the
    predicated XOR/POPC scan on ``dev`` (diagnostics / tests; synchronises)."""
    d = device(dev)
    acc = _overflow_acc.get(d.index if d.index is not None else torch.cuda.current_device())
    return 0 if acc is None else int(acc.item())


def hamming_scan_tc_supported(U: int, W: int, Q: int, k: int) -> bool:
    return bool(_lib.load().sb_hamming_scan_tc_supported(U, W, Q, k))


#: operand format of the tensor-core scan: "fp4" (packed E2M1, kind::mxf4: half the tensor-pipe time per MMA;
#: several queries share one FP32 accumulator column so the tensor-memory read of the epilogue does not grow --
#: the default) or "fp8" (E4M3, kind::f8f6f4, FP16 accumulators read packed); same keys.  Shapes
#: neither kernel takes (k > 256, codes wider than 256 bits) run the XOR/POPC scan
TC_SCAN_FORMAT = os.environ.get("SB_TC_SCAN_FORMAT", "fp4")


def hamming_scan_keys_tc(db: torch.Tensor, q: torch.Tensor, k: int, idx_base: int = 0, fmt: Optional[str] = None):
    """Tensor-core scan (``sb_hamming_scan_tc4`` / ``sb_hamming_scan_tc``) -> (keys int64[Q, k], overflow int32[1])."""
    U, W = db.shape
    Q = q.shape[0]
    lib = _lib.load()
    fmt = fmt or TC_SCAN_FORMAT
    fp4 = fmt == "fp4" and bool(lib.sb_hamming_scan_tc4_supported(U, W, Q, k))
    ws_fn, scan_fn = ((lib.sb_hamming_scan_tc4_workspace_bytes, lib.sb_hamming_scan_tc4) if fp4 else
                      (lib.sb_hamming_scan_tc_workspace_bytes, lib.sb_hamming_scan_tc))
    ws_bytes = ws_fn(U, W, Q, k)
    ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=db.device)
    keys = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    flag = torch.empty((1,), dtype=torch.int32, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(scan_fn(_ptr(db), U, W, _ptr(q), Q, k, idx_base, _ptr(keys), _ptr(flag), _ptr(ws), ws_bytes, _stream()))
    return keys, flag


_scan_tls = threading.local()       # per-thread dispatch state: callers may drive several indexes from several threads


class force_popc:
    """Context manager: ``hamming_scan_keys`` / ``hamming_topk`` stay on the XOR/POPC scan."""

    @staticmethod
    def on() -> bool:
        return bool(getattr(_scan_tls, "force_popc", False))

    def __enter__(self):
        self._prev = force_popc.on()
        _scan_tls.force_popc = True

    def __exit__(self, *exc):
        _scan_tls.force_popc = self._prev


#: longest per-query list the scan kernels keep; beyond it the exhaustive sorted path answers
SCAN_MAX_K = 2048
#: (query, row) keys materialised per call of the exhaustive path (16 bytes each + sort workspace)
SORTED_TOPK_MAX_PAIRS = 1 << 26


def hamming_topk_sorted_keys(db: torch.Tensor, q: torch.Tensor, k: int, idx_base: int = 0) -> torch.Tensor:
    """Packed keys int64[Q, k] for ANY k (``sb_hamming_topk_sorted``: all (query, row) keys, one radix
    sort) -- the reference's ``heapq.nsmallest`` takes any n (linear.py:232-240).  Queries are processed
    in chunks of at most ``SORTED_TOPK_MAX_PAIRS`` (query, row) pairs."""
    require_cuda()
    U, W = db.shape
    Q = q.shape[0]
    keys = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    if U == 0:
        return keys.fill_(-1)
    if U > SORTED_TOPK_MAX_PAIRS * 8:
        raise ValueError("n=%d neighbours over %d codes: the exhaustive path holds at most %d codes"
                         % (k, U, SORTED_TOPK_MAX_PAIRS * 8))
    lib = _lib.load()
    per = max(1, SORTED_TOPK_MAX_PAIRS // U)
    with torch.cuda.device(db.device):
        for q0 in range(0, Q, per):
            qc = min(per, Q - q0)
            ws_bytes = lib.sb_hamming_topk_sorted_workspace_bytes(U, qc)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=db.device)
            _lib.check(lib.sb_hamming_topk_sorted(_ptr(db), U, W, _ptr(q[q0:q0 + qc]), qc, k, idx_base,
                                                  _ptr(keys[q0:q0 + qc]), _ptr(ws), ws_bytes, _stream()))
    return keys


def decode_keys(keys: torch.Tensor):
    """Packed keys -> (dist int32, idx int64), -1 for empty slots (elementwise decode of the ABI's
    key format; the merge kernel does the same for lists it has to merge)."""
    empty = keys == -1
    dist = torch.where(empty, torch.full_like(keys, -1), (keys >> _lib.KEY_ROW_BITS) & 0xFFFFFF).to(torch.int32)
    idx = torch.where(empty, torch.full_like(keys, -1), keys & ((1 << _lib.KEY_ROW_BITS) - 1))
    return dist, idx


def hamming_scan_keys(db: torch.Tensor, q: torch.Tensor, k: int, idx_base: int = 0,
                      variant: int = 0) -> torch.Tensor:
    """Local top-k as packed keys int64[Q, k] (uint64 bit pattern, ascending,
    SB_KEY_EMPTY padded).  ``variant`` 0 picks the kernel: the tensor-core scan for large batches
    over large tables, the XOR/POPC scan otherwise (1, 2: POPC formulations, 3: tensor cores)."""
    require_cuda()
    _chk(db, torch.int32, "db")
    _chk(q, torch.int32, "q")
    U, W = db.shape
    Q = q.shape[0]
    if q.shape[1] != W:
        raise ValueError("query codes have %d words, table has %d" % (q.shape[1], W))
    if variant == SCAN_VARIANT_TC and not hamming_scan_tc_supported(U, W, Q, k):
        raise ValueError("tensor-core scan does not support U=%d W=%d Q=%d k=%d" % (U, W, Q, k))
    if k > SCAN_MAX_K:
        return hamming_topk_sorted_keys(db, q, k, idx_base)
    if variant == SCAN_VARIANT_TC or (variant == 0 and not force_popc.on() and _tc_scan_pays(U, W, Q, k)):
        # tensor-core scan, then the XOR/POPC scan PREDICATED on its overflow flag: every kernel of the second
        # scan returns at once unless a candidate buffer overflowed (tables of massively tied codes), in which
        # case it overwrites the keys with the exact result.  No host round trip, same keys either way.
        keys, flag = hamming_scan_keys_tc(db, q, k, idx_base)
        lib = _lib.load()
        ws_bytes = lib.sb_hamming_scan_workspace_bytes(U, W, Q, k)
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=db.device)
        with torch.cuda.device(db.device):
            _lib.check(lib.sb_hamming_scan_if(_ptr(db), U, W, _ptr(q), Q, k, idx_base, _ptr(keys), _ptr(flag),
                                              _ptr(ws), ws_bytes, _stream()))
        if not torch.cuda.is_current_stream_capturing():
            acc = _overflow_acc.get(db.device.index)
            if acc is None:
                acc = _overflow_acc[db.device.index] = torch.zeros((1,), dtype=torch.int32, device=db.device)
            acc.add_(flag)
        return keys
    if variant == SCAN_VARIANT_TC:
        variant = 0
    lib = _lib.load()
    ws_bytes = lib.sb_hamming_scan_workspace_bytes(U, W, Q, k)
    ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=db.device)
    keys = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(lib.sb_hamming_scan_variant(_ptr(db), U, W, _ptr(q), Q, k, idx_base, _ptr(keys),
                                               _ptr(ws), ws_bytes, variant, _stream()))
    return keys


def topk_merge(keys: torch.Tensor, want_keys: bool = False):
    """keys int64[parts, Q, k] -> (dist int32[Q, k], idx int64[Q, k]) in canonical
    (distance, row) order; -1 marks empty slots."""
    require_cuda()
    _chk(keys, torch.int64, "keys")
    parts, Q, k = keys.shape
    dist = torch.empty((Q, k), dtype=torch.int32, device=keys.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=keys.device)
    okeys = torch.empty((Q, k), dtype=torch.int64, device=keys.device) if want_keys else None
    with torch.cuda.device(keys.device):
        _lib.check(_lib.load().sb_topk_merge(_ptr(keys), parts, Q, k, _ptr(okeys), _ptr(dist), _ptr(idx), _stream()))
    return (dist, idx, okeys) if want_keys else (dist, idx)


def hamming_topk(db: torch.Tensor, q: torch.Tensor, k: int, idx_base: int = 0):
    """(dist int32[Q, k], idx int64[Q, k]); -1 where the table has fewer than k rows."""
    require_cuda()
    _chk(db, torch.int32, "db")
    _chk(q, torch.int32, "q")
    U, W = db.shape
    Q = q.shape[0]
    if q.shape[1] != W:
        raise ValueError("query codes have %d words, table has %d" % (q.shape[1], W))
    if k > SCAN_MAX_K:
        return decode_keys(hamming_topk_sorted_keys(db, q, k, idx_base))
    if not force_popc.on() and _tc_scan_pays(U, W, Q, k):
        # large batch over a large table: tensor-core scan (falls back to XOR/POPC on overflow), then decode
        return topk_merge(hamming_scan_keys(db, q, k, idx_base).unsqueeze(0))
    lib = _lib.load()
    ws_bytes = lib.sb_hamming_scan_workspace_bytes(U, W, Q, k)
    ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=db.device)
    dist = torch.empty((Q, k), dtype=torch.int32, device=db.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(lib.sb_hamming_topk(_ptr(db), U, W, _ptr(q), Q, k, idx_base, _ptr(dist), _ptr(idx),
                                       _ptr(ws), ws_bytes, _stream()))
    return dist, idx


# --------------------------------------------------------------------- stage 3
def rerank(db: torch.Tensor, q: torch.Tensor, cand_idx: torch.Tensor, cand_off: torch.Tensor,
           metric: str, pitch: int = 0) -> torch.Tensor:
    """float64[M] distances of each candidate row to its query.  ``pitch`` > 0: the candidates are in
    the fixed-pitch layout of ``expand_candidates`` (M = Q * pitch)."""
    require_cuda()
    if db.dtype != torch.float32 or q.dtype != torch.float32 or db.stride(1) != 1 or q.stride(1) != 1:
        raise ValueError("db and q must be float32 with unit column stride")
    _chk(cand_idx, torch.int64, "cand_idx")
    _chk(cand_off, torch.int64, "cand_off")
    if metric not in _lib.METRICS:
        raise ValueError("Invalid distance method label. Must be one of "
                         "['euclidean' | 'cosine' | 'hik']")
    N, D = db.shape
    Q = q.shape[0]
    if q.shape[1] != D or cand_off.numel() != Q + 1:
        raise ValueError("shape mismatch between db, q and cand_off")
    M = cand_idx.numel()
    out = torch.empty((M,), dtype=torch.float64, device=db.device)
    with torch.cuda.device(db.device):
        if pitch > 0:
            if M != Q * pitch:
                raise ValueError("fixed-pitch candidates: %d slots for %d queries x pitch %d" % (M, Q, pitch))
            _lib.check(_lib.load().sb_rerank_pitched(_ptr(db), N, D, max(db.stride(0), D), _ptr(q), Q, max(q.stride(0), D),
                                                     _ptr(cand_idx), _ptr(cand_off), pitch, _lib.METRICS[metric], _ptr(out),
                                                     _stream()))
        else:
            _lib.check(_lib.load().sb_rerank(_ptr(db), N, D, max(db.stride(0), D), _ptr(q), Q, max(q.stride(0), D),
                                             _ptr(cand_idx), _ptr(cand_off), M, _lib.METRICS[metric], _ptr(out),
                                             _stream()))
    return out


def rerank_shard(db: torch.Tensor, row_base: int, q: torch.Tensor, cand_idx: torch.Tensor, cand_off: torch.Tensor,
                 metric: str) -> torch.Tensor:
    """float64[M] distances of the candidates whose GLOBAL row lies in
    [row_base, row_base + len(db)); 0.0 for the others (sum over shards = all distances)."""
    require_cuda()
    if db.dtype != torch.float32 or q.dtype != torch.float32 or db.stride(1) != 1 or q.stride(1) != 1:
        raise ValueError("db and q must be float32 with unit column stride")
    _chk(cand_idx, torch.int64, "cand_idx")
    _chk(cand_off, torch.int64, "cand_off")
    N, D = db.shape
    Q = q.shape[0]
    M = cand_idx.numel()
    out = torch.empty((M,), dtype=torch.float64, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(_lib.load().sb_rerank_shard(_ptr(db), N, row_base, D, max(db.stride(0), D), _ptr(q), Q,
                                               max(q.stride(0), D), _ptr(cand_idx), _ptr(cand_off), M,
                                               _lib.METRICS[metric], _ptr(out), _stream()))
    return out


def rerank_peer(shards, q: torch.Tensor, cand_idx: torch.Tensor, cand_off: torch.Tensor, metric: str,
                pitch: int = 0) -> torch.Tensor:
    """float64[M] distances of candidate GLOBAL rows against a row-sharded table whose shards live on
    several GPUs (``peer.PeerShards``): each row is loaded from the GPU that holds it (NVLink)."""
    require_cuda()
    if q.dtype != torch.float32 or q.stride(1) != 1:
        raise ValueError("q must be float32 with unit column stride")
    _chk(cand_idx, torch.int64, "cand_idx")
    _chk(cand_off, torch.int64, "cand_off")
    if metric not in _lib.METRICS:
        raise ValueError("Invalid distance method label. Must be one of "
                         "['euclidean' | 'cosine' | 'hik']")
    Q, D = q.shape
    if D != shards.dim or cand_off.numel() != Q + 1:
        raise ValueError("shape mismatch between the shards, q and cand_off")
    M = cand_idx.numel()
    out = torch.empty((M,), dtype=torch.float64, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().sb_rerank_peer(_ptr(shards.ptr_table), _ptr(shards.bound_table), shards.n_shards, D,
                                              shards.ld, _ptr(q), Q, max(q.stride(0), D), _ptr(cand_idx), _ptr(cand_off), M,
                                              pitch, _lib.METRICS[metric], 1 if shards.aligned16 else 0, _ptr(out), _stream()))
    return out


def expand_candidates(code_rows: torch.Tensor, csr_off: torch.Tensor, csr_rows: torch.Tensor, pitch: int):
    """Fixed-pitch candidate expansion on the device (no host sync):
    (cand_idx int64[Q * pitch] padded with -1, cand_off int64[Q + 1], cand_cnt int64[Q])."""
    require_cuda()
    _chk(code_rows, torch.int64, "code_rows")
    _chk(csr_off, torch.int64, "csr_off")
    _chk(csr_rows, torch.int64, "csr_rows")
    Q, n = code_rows.shape
    cand_idx = torch.empty((Q * pitch,), dtype=torch.int64, device=code_rows.device)
    cand_off = torch.empty((Q + 1,), dtype=torch.int64, device=code_rows.device)
    cand_cnt = torch.empty((Q,), dtype=torch.int64, device=code_rows.device)
    with torch.cuda.device(code_rows.device):
        _lib.check(_lib.load().sb_expand_candidates(_ptr(code_rows), Q, n, _ptr(csr_off), _ptr(csr_rows), pitch,
                                                    _ptr(cand_idx), _ptr(cand_off), _ptr(cand_cnt), _stream()))
    return cand_idx, cand_off, cand_cnt


#: longest candidate list of ONE query the rank-count selection kernel orders (it stages the list in
#: shared memory and counts, O(m^2)); longer lists take the radix-sorted selection
SELECT_RANK_MAX = 4096


def _select_sorted(dist, cand_off, cand_cnt, cand_idx, n: int, tie_by_row: bool):
    Q = cand_off.numel() - 1
    M = dist.numel()
    lib = _lib.load()
    pos = torch.empty((Q, n), dtype=torch.int64, device=dist.device)
    od = torch.empty((Q, n), dtype=torch.float64, device=dist.device)
    if M == 0:
        return pos.fill_(-1), od.fill_(float("nan"))
    ws_bytes = lib.sb_rerank_select_sorted_workspace_bytes(M)
    if ws_bytes == 0:
        raise ValueError("candidate list of %d entries is beyond the selection's range" % M)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(lib.sb_rerank_select_sorted(_ptr(dist), _ptr(cand_off), _ptr(cand_cnt), _ptr(cand_idx), M, Q, n,
                                               1 if tie_by_row else 0, _ptr(pos), _ptr(od), _ptr(ws), ws_bytes, _stream()))
    return pos, od


def rerank_select_rows(dist: torch.Tensor, cand_off: torch.Tensor, cand_cnt: Optional[torch.Tensor],
                       cand_idx: torch.Tensor, n: int, tie_by_row: bool = False, max_m: Optional[int] = None):
    """Per query the first ``n`` candidates by (distance, position), as candidate ROWS:
    (rows int64[Q, n] or -1, dist float64[Q, n]).  ``max_m``: an upper bound on one query's list
    length when the caller knows it (default: the total) -- lists beyond ``SELECT_RANK_MAX`` are
    ordered by the radix-sorted selection."""
    require_cuda()
    _chk(dist, torch.float64, "dist")
    _chk(cand_off, torch.int64, "cand_off")
    _chk(cand_idx, torch.int64, "cand_idx")
    Q = cand_off.numel() - 1
    if (dist.numel() if max_m is None else max_m) > SELECT_RANK_MAX:
        return _select_sorted(dist, cand_off, cand_cnt, cand_idx, n, tie_by_row)
    rows = torch.empty((Q, n), dtype=torch.int64, device=dist.device)
    od = torch.empty((Q, n), dtype=torch.float64, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(_lib.load().sb_rerank_select_rows(_ptr(dist), _ptr(cand_off), _ptr(cand_cnt), _ptr(cand_idx), Q, n,
                                                     1 if tie_by_row else 0, _ptr(rows), _ptr(od), _stream()))
    return rows, od


def rerank_select(dist: torch.Tensor, cand_off: torch.Tensor, n: int, max_m: Optional[int] = None):
    """Per query the first ``n`` candidates by (distance, position):
    (pos int64[Q, n] into the candidate list or -1, dist float64[Q, n])."""
    require_cuda()
    _chk(dist, torch.float64, "dist")
    _chk(cand_off, torch.int64, "cand_off")
    Q = cand_off.numel() - 1
    if (dist.numel() if max_m is None else max_m) > SELECT_RANK_MAX:
        return _select_sorted(dist, cand_off, None, None, n, False)
    pos = torch.empty((Q, n), dtype=torch.int64, device=dist.device)
    od = torch.empty((Q, n), dtype=torch.float64, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(_lib.load().sb_rerank_select(_ptr(dist), _ptr(cand_off), Q, n, _ptr(pos), _ptr(od), _stream()))
    return pos, od


# --------------------------------------------------------------------- flat L2 index
#: queries of the last ``l2_topk`` call that overflowed their survivor buffer and were redone by
#: brute force (diagnostics / tests)
LAST_L2_OVERFLOW = 0


def l2_prepare(db: torch.Tensor):
    """(xn f32[N] = |row|^2, xn_max f32[1]) of a float32 table (once per table)."""
    require_cuda()
    if db.dtype != torch.float32 or db.dim() != 2 or db.stride(1) != 1:
        raise ValueError("db must be float32[N, D] with unit column stride")
    N, D = db.shape
    xn = torch.empty((N,), dtype=torch.float32, device=db.device)
    xmax = torch.empty((1,), dtype=torch.float32, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(_lib.load().sb_l2_prepare(_ptr(db), N, D, max(db.stride(0), D), _ptr(xn), _ptr(xmax), _stream()))
    return xn, xmax


def l2_brute_force(db: torch.Tensor, q: torch.Tensor, k: int):
    """Exact L2 top-k by the re-rank kernel over EVERY row (any shape; the fallback of
    ``l2_topk`` and the path for shapes the tensor-core filter does not take).
    (idx int64[Q, k], dist float64[Q, k]); -1 / NaN when the table has fewer than k rows."""
    N = db.shape[0]
    Q = q.shape[0]
    idx = torch.full((Q, k), -1, dtype=torch.int64, device=db.device)
    dist = torch.full((Q, k), float("nan"), dtype=torch.float64, device=db.device)
    if N == 0:
        return idx, dist
    # query chunks of at most 2^25 (query, row) distances: candidate list of query j = every row
    per = max(1, (1 << 25) // N)
    for q0 in range(0, Q, per):
        qc = min(per, Q - q0)
        cand = torch.arange(N, dtype=torch.int64, device=db.device).repeat(qc)
        off = torch.arange(qc + 1, dtype=torch.int64, device=db.device) * N
        d = rerank(db, q[q0:q0 + qc], cand, off, "euclidean")
        r, od = rerank_select_rows(d, off, None, cand, k, tie_by_row=True, max_m=N)
        idx[q0:q0 + qc], dist[q0:q0 + qc] = r, od
    return idx, dist


def l2_topk(db: torch.Tensor, q: torch.Tensor, k: int, prepared=None):
    """Exact k nearest rows by euclidean distance, (distance, row) order:
    (idx int64[Q, k], dist float64[Q, k]).  ``prepared`` = ``l2_prepare(db)`` (cached by the
    index).  Queries whose survivor buffer overflowed are redone by brute force."""
    require_cuda()
    if db.dtype != torch.float32 or q.dtype != torch.float32 or db.stride(1) != 1 or q.stride(1) != 1:
        raise ValueError("db and q must be float32 with unit column stride")
    N, D = db.shape
    Q = q.shape[0]
    if q.shape[1] != D:
        raise ValueError("query dimension %d != table dimension %d" % (q.shape[1], D))
    lib = _lib.load()
    ldd = max(db.stride(0), D)
    if not lib.sb_l2_topk_supported(N, D, ldd, k) or db.data_ptr() % 16 != 0:
        return l2_brute_force(db, q, k)
    xn, xmax = prepared if prepared is not None else l2_prepare(db)
    ws_bytes = lib.sb_l2_topk_workspace_bytes(D, Q, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=db.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    dist = torch.empty((Q, k), dtype=torch.float64, device=db.device)
    overflow = torch.empty((Q,), dtype=torch.int32, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(lib.sb_l2_topk(_ptr(db), N, D, ldd, _ptr(xn), _ptr(xmax), _ptr(q), Q, max(q.stride(0), D), k,
                                  _ptr(idx), _ptr(dist), _ptr(overflow), _ptr(ws), ws_bytes, _stream()))
    global LAST_L2_OVERFLOW
    bad = torch.nonzero(overflow).reshape(-1)
    LAST_L2_OVERFLOW = int(bad.numel())
    if bad.numel():
        bi, bd = l2_brute_force(db, q[bad].contiguous(), k)
        idx[bad], dist[bad] = bi, bd
    return idx, dist
