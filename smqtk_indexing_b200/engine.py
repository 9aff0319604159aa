"""
Device-resident LSH index: the array-form state behind
``LSHNearestNeighborIndex.nn_batch`` -- descriptor matrix, per-row packed codes,
sorted unique code table and the ``code -> rows`` CSR that replace the
reference's ``set`` of ints (linear.py:163) and ``hash2uuids`` KVS
(lsh.py:316-323) on the query path.

Query pipeline (reference lsh.py:470-519 for a whole batch, nothing leaves HBM):
  sb_itq_hash -> sb_hamming_topk over the unique table -> CSR expansion of the
  n nearest codes into candidate rows -> sb_rerank -> sb_rerank_select.
"""
from typing import Optional, Tuple

import torch

from . import codes as codeops
from . import device

#: Largest Q * pitch (candidate slots) served by the fixed-pitch device path.
FIXED_PITCH_LIMIT = 1 << 25

#: When set to a list, every query stage appends (name, start_event, end_event)
#: recorded on the current stream -- used by bench.py to attribute step time to
#: kernels without a profiler.  None (default) = no events are recorded.
STAGE_EVENTS = None


class _stage:
    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self):
        if STAGE_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STAGE_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STAGE_EVENTS.append((self.name, self.e0, e1))
        return False


def expand_candidates(code_rows: torch.Tensor, csr_off: torch.Tensor, csr_rows: torch.Tensor):
    """Ragged expansion ``near codes -> descriptor rows`` on the device.

    :param code_rows: int64[Q, n] rows of the unique-code table (-1 = none)
    :return: (cand_idx int64[M], cand_off int64[Q + 1]); candidates of a query are
        listed in (code rank, row) order.
    """
    Q, n = code_rows.shape
    flat = code_rows.reshape(-1)
    ok = flat >= 0
    safe = torch.where(ok, flat, torch.zeros_like(flat))
    start = csr_off[safe]
    cnt = torch.where(ok, csr_off[safe + 1] - start, torch.zeros_like(start))
    seg_end = torch.cumsum(cnt, 0)
    total = int(seg_end[-1].item()) if flat.numel() else 0
    cand_off = torch.zeros(Q + 1, dtype=torch.int64, device=flat.device)
    cand_off[1:] = seg_end.reshape(Q, n)[:, -1]
    if total == 0:
        return torch.empty(0, dtype=torch.int64, device=flat.device), cand_off
    seg = torch.repeat_interleave(torch.arange(flat.numel(), device=flat.device), cnt, output_size=total)
    within = torch.arange(total, device=flat.device) - (seg_end - cnt)[seg]
    return csr_rows[start[seg] + within], cand_off


class DeviceLshIndex:
    """Descriptor rows + codes + unique-code table + CSR, all on one device."""

    def __init__(self) -> None:
        self.x: Optional[torch.Tensor] = None          # float32[N, D]
        self.codes: Optional[torch.Tensor] = None      # int32[N, W]
        self.table: Optional[torch.Tensor] = None      # int32[U, W] sorted unique
        self.row_code: Optional[torch.Tensor] = None   # int64[N] row -> table row
        self.csr_off: Optional[torch.Tensor] = None    # int64[U + 1]
        self.csr_rows: Optional[torch.Tensor] = None   # int64[N]
        self.max_rows_per_code: int = 0

    def clear(self) -> None:
        self.__init__()

    @property
    def num_rows(self) -> int:
        return 0 if self.codes is None else int(self.codes.shape[0])

    @property
    def num_codes(self) -> int:
        return 0 if self.table is None else int(self.table.shape[0])

    def set_rows(self, x: Optional[torch.Tensor], codes: torch.Tensor) -> None:
        """Adopt descriptor rows (may be None for a codes-only index) and their codes."""
        self.x, self.codes = x, codes
        self.reindex()

    def reindex(self) -> None:
        """Recompute the unique table and the CSR from ``codes``."""
        if self.codes is None or self.codes.shape[0] == 0:
            self.table = self.csr_off = self.csr_rows = self.row_code = None
            return
        self.table, self.row_code, self.csr_off, self.csr_rows = codeops.build_table(self.codes)
        # most rows sharing one code: sizes the fixed-pitch candidate segments (one sync per re-index)
        self.max_rows_per_code = int((self.csr_off[1:] - self.csr_off[:-1]).max().item())

    # ------------------------------------------------------------------ query stages
    def near_codes(self, q_codes: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(dist int32[Q, n], table row int64[Q, n]) of the n nearest unique codes."""
        table = self.table
        w = max(table.shape[1], q_codes.shape[1])
        with _stage("hamming_scan"):
            return device.hamming_topk(codeops.widen(table, w), codeops.widen(q_codes, w).contiguous(), n)

    def rerank(self, q: torch.Tensor, code_rows: torch.Tensor, n: int, distance_method: str):
        """Candidate rows of the given codes, re-ranked: (rows int64[Q, n], dists f64[Q, n])."""
        Q = code_rows.shape[0]
        pitch = n * max(self.max_rows_per_code, 1)
        if n <= 2048 and Q * pitch <= FIXED_PITCH_LIMIT:
            # device-only path: fixed-pitch segments, no host round trip
            with _stage("expand"):
                cand_idx, cand_off, cand_cnt = device.expand_candidates(code_rows.contiguous(), self.csr_off,
                                                                        self.csr_rows, pitch)
            with _stage("rerank"):
                dist = device.rerank(self.x, q, cand_idx, cand_off, distance_method)
                return device.rerank_select_rows(dist, cand_off, cand_cnt, cand_idx, n)
        # heavy code collisions: exact-size ragged expansion (needs the total on the host)
        with _stage("expand"):
            cand_idx, cand_off = expand_candidates(code_rows, self.csr_off, self.csr_rows)
        with _stage("rerank"):
            dist = device.rerank(self.x, q, cand_idx, cand_off, distance_method)
            pos, od = device.rerank_select(dist, cand_off, n)
        if cand_idx.numel() == 0:
            return pos, od
        rows = torch.where(pos >= 0, cand_idx[pos.clamp(min=0)], pos)
        return rows, od

    def query(self, functor, q: torch.Tensor, n: int, distance_method: str):
        """hash -> Hamming top-n unique codes -> candidates -> re-rank -> top-n."""
        with _stage("itq_hash"):
            q_codes = functor.get_hash_packed(q)
        # the tensor-core scan's overflow flag is read once, after the last stage has been launched
        with device.deferred_scan_check() as chk:
            _, code_rows = self.near_codes(q_codes, n)
            out = self.rerank(q, code_rows, n, distance_method)
        if chk.overflowed():
            with device.force_popc():
                _, code_rows = self.near_codes(q_codes, n)
                out = self.rerank(q, code_rows, n, distance_method)
        return out
