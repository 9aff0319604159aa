"""
Multi-GPU LSH index: one process per GPU (``torch.distributed``, NCCL over
NVLink 5 / NVSwitch; gloo on CPU for tests).

Sharding (DESIGN.md section "Multi-GPU"):
  * the descriptor matrix is sharded by ROW RANGE: rank r holds the rows it
    ingested; global row id = rank-major concatenation;
  * per-row codes are all-gathered once at build time (32 B per row at 256 bits)
    and every rank derives the SAME global sorted-unique code table and
    code -> rows CSR -- exactly the single-GPU structures;
  * the Hamming scan is partitioned either over that table (``scan_partition="rows"``: rank r
    scans table rows [lo_r, hi_r) with ``idx_base = lo_r``, so its keys are already global) or over
    the QUERIES (``"queries"``: every rank holds the whole table anyway -- 32 B per code -- and scans
    all of it for its slice of the batch; its keys are final, the all-gather assembles the batch and
    nothing is merged).  Same keys either way.  The per-batch costs that do not shrink with the table
    (threshold bookkeeping, survivor re-checks of the tensor-core scan) shrink with the query slice,
    so ``"auto"`` partitions by queries once every rank gets at least 64 of them.

Per query batch there are two small collectives:
  1. all-gather of the per-rank top-n keys (Q*n*8 bytes per rank) followed by
     ``sb_topk_merge`` -> the global n nearest unique codes, identical on all ranks;
  2. all-reduce (SUM) of candidate distances: every rank re-ranks the candidate rows
     that live in its descriptor shard and contributes an exact 0.0 for the others.
Because keys and candidate order are global, the result is bit-identical to the
single-GPU index, ties included.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import codes as codeops
from .engine import _stage, expand_candidates


class DeviceOps:
    """The CUDA implementation of the per-rank compute steps (default)."""

    def __init__(self, functor, distance_method: str) -> None:
        self.functor = functor
        self.distance_method = distance_method

    def hash(self, x: torch.Tensor) -> torch.Tensor:
        return self.functor.get_hash_packed(x)

    def scan_keys(self, table: torch.Tensor, q_codes: torch.Tensor, n: int, idx_base: int) -> torch.Tensor:
        from . import device
        return device.hamming_scan_keys(table, q_codes, n, idx_base)

    def merge_keys(self, keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        from . import device
        return device.topk_merge(keys)

    def rerank(self, x, q, cand_idx, cand_off) -> torch.Tensor:
        from . import device
        return device.rerank(x, q, cand_idx, cand_off, self.distance_method)

    def rerank_select(self, d, cand_off, n):
        from . import device
        return device.rerank_select(d, cand_off, n)

    def rerank_shard(self, x, row_base, q, cand_idx, cand_off) -> torch.Tensor:
        from . import device
        return device.rerank_shard(x, row_base, q, cand_idx, cand_off, self.distance_method)

    def expand(self, code_rows, csr_off, csr_rows, pitch):
        from . import device
        return device.expand_candidates(code_rows.contiguous(), csr_off, csr_rows, pitch)

    def rerank_select_rows(self, d, cand_off, cand_cnt, cand_idx, n):
        from . import device
        # fixed-pitch layout: every query owns numel / Q slots
        return device.rerank_select_rows(d, cand_off, cand_cnt, cand_idx, n,
                                         max_m=cand_idx.numel() // max(cand_off.numel() - 1, 1))


def _all_gather_rows(t: torch.Tensor, group) -> Tuple[torch.Tensor, List[int]]:
    """Concatenate row blocks of differing length from all ranks (rank-major)."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0), sizes


def partition_bounds(total: int, world: int, align: int = 4) -> List[int]:
    """``world + 1`` ascending cut points over ``total`` table rows; interior cuts
    are multiples of ``align`` rows so every slice starts 16-byte aligned."""
    cuts = [0]
    for r in range(1, world):
        c = (total * r // world) // align * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(total)
    return cuts


class ShardedLshIndex:
    """Row-sharded descriptors, range-partitioned Hamming scan, exact merge."""

    #: a rank's query slice must be at least this long for ``scan_partition="auto"`` to split the batch
    MIN_QUERIES_PER_RANK = 64

    def __init__(self, functor, distance_method: str = "euclidean", group=None, ops=None,
                 scan_partition: str = "auto") -> None:
        if scan_partition not in ("auto", "rows", "queries"):
            raise ValueError("scan_partition must be 'auto', 'rows' or 'queries'")
        self.scan_partition = scan_partition
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ops = ops if ops is not None else DeviceOps(functor, distance_method)
        self.x_local: Optional[torch.Tensor] = None
        self.row_bounds: List[int] = []
        self.table = self.csr_off = self.csr_rows = None
        self.max_rows_per_code = 0
        self.scan_lo = self.scan_hi = 0

    @property
    def num_rows(self) -> int:
        return self.row_bounds[-1] if self.row_bounds else 0

    @property
    def num_codes(self) -> int:
        return 0 if self.table is None else int(self.table.shape[0])

    def build(self, x_local: torch.Tensor) -> None:
        """Collective: every rank passes the descriptor rows it owns."""
        self.x_local = x_local
        codes_local = self.ops.hash(x_local)
        codes_all, sizes = _all_gather_rows(codes_local, self.group)
        self.row_bounds = [0]
        for s in sizes:
            self.row_bounds.append(self.row_bounds[-1] + s)
        self.table, _, self.csr_off, self.csr_rows, self.max_rows_per_code = codeops.build_table(codes_all, with_max=True)
        cuts = partition_bounds(int(self.table.shape[0]), self.world)
        self.scan_lo, self.scan_hi = cuts[self.rank], cuts[self.rank + 1]

    def near_codes(self, q_codes: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global n nearest unique codes: local scan, all-gather, merge."""
        Q = q_codes.shape[0]
        by_queries = self.scan_partition == "queries" or (
            self.scan_partition == "auto" and Q >= self.MIN_QUERIES_PER_RANK * self.world)
        if by_queries and self.world > 1:
            per = (Q + self.world - 1) // self.world              # equal slices (the last ones padded) for the collective
            lo = min(Q, self.rank * per)
            hi = min(Q, lo + per)
            mine = q_codes[lo:hi]
            if hi - lo < per:
                mine = torch.cat([mine, torch.zeros((per - (hi - lo), q_codes.shape[1]), dtype=q_codes.dtype,
                                                    device=q_codes.device)], dim=0)
            with _stage("hamming_scan"):
                keys = self.ops.scan_keys(self.table, mine.contiguous(), n, 0).contiguous()
            with _stage("allgather_merge"):
                gathered = torch.empty((self.world * per,) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
                dist.all_gather_into_tensor(gathered, keys, group=self.group)  # rank-major = query order
                return self.ops.merge_keys(gathered[:Q].contiguous().unsqueeze(0))   # one part: decode only
        with _stage("hamming_scan"):
            keys = self.ops.scan_keys(self.table[self.scan_lo:self.scan_hi], q_codes, n, self.scan_lo).contiguous()
        with _stage("allgather_merge"):
            Q = keys.shape[0]
            gathered = torch.empty((self.world * Q,) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
            dist.all_gather_into_tensor(gathered, keys, group=self.group)      # rank-major concatenation
            return self.ops.merge_keys(gathered.view((self.world, Q) + tuple(keys.shape[1:])))

    def query(self, q: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Collective: all ranks pass the SAME queries; all get the full result
        (global rows int64[Q, n], dists f64[Q, n])."""
        from .engine import FIXED_PITCH_LIMIT
        with _stage("itq_hash"):
            q_codes = self.ops.hash(q)
        from . import device
        with device.deferred_scan_check() as chk:
            out = self._query_hashed(q, q_codes, n)
            # every rank must take the same branch afterwards: fold the flags (a rank that did not use the
            # tensor-core scan contributes 0); still no host synchronisation
            flag = chk.flag_tensor()
            if flag is None:
                flag = torch.zeros((1,), dtype=torch.int32, device=q.device)
                chk.flags = [flag]
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if chk.overflowed():
            with device.force_popc():
                out = self._query_hashed(q, q_codes, n)
        return out

    def _query_hashed(self, q: torch.Tensor, q_codes: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        from .engine import FIXED_PITCH_LIMIT
        _, code_rows = self.near_codes(q_codes, n)
        pitch = n * max(self.max_rows_per_code, 1)
        fixed = hasattr(self.ops, "expand") and n <= 2048 and q.shape[0] * pitch <= FIXED_PITCH_LIMIT
        cand_cnt = None
        with _stage("expand"):
            if fixed:       # device-only, fixed-pitch segments padded with -1 (no host round trip)
                cand_idx, cand_off, cand_cnt = self.ops.expand(code_rows, self.csr_off, self.csr_rows, pitch)
            else:
                cand_idx, cand_off = expand_candidates(code_rows, self.csr_off, self.csr_rows)
        lo, hi = self.row_bounds[self.rank], self.row_bounds[self.rank + 1]
        with _stage("rerank"):
            # every candidate row lives in exactly one shard: its owner computes the distance, the
            # others contribute an exact 0.0, so a SUM all-reduce assembles the vector bit-exactly
            if hasattr(self.ops, "rerank_shard"):
                d = self.ops.rerank_shard(self.x_local, lo, q, cand_idx, cand_off)
            else:
                d = self.ops.rerank(self.x_local, q, (cand_idx - lo).contiguous(), cand_off)
                d = torch.where((cand_idx >= lo) & (cand_idx < hi), d, torch.zeros_like(d))
        with _stage("allreduce_dist"):
            if d.numel():
                dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.group)
        with _stage("rerank"):
            if fixed:
                return self.ops.rerank_select_rows(d, cand_off, cand_cnt, cand_idx, n)
            pos, od = self.ops.rerank_select(d, cand_off, n)
        rows = torch.where(pos >= 0, cand_idx[pos.clamp(min=0)], pos) if cand_idx.numel() else pos
        return rows, od


# ----------------------------------------------------------------------------- flat L2 index
class DeviceFlatOps:
    """CUDA implementation of the per-rank steps of the sharded flat L2 index."""

    def prepare(self, x: torch.Tensor):
        from . import device
        return device.l2_prepare(x) if len(x) else None

    def l2_topk(self, x, q, n, prepared):
        from . import device
        if len(x) == 0:
            return (torch.full((q.shape[0], n), -1, dtype=torch.int64, device=q.device),
                    torch.full((q.shape[0], n), float("nan"), dtype=torch.float64, device=q.device))
        return device.l2_topk(x, q, n, prepared=prepared)

    def select(self, dist_all: torch.Tensor, idx_all: torch.Tensor, n: int):
        """Per query the n best of ``m`` (distance, global row) pairs, ties by row; NaN / -1 last."""
        from . import device
        Q, m = dist_all.shape
        off = torch.arange(Q + 1, dtype=torch.int64, device=dist_all.device) * m
        return device.rerank_select_rows(dist_all.reshape(-1).contiguous(), off, None,
                                         idx_all.reshape(-1).contiguous(), n, tie_by_row=True, max_m=m)


class ShardedFlatL2Index:
    """Row-sharded exact L2 index (BASELINE config 3: 100M x 128-d over 8 GPUs).  Each rank
    searches its own rows (local exact top-n), one all-gather of ``Q*n*(8+8)`` bytes per rank,
    then the (distance, global row) merge -- identical to the single-device result."""

    def __init__(self, group=None, ops=None) -> None:
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ops = ops if ops is not None else DeviceFlatOps()
        self.x_local: Optional[torch.Tensor] = None
        self.prepared = None
        self.row_bounds: List[int] = []

    @property
    def num_rows(self) -> int:
        return self.row_bounds[-1] if self.row_bounds else 0

    def build(self, x_local: torch.Tensor) -> None:
        """Collective: every rank passes the descriptor rows it owns (rank-major global rows)."""
        self.x_local = x_local
        n_local = torch.tensor([x_local.shape[0]], dtype=torch.int64, device=x_local.device)
        sizes = [torch.zeros_like(n_local) for _ in range(self.world)]
        dist.all_gather(sizes, n_local, group=self.group)
        self.row_bounds = [0]
        for s in sizes:
            self.row_bounds.append(self.row_bounds[-1] + int(s.item()))
        self.prepared = self.ops.prepare(x_local)

    def query(self, q: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Collective: all ranks pass the SAME queries; all get (global rows int64[Q, n], dists f64[Q, n])."""
        lo = self.row_bounds[self.rank]
        with _stage("l2_topk"):
            idx, d = self.ops.l2_topk(self.x_local, q, n, self.prepared)
            idx = torch.where(idx >= 0, idx + lo, idx)
        with _stage("allgather_merge"):
            Q = q.shape[0]
            g_idx = torch.empty((self.world * Q, n), dtype=idx.dtype, device=idx.device)
            g_d = torch.empty((self.world * Q, n), dtype=d.dtype, device=d.device)
            dist.all_gather_into_tensor(g_idx, idx.contiguous(), group=self.group)
            dist.all_gather_into_tensor(g_d, d.contiguous(), group=self.group)
            # [world, Q, n] -> [Q, world * n]
            g_idx = g_idx.view(self.world, Q, n).permute(1, 0, 2).reshape(Q, self.world * n)
            g_d = g_d.view(self.world, Q, n).permute(1, 0, 2).reshape(Q, self.world * n)
            return self.ops.select(g_d, g_idx, n)
