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
    scans table rows [lo_r, hi_r) with ``idx_base = lo_r``, so its keys are already global; one
    all-gather of the per-rank top-n keys + ``sb_topk_merge`` gives the global n nearest codes) or over
    the QUERIES (``"queries"``: every rank holds the whole table anyway -- 32 B per code -- and scans
    all of it for its slice of the batch; its keys are final).  Same keys either way.  The per-batch
    costs that do not shrink with the table (threshold bookkeeping, survivor re-checks of the
    tensor-core scan) shrink with the query slice, so ``"auto"`` partitions by queries once every
    rank gets at least 64 of them.

Re-rank, two modes with identical results:
  * ``rerank="peer"`` (default on CUDA when the GPUs can reach each other): every rank finishes ITS
    slice of the batch alone -- candidate expansion, distances, selection -- loading each candidate
    row from the GPU that owns it through CUDA-IPC mappings (``peer.py``, ``sb_rerank_peer``: NVLink
    loads, no collective), then ONE all-gather assembles the (row, distance) results of all slices.
    Per batch: 1 collective with the scan partitioned over queries, 2 (keys, results) over rows; no
    host synchronisation anywhere (an overflowed tensor-core scan is redone by a predicated kernel).
  * ``rerank="allreduce"`` (any backend, the gloo tests): every rank expands ALL candidates, computes the
    distances of the rows it owns (exact 0.0 elsewhere) and a SUM all-reduce assembles the vector.
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

    def rerank_peer(self, shards, q, cand_idx, cand_off, pitch: int = 0) -> torch.Tensor:
        from . import device
        return device.rerank_peer(shards, q, cand_idx, cand_off, self.distance_method, pitch)

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
    """Row-sharded descriptors, partitioned Hamming scan, exact merge."""

    #: a rank's query slice must be at least this long for ``scan_partition="auto"`` to split the batch
    MIN_QUERIES_PER_RANK = 64
    #: capture the pipeline of a (Q, n) shape into a CUDA graph once it has been asked for this often
    GRAPH_AFTER = 2

    def __init__(self, functor, distance_method: str = "euclidean", group=None, ops=None,
                 scan_partition: str = "auto", rerank: str = "auto", graph: bool = True) -> None:
        if scan_partition not in ("auto", "rows", "queries"):
            raise ValueError("scan_partition must be 'auto', 'rows' or 'queries'")
        if rerank not in ("auto", "peer", "allreduce"):
            raise ValueError("rerank must be 'auto', 'peer' or 'allreduce'")
        self.scan_partition = scan_partition
        self.rerank_mode = rerank
        self.graph = graph
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ops = ops if ops is not None else DeviceOps(functor, distance_method)
        self.x_local: Optional[torch.Tensor] = None
        self.row_bounds: List[int] = []
        self.table = self.csr_off = self.csr_rows = None
        self.max_rows_per_code = 0
        self.scan_lo = self.scan_hi = 0
        self.peers = None                 # peer.PeerShards when the re-rank reads the owners' HBM directly
        self.peer_error: Optional[str] = None
        self._graphs: dict = {}

    @property
    def num_rows(self) -> int:
        return self.row_bounds[-1] if self.row_bounds else 0

    @property
    def num_codes(self) -> int:
        return 0 if self.table is None else int(self.table.shape[0])

    def build(self, x_local: torch.Tensor) -> None:
        """Collective: every rank passes the descriptor rows it owns."""
        self.x_local = x_local
        self._graphs = {}
        codes_local = self.ops.hash(x_local)
        codes_all, sizes = _all_gather_rows(codes_local, self.group)
        self.row_bounds = [0]
        for s in sizes:
            self.row_bounds.append(self.row_bounds[-1] + s)
        self.table, _, self.csr_off, self.csr_rows, self.max_rows_per_code = codeops.build_table(codes_all, with_max=True)
        cuts = partition_bounds(int(self.table.shape[0]), self.world)
        self.scan_lo, self.scan_hi = cuts[self.rank], cuts[self.rank + 1]
        self.peers, self.peer_error = None, None
        if self.rerank_mode != "allreduce" and x_local.is_cuda and self.world > 1 and hasattr(self.ops, "rerank_peer"):
            # map every peer's shard into this process (CUDA IPC); all ranks must agree on the outcome
            try:
                from . import peer
                self.peers = peer.share_rows(x_local, self.row_bounds, self.group)
            except Exception as e:                               # no P2P path between the devices, IPC refused, ...
                self.peer_error = "%s: %s" % (type(e).__name__, str(e)[:200])
            ok = torch.tensor([1 if self.peers is not None else 0], dtype=torch.int32, device=x_local.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self.peers = None
                if self.rerank_mode == "peer":
                    raise RuntimeError("rerank='peer' needs CUDA IPC + peer access between all ranks' GPUs (%s)"
                                       % (self.peer_error or "a peer rank failed"))

    def close(self) -> None:
        """Drop the captured CUDA graphs (they hold NCCL work) and the peer mappings.  Call on every rank
        before ``destroy_process_group``."""
        self._graphs = {}
        if self.peers is not None:
            self.peers.release()
            self.peers = None

    # ------------------------------------------------------------------ partitions
    def _by_queries(self, Q: int) -> bool:
        return self.world > 1 and (self.scan_partition == "queries" or (
            self.scan_partition == "auto" and Q >= self.MIN_QUERIES_PER_RANK * self.world))

    def _slice(self, Q: int) -> Tuple[int, int, int]:
        per = (Q + self.world - 1) // self.world                 # equal slices (the last ones padded) for the collectives
        lo = min(Q, self.rank * per)
        return per, lo, min(Q, lo + per)

    @staticmethod
    def _padded(t: torch.Tensor, lo: int, hi: int, per: int, fill=0) -> torch.Tensor:
        mine = t[lo:hi]
        if hi - lo < per:
            pad = torch.full((per - (hi - lo),) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
            mine = torch.cat([mine, pad], dim=0)
        return mine.contiguous()

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)      # rank-major concatenation
        return out

    def near_codes(self, q_codes: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global n nearest unique codes of ALL queries on every rank: local scan, all-gather, merge."""
        Q = q_codes.shape[0]
        if self._by_queries(Q):
            per, lo, hi = self._slice(Q)
            with _stage("hamming_scan"):
                keys = self.ops.scan_keys(self.table, self._padded(q_codes, lo, hi, per), n, 0).contiguous()
            with _stage("allgather_merge"):
                return self.ops.merge_keys(self._gather(keys)[:Q].contiguous().unsqueeze(0))   # one part: decode only
        with _stage("hamming_scan"):
            keys = self.ops.scan_keys(self.table[self.scan_lo:self.scan_hi], q_codes, n, self.scan_lo).contiguous()
        with _stage("allgather_merge"):
            return self.ops.merge_keys(self._gather(keys).view((self.world, Q) + tuple(keys.shape[1:])))

    # ------------------------------------------------------------------ queries
    def query(self, q: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Collective: all ranks pass the SAME queries; all get the full result
        (global rows int64[Q, n], dists f64[Q, n])."""
        Q = q.shape[0]
        if self.peers is not None and self._by_queries(Q):
            # every rank needs only ITS slice of the batch on the device (hash, scan, re-rank are per slice)
            per, lo, hi = self._slice(Q)
            return self._unpack(self._run_slice(self._padded(q, lo, hi, per), Q, n), Q, n)
        from . import engine
        if not (self.graph and q.is_cuda and self.peers is not None) or engine.STAGE_EVENTS is not None:
            return self._query(q, n)
        key = ("full", int(Q), int(q.shape[1]), int(n), q.dtype, self.scan_partition)
        ent = self._graphs.get(key)
        if isinstance(ent, engine.GraphedCall):
            return ent(q)
        seen = (ent or 0) + 1
        if seen >= self.GRAPH_AFTER and self._fixed_pitch(Q, n):
            engine.evict_graphs(self._graphs, engine.DeviceLshIndex.MAX_GRAPHS - 1)
            self._graphs[key] = g = engine.GraphedCall(lambda qq: self._query(qq, n), q.contiguous())
            return g(q)
        self._graphs[key] = seen
        return self._query(q, n)

    def query_host(self, q_host: torch.Tensor, n: int):
        """``query`` for a HOST batch (pinned memory makes the copy asynchronous): with the scan split over the
        queries only this rank's slice crosses PCIe (Q / N rows instead of Q), and the result comes back in ONE
        device-to-host copy.  Returns numpy ``(rows int64[Q, n], dists float64[Q, n])``."""
        Q = q_host.shape[0]
        dev = self.x_local.device
        if self.peers is not None and self._by_queries(Q):
            from . import engine
            per, lo, hi = self._slice(Q)
            g = self._graphs.get(("slice", int(per), int(q_host.shape[1]), int(n), torch.float32))
            if self.graph and isinstance(g, engine.GraphedCall) and engine.STAGE_EVENTS is None and hi - lo == per:
                # steady state: host slice -> the graph's static input, replay, static output -> host (no allocation,
                # no intermediate copies)
                g.static_in[0].copy_(q_host[lo:hi], non_blocking=True)
                g.graph.replay()
                host = g.static_out[0][:Q].cpu()
            else:
                mine = torch.zeros((per, q_host.shape[1]), dtype=torch.float32, device=dev)
                if hi > lo:
                    mine[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
                host = self._run_slice(mine, Q, n).cpu()          # one copy, one synchronisation
            return host[:, :n].numpy(), host[:, n:].contiguous().view(torch.float64).numpy()
        rows, d = self.query(q_host.to(dev, non_blocking=True), n)
        return rows.cpu().numpy(), d.cpu().numpy()

    @staticmethod
    def _unpack(packed: torch.Tensor, Q: int, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return packed[:, :n].contiguous(), packed[:, n:].contiguous().view(torch.float64)

    def _run_slice(self, q_mine: torch.Tensor, Q: int, n: int) -> torch.Tensor:
        """This rank's query slice through the pipeline (graph replay once the shape repeats) ->
        the all-gathered packed result int64[Q, 2n] = (rows | distance bits)."""
        from . import engine
        per = q_mine.shape[0]
        if not self.graph or engine.STAGE_EVENTS is not None or not self._fixed_pitch(per, n):
            return self._slice_pipeline(q_mine, n)[:Q]
        key = ("slice", int(per), int(q_mine.shape[1]), int(n), q_mine.dtype)
        ent = self._graphs.get(key)
        if isinstance(ent, engine.GraphedCall):
            return ent(q_mine)[0][:Q]
        seen = (ent or 0) + 1
        if seen >= self.GRAPH_AFTER:
            engine.evict_graphs(self._graphs, engine.DeviceLshIndex.MAX_GRAPHS - 1)
            self._graphs[key] = g = engine.GraphedCall(lambda qq: (self._slice_pipeline(qq, n),), q_mine.contiguous())
            return g(q_mine)[0][:Q]
        self._graphs[key] = seen
        return self._slice_pipeline(q_mine, n)[:Q]

    def _fixed_pitch(self, Q: int, n: int) -> bool:
        from .engine import FIXED_PITCH_LIMIT
        return hasattr(self.ops, "expand") and n <= 2048 and Q * n * max(self.max_rows_per_code, 1) <= FIXED_PITCH_LIMIT

    def _query(self, q: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.peers is not None:
            return self._query_peer(q, n)
        with _stage("itq_hash"):
            q_codes = self.ops.hash(q)
        return self._query_allreduce(q, q_codes, n)

    def _slice_pipeline(self, q_mine: torch.Tensor, n: int) -> torch.Tensor:
        """Scan split over the queries: hash, scan of the whole table, expansion, peer re-rank and selection for
        this rank's slice, then ONE all-gather of the packed (row | distance bits) results of all slices."""
        with _stage("itq_hash"):
            codes_mine = self.ops.hash(q_mine)
        with _stage("hamming_scan"):
            keys = self.ops.scan_keys(self.table, codes_mine, n, 0).contiguous()
            _, code_rows = self.ops.merge_keys(keys.unsqueeze(0))                 # one part: decode only
        return self._finish_slice(q_mine, code_rows, n)

    def _finish_slice(self, q_mine: torch.Tensor, code_rows: torch.Tensor, n: int) -> torch.Tensor:
        per = q_mine.shape[0]
        pitch = n * max(self.max_rows_per_code, 1)
        with _stage("expand"):
            if self._fixed_pitch(per, n):                      # device-only, fixed-pitch segments padded with -1
                cand_idx, cand_off, cand_cnt = self.ops.expand(code_rows, self.csr_off, self.csr_rows, pitch)
            else:
                cand_idx, cand_off = expand_candidates(code_rows, self.csr_off, self.csr_rows)
                cand_cnt = None
        with _stage("rerank"):
            d = self.ops.rerank_peer(self.peers, q_mine, cand_idx, cand_off, pitch if cand_cnt is not None else 0)
            if cand_cnt is not None:
                rows, od = self.ops.rerank_select_rows(d, cand_off, cand_cnt, cand_idx, n)
            else:
                pos, od = self.ops.rerank_select(d, cand_off, n)
                rows = torch.where(pos >= 0, cand_idx[pos.clamp(min=0)], pos) if cand_idx.numel() else pos
        with _stage("allgather_results"):
            return self._gather(torch.cat([rows, od.view(torch.int64)], dim=1))     # (row | distance bits)[N * per, 2n]

    def _query_peer(self, q: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every rank answers ITS slice of the batch; one all-gather assembles the batch."""
        Q = q.shape[0]
        per, lo, hi = self._slice(Q)
        q_mine = self._padded(q, lo, hi, per)
        if self._by_queries(Q):
            return self._unpack(self._slice_pipeline(q_mine, n)[:Q], Q, n)
        with _stage("itq_hash"):
            q_codes = self.ops.hash(q)
        _, code_rows_all = self.near_codes(q_codes, n)
        code_rows = self._padded(code_rows_all, lo, hi, per, fill=-1)
        return self._unpack(self._finish_slice(q_mine, code_rows, n)[:Q], Q, n)

    def _query_allreduce(self, q: torch.Tensor, q_codes: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        _, code_rows = self.near_codes(q_codes, n)
        pitch = n * max(self.max_rows_per_code, 1)
        fixed = self._fixed_pitch(q.shape[0], n)
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
