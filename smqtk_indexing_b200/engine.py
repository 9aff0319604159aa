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


class GraphedCall:
    """A query pipeline captured once into a CUDA graph and replayed for every later batch of the same
    shape.  The pipeline is ~45 short kernel launches (hash, 9 scan chunks x 3-4 kernels, expand,
    re-rank, select [, one NCCL all-gather]); replaying it as one graph removes the per-launch host
    cost, which at 8 GPUs (1/8 of the tensor work per rank) is otherwise a third of the step.
    ``fn(static_inputs...) -> tuple of tensors`` must be free of host synchronisation -- the
    pipelines here are (an overflowed tensor-core scan is redone by a predicated kernel, not by the host).
    Outputs are cloned out of the graph's static buffers on every call."""

    def __init__(self, fn, *example_inputs: torch.Tensor) -> None:
        self.static_in = [t.clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                          # warm-up off the capture: lazy inits, allocator growth
            for _ in range(2):
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = fn(*self.static_in)
        self.static_out = tuple(out)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)                      # (host sources: pinned memory makes this asynchronous)
        self.graph.replay()
        return tuple(o.clone() for o in self.static_out)

    def run_to_host(self, *inputs: torch.Tensor):
        """Same, returning HOST copies straight from the static outputs (no device-side clone)."""
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return tuple(o.cpu() for o in self.static_out)


def evict_graphs(graphs: dict, keep: int) -> None:
    """Drop the oldest captured graphs (and stale shape counters) so that at most ``keep`` graphs remain."""
    captured = [k for k, v in graphs.items() if isinstance(v, GraphedCall)]
    for k in captured[:max(0, len(captured) - keep)]:
        del graphs[k]
    if len(graphs) > 64:                                   # shape counters of one-off batch sizes
        for k in [k for k, v in graphs.items() if not isinstance(v, GraphedCall)][:len(graphs) - 64]:
            del graphs[k]


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
    """Descriptor rows + codes + unique-code table + CSR, all on one device.

    Ingest is incremental (reference lsh.py:331-450): rows are appended into
    buffers that grow geometrically, removed rows become tombstones (their row numbers stay
    valid for the layers that map rows to uuids) and ``reindex`` rebuilds the unique table and
    the CSR over the live rows on the device (``sb_unique_codes``)."""

    #: buffers grow by this factor when an append does not fit
    GROWTH = 1.5

    def __init__(self) -> None:
        self.x: Optional[torch.Tensor] = None          # float32[N, D]   (view of the first N buffer rows)
        self.codes: Optional[torch.Tensor] = None      # int32[N, W]
        self.alive: Optional[torch.Tensor] = None      # bool[N], None = every row is live
        self.table: Optional[torch.Tensor] = None      # int32[U, W] sorted unique (live rows only)
        self.row_code: Optional[torch.Tensor] = None   # int64[N] row -> table row (-1 = removed)
        self.csr_off: Optional[torch.Tensor] = None    # int64[U + 1]
        self.csr_rows: Optional[torch.Tensor] = None   # int64[live rows]
        self.max_rows_per_code: int = 0
        self._x_buf: Optional[torch.Tensor] = None
        self._codes_buf: Optional[torch.Tensor] = None
        self.num_dead: int = 0
        self._graphs: dict = {}                          # (Q, D, n, metric, dtype) -> GraphedCall | times seen

    def clear(self) -> None:
        self.__init__()

    @property
    def num_rows(self) -> int:
        """Physical rows (live + tombstones)."""
        return 0 if self.codes is None else int(self.codes.shape[0])

    @property
    def num_live(self) -> int:
        return self.num_rows - self.num_dead

    @property
    def num_codes(self) -> int:
        return 0 if self.table is None else int(self.table.shape[0])

    def set_rows(self, x: Optional[torch.Tensor], codes: torch.Tensor) -> None:
        """Adopt descriptor rows (may be None for a codes-only index) and their codes."""
        self.x, self.codes = x, codes
        self._x_buf, self._codes_buf = x, codes
        self.alive, self.num_dead = None, 0
        self.reindex()

    @staticmethod
    def _grown(buf: Optional[torch.Tensor], used: int, new: torch.Tensor, growth: float) -> torch.Tensor:
        """``buf`` (capacity rows) with ``new`` copied in at row ``used``; reallocated when too small."""
        need = used + new.shape[0]
        if buf is None or buf.shape[0] < need or buf.shape[1:] != new.shape[1:]:
            cap = max(need, int(used * growth) + 1)
            grown = torch.empty((cap,) + tuple(new.shape[1:]), dtype=new.dtype, device=new.device)
            if used:
                grown[:used] = buf[:used]
            buf = grown
        buf[used:need] = new
        return buf

    def append_rows(self, x: Optional[torch.Tensor], codes: torch.Tensor) -> int:
        """Append rows (no re-index: call ``reindex`` after the batch).  Returns the first new row."""
        n0 = self.num_rows
        if n0 and codes.shape[1] != self.codes.shape[1]:
            w = max(codes.shape[1], self.codes.shape[1])
            codes = codeops.widen(codes, w)
            if self.codes.shape[1] != w:
                self._codes_buf = codeops.widen(self.codes, w)
        self._codes_buf = self._grown(self._codes_buf, n0, codes, self.GROWTH)
        self.codes = self._codes_buf[:n0 + codes.shape[0]]
        if x is not None:
            self._x_buf = self._grown(self._x_buf, n0, x, self.GROWTH)
            self.x = self._x_buf[:n0 + x.shape[0]]
        if self.alive is not None:
            self.alive = torch.cat([self.alive, torch.ones(codes.shape[0], dtype=torch.bool, device=codes.device)])
        return n0

    def overwrite_rows(self, rows: torch.Tensor, x: Optional[torch.Tensor], codes: torch.Tensor) -> None:
        """Replace the vectors / codes of existing rows (a uuid added again)."""
        if codes.shape[1] != self.codes.shape[1]:
            codes = codeops.widen(codes, self.codes.shape[1])
        self.codes[rows] = codes
        if x is not None and self.x is not None:
            self.x[rows] = x

    def remove_rows(self, rows: torch.Tensor) -> None:
        """Tombstone rows (int64 tensor of live rows; no re-index)."""
        if self.alive is None:
            self.alive = torch.ones(self.num_rows, dtype=torch.bool, device=self.codes.device)
        self.alive[rows] = False
        self.num_dead = self.num_rows - int(self.alive.sum().item())

    def compact(self) -> Optional[torch.Tensor]:
        """Drop the tombstones physically.  Returns old-row -> new-row (int64, -1 = removed) or None
        when nothing was dead; the caller re-maps whatever it keys by row."""
        if self.alive is None or self.num_dead == 0:
            self.alive, self.num_dead = None, 0
            return None
        live = torch.nonzero(self.alive).reshape(-1)
        remap = torch.full((self.num_rows,), -1, dtype=torch.int64, device=live.device)
        remap[live] = torch.arange(live.numel(), device=live.device)
        x = self.x[live].contiguous() if self.x is not None else None
        self.set_rows(x, self.codes[live].contiguous())
        return remap

    # ------------------------------------------------------------------ snapshot
    def state_dict(self) -> dict:
        """Host copy of what cannot be recomputed cheaply: descriptor rows, their codes, tombstones.
        The unique table and the CSR are NOT stored -- ``load_state_dict`` re-derives them on the
        device (a sort of the codes), which also keeps the format independent of their layout."""
        return {
            "format": "smqtk_indexing_b200.DeviceLshIndex/1",
            "x": None if self.x is None else self.x.cpu(),
            "codes": None if self.codes is None else self.codes.cpu(),
            "alive": None if self.alive is None else self.alive.cpu(),
        }

    def load_state_dict(self, state: dict, dev) -> None:
        if state.get("format") != "smqtk_indexing_b200.DeviceLshIndex/1":
            raise ValueError("not a DeviceLshIndex snapshot: %r" % (state.get("format"),))
        self.clear()
        if state["codes"] is None:
            return
        x = None if state["x"] is None else state["x"].to(dev)
        self.x, self.codes = x, state["codes"].to(dev)
        self._x_buf, self._codes_buf = self.x, self.codes
        if state["alive"] is not None:
            self.alive = state["alive"].to(dev)
            self.num_dead = self.num_rows - int(self.alive.sum().item())
        self.reindex()

    def reindex(self) -> None:
        """Recompute the unique table and the CSR from the codes of the live rows."""
        self._graphs = {}                                # captured pipelines point at the old table / CSR
        if self.codes is None or self.num_live == 0:
            self.table = self.csr_off = self.csr_rows = self.row_code = None
            self.max_rows_per_code = 0
            return
        # sb_unique_codes: radix sort of a row permutation + boundaries + CSR in one device pipeline; the
        # one host read per re-index returns U and the most rows sharing one code (sizes the fixed-pitch
        # candidate segments)
        live = None if (self.alive is None or self.num_dead == 0) else torch.nonzero(self.alive).reshape(-1)
        self.table, self.row_code, self.csr_off, self.csr_rows, self.max_rows_per_code = codeops.build_table(
            self.codes, rows=live, with_max=True)

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
                dist = device.rerank(self.x, q, cand_idx, cand_off, distance_method, pitch=pitch)
                return device.rerank_select_rows(dist, cand_off, cand_cnt, cand_idx, n, max_m=pitch)
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

    #: capture the pipeline of a (Q, n) shape into a CUDA graph once it has been asked for this often
    GRAPH_AFTER = 2
    #: captured shapes kept per index (each holds its workspaces: ~0.3 GB at 4096 queries); oldest evicted
    MAX_GRAPHS = 4

    def query_graphed(self, functor, q: torch.Tensor, n: int, distance_method: str, to_host: bool = False):
        """``query`` through a per-shape CUDA graph (captured on the ``GRAPH_AFTER``-th batch of the
        shape; any re-index drops the graphs).  Shapes whose candidate expansion needs the host
        (heavy code collisions) and batches recorded with stage events stay eager."""
        key = (int(q.shape[0]), int(q.shape[1]), int(n), distance_method, q.dtype)
        pitch = n * max(self.max_rows_per_code, 1)
        eager = (STAGE_EVENTS is not None or n > 2048 or q.shape[0] * pitch > FIXED_PITCH_LIMIT
                 or torch.cuda.is_current_stream_capturing())
        if eager:
            if q.device != self.x.device:
                q = q.to(self.x.device, non_blocking=True)
            out = self.query(functor, q, n, distance_method)
            return tuple(o.cpu() for o in out) if to_host else out
        ent = self._graphs.get(key)
        if isinstance(ent, GraphedCall):
            return ent.run_to_host(q) if to_host else ent(q)      # q may live on the host: copied into the static input
        if q.device != self.x.device:
            q = q.to(self.x.device, non_blocking=True)
        seen = (ent or 0) + 1
        if seen >= self.GRAPH_AFTER:
            evict_graphs(self._graphs, self.MAX_GRAPHS - 1)
            self._graphs[key] = g = GraphedCall(lambda qq: self.query(functor, qq, n, distance_method), q)
            out = g(q)
        else:
            self._graphs[key] = seen
            out = self.query(functor, q, n, distance_method)
        return tuple(o.cpu() for o in out) if to_host else out

    def query(self, functor, q: torch.Tensor, n: int, distance_method: str):
        """hash -> Hamming top-n unique codes -> candidates -> re-rank -> top-n."""
        with _stage("itq_hash"):
            q_codes = functor.get_hash_packed(q)
        # no host round trip anywhere: an overflowed tensor-core scan is redone on the device by the
        # predicated XOR/POPC scan (device.hamming_scan_keys)
        _, code_rows = self.near_codes(q_codes, n)
        return self.rerank(q, code_rows, n, distance_method)
