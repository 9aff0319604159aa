"""
Peer memory over NVLink / NVSwitch for the row-sharded descriptor table.

One process per GPU (``torch.distributed``); every rank owns a contiguous range of descriptor rows in
its own HBM.  ``share_rows`` maps every peer's shard into the calling process with CUDA IPC
(``cudaIpcGetMemHandle`` / ``cudaIpcOpenMemHandle`` through torch's storage sharing, the mechanism
``torch.multiprocessing`` uses) and enables peer access, so that a kernel on this GPU can LOAD a
candidate row straight from the GPU that holds it: the re-rank stage (reference lsh.py:500-519)
then needs no collective at all -- ~10 rows x 2 KB per query cross NVLink, nothing else.

Only plumbing lives here (handle exchange, pointer table); the loads are issued by
``rerank_kernel`` (csrc/rerank.cu, ``sb_rerank_peer``).
"""
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


class PeerShards:
    """Pointer table of a row-sharded float32 matrix: shard r = global rows [bounds[r], bounds[r+1])."""

    def __init__(self, ptrs: List[int], bounds: List[int], dim: int, ld: int, device: torch.device, keep) -> None:
        self.ptrs = ptrs
        self.bounds = bounds
        self.dim = dim
        self.ld = ld
        self.keep = keep                               # mapped storages / the local tensor: alive as long as the table
        self.ptr_table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.bound_table = torch.tensor(bounds, dtype=torch.int64, device=device)
        self.aligned16 = all(p % 16 == 0 for p in ptrs) and ld % 4 == 0

    @property
    def n_shards(self) -> int:
        return len(self.ptrs)


def share_rows(x_local: torch.Tensor, bounds: List[int], group=None) -> PeerShards:
    """Collective: every rank passes its shard (float32 CUDA, contiguous rows); returns the pointer
    table of all shards as seen from THIS process.  Raises when the devices cannot reach each other
    (the caller then falls back to the collective re-rank)."""
    if not (x_local.is_cuda and x_local.dtype == torch.float32 and x_local.dim() == 2 and x_local.stride(1) == 1):
        raise ValueError("x_local must be a float32 CUDA matrix with unit column stride")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = x_local.device
    lib = _lib.load()
    meta = x_local.untyped_storage()._share_cuda_()
    info = (meta, x_local.storage_offset() * x_local.element_size(), int(x_local.stride(0)) if x_local.shape[0] > 1 else int(x_local.shape[1]),
            int(x_local.shape[1]))
    infos: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(infos, info, group=group)
    lds = {i[2] for i in infos}
    dims = {i[3] for i in infos}
    if len(lds) != 1 or len(dims) != 1:
        raise ValueError("descriptor shards differ in width or row stride: %s / %s" % (sorted(dims), sorted(lds)))
    ptrs, keep = [], [x_local]
    with torch.cuda.device(dev):
        for r, (m, off, _ld, _d) in enumerate(infos):
            if r == rank:
                ptrs.append(x_local.data_ptr())
                continue
            src_dev = int(m[0])
            storage = torch.UntypedStorage._new_shared_cuda(*m)
            keep.append(storage)
            ptrs.append(storage.data_ptr() + off)
            if src_dev != dev.index:
                _lib.check(lib.sb_enable_peer_access(src_dev))
    return PeerShards(ptrs, list(bounds), dims.pop(), lds.pop(), dev, keep)
