"""
Peer memory over NVLink / NVSwitch for the row-sharded descriptor table.

One process per GPU (``torch.distributed``); every rank owns a contiguous range of descriptor rows in
its own HBM.  ``share_rows`` maps every peer's shard into the calling process with CUDA IPC
(``sb_ipc_export`` / ``sb_ipc_import``: ``cudaIpcGetMemHandle`` on the exporter,
``cudaIpcOpenMemHandle(..., cudaIpcMemLazyEnablePeerAccess)`` on every importer WITH ITS OWN DEVICE
CURRENT -- that is what makes the mapping loadable from the importer's kernels), so that a kernel on
this GPU can LOAD a candidate row straight from the GPU that holds it: the re-rank stage (reference
lsh.py:500-519) then needs no collective at all -- ~10 rows x 2 KB per query cross NVLink, nothing else.

Only plumbing lives here (handle exchange, pointer table); the loads are issued by
``rerank_kernel`` (csrc/rerank.cu, ``sb_rerank_peer``).
"""
import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


class PeerShards:
    """Pointer table of a row-sharded float32 matrix: shard r = global rows [bounds[r], bounds[r+1])."""

    def __init__(self, ptrs: List[int], bounds: List[int], dim: int, ld: int, device: torch.device, keep, bases) -> None:
        self.ptrs = ptrs
        self.bounds = bounds
        self.dim = dim
        self.ld = ld
        self.keep = keep                               # the local tensor: alive as long as the table
        self._bases = bases                            # imported allocation bases (closed on release)
        self.device = device
        self.ptr_table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.bound_table = torch.tensor(bounds, dtype=torch.int64, device=device)
        self.aligned16 = all(p % 16 == 0 for p in ptrs) and ld % 4 == 0

    @property
    def n_shards(self) -> int:
        return len(self.ptrs)

    def release(self) -> None:
        """Close the imported mappings (the peers' memory itself is untouched)."""
        bases, self._bases = self._bases, []
        if bases:
            lib = _lib.load()
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                for b in bases:
                    lib.sb_ipc_release(ctypes.c_void_p(b))

    # no __del__: at interpreter teardown the CUDA context / the peer process may already be gone, and the
    # driver unmaps everything when the process exits; call release() to drop the mappings earlier


def share_rows(x_local: torch.Tensor, bounds: List[int], group=None) -> PeerShards:
    """Collective: every rank passes its shard (float32 CUDA, contiguous rows); returns the pointer
    table of all shards as seen from THIS process.  Raises when the devices cannot reach each other
    (the caller then falls back to the collective re-rank)."""
    if not (x_local.is_cuda and x_local.dtype == torch.float32 and x_local.dim() == 2 and x_local.stride(1) == 1):
        raise ValueError("x_local must be a float32 CUDA matrix with unit column stride")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = x_local.device
    lib = _lib.load()
    handle = ctypes.create_string_buffer(64)
    offset = ctypes.c_int64(0)
    err: Optional[str] = None
    with torch.cuda.device(dev):
        try:
            _lib.check(lib.sb_ipc_export(ctypes.c_void_p(x_local.data_ptr()), ctypes.cast(handle, ctypes.c_void_p),
                                         ctypes.byref(offset)))
        except RuntimeError as e:
            err = str(e)
    ld = int(x_local.stride(0)) if x_local.shape[0] > 1 else int(x_local.shape[1])
    info = (bytes(handle.raw), int(offset.value), ld, int(x_local.shape[1]), int(dev.index), err)
    infos: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(infos, info, group=group)
    errs = [i[5] for i in infos if i[5]]
    if errs:
        raise RuntimeError("a rank could not export its shard: %s" % errs[0])
    lds = {i[2] for i in infos}
    dims = {i[3] for i in infos}
    if len(lds) != 1 or len(dims) != 1:
        raise ValueError("descriptor shards differ in width or row stride: %s / %s" % (sorted(dims), sorted(lds)))
    ptrs, bases = [], []
    with torch.cuda.device(dev):
        for r, (h, off, _ld, _d, _src, _e) in enumerate(infos):
            if r == rank:
                ptrs.append(x_local.data_ptr())
                continue
            base = ctypes.c_void_p(0)
            hbuf = ctypes.create_string_buffer(h, 64)
            _lib.check(lib.sb_ipc_import(ctypes.cast(hbuf, ctypes.c_void_p), ctypes.byref(base)))
            bases.append(int(base.value))
            ptrs.append(int(base.value) + off)
    return PeerShards(ptrs, list(bounds), dims.pop(), lds.pop(), dev, [x_local], bases)
