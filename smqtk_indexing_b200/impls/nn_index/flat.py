"""
B200 exact flat (brute-force) L2 nearest-neighbour index.

Drop-in for the reference's flat configuration of its FAISS wrapper --
``FaissNearestNeighborsIndex(factory_string='IDMap,Flat', metric_type='l2')``
(reference: smqtk_indexing/impls/nn_index/faiss.py:486-559 build, :561-679
update / remove, :751-831 nn) -- without FAISS: the k rows with the smallest
euclidean distance, distances as in ``utils/metrics.py:73-86``, computed by
``sb_l2_topk`` (tensor-core filter + exact re-rank, csrc/flat_l2.cu).  Results are
in (distance, insertion row) order; unlike the reference (faiss.py:826-831 sorts the
uuids and distances but returns the descriptors unsorted) descriptors, uuids and
distances stay aligned.

Same plugin surface as every ``NearestNeighborsIndex``: ``build_index`` /
``update_index`` / ``remove_from_index`` / ``nn`` / ``count``, ``get_config`` /
``from_config``; plus the batch entry points ``build_index_matrix`` / ``nn_batch``.
"""
import logging
import threading
from typing import Any, Dict, Hashable, Iterable, List, Optional, Sequence, Tuple, Type, TypeVar

import numpy

from smqtk_core.configuration import from_config_dict, make_default_config, to_config_dict
from smqtk_core.dict import merge_dict
from smqtk_dataprovider.exceptions import ReadOnlyError
from smqtk_descriptors import DescriptorElement, DescriptorSet

from smqtk_indexing_b200.interfaces import NearestNeighborsIndex

LOG = logging.getLogger(__name__)
T_FLAT = TypeVar("T_FLAT", bound="FlatL2NearestNeighborsIndex")


class FlatL2NearestNeighborsIndex(NearestNeighborsIndex):
    """Exact L2 k-nearest-neighbour index over a device-resident descriptor matrix."""

    @classmethod
    def is_usable(cls) -> bool:
        # library built and a CUDA device visible (no CPU implementation exists to fall back to)
        from smqtk_indexing_b200 import _lib
        return _lib.usable()

    @classmethod
    def get_default_config(cls) -> Dict[str, Any]:
        default = super(FlatL2NearestNeighborsIndex, cls).get_default_config()
        default['descriptor_set'] = make_default_config(DescriptorSet.get_impls())
        return default

    @classmethod
    def from_config(cls: Type[T_FLAT], config_dict: Dict, merge_default: bool = True) -> T_FLAT:
        if merge_default:
            cfg = cls.get_default_config()
            merge_dict(cfg, config_dict)
        else:
            cfg = config_dict
        cfg['descriptor_set'] = from_config_dict(cfg['descriptor_set'], DescriptorSet.get_impls())
        return super(FlatL2NearestNeighborsIndex, cls).from_config(cfg, False)

    def __init__(self, descriptor_set: DescriptorSet, read_only: bool = False):
        super(FlatL2NearestNeighborsIndex, self).__init__()
        self._descriptor_set = descriptor_set
        self.read_only = read_only
        self._model_lock = threading.RLock()
        self._x = None                               # float32[N, D] on the device
        self._prepared = None                        # (|row|^2, max) for sb_l2_topk
        self._uuids: List[Hashable] = []             # row -> uuid
        self._row_of: Optional[Dict[Hashable, int]] = {}   # uuid -> row; None = not built yet (matrix-built index)
        self._adopted = False                        # self._x is the CALLER's tensor: copy before writing into it

    def get_config(self) -> Dict[str, Any]:
        return {
            "descriptor_set": to_config_dict(self._descriptor_set),
            "read_only": self.read_only,
        }

    def count(self) -> int:
        with self._model_lock:
            return len(self._uuids)

    # ------------------------------------------------------------------ device table
    def _set_matrix(self, x, uuids: Sequence[Hashable], adopted: bool = False) -> None:
        from smqtk_indexing_b200 import device
        import torch
        adopted = adopted and isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32))
        if not x.is_cuda:
            x = x.to(device.device())
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        self._x = x.contiguous()
        self._uuids = list(uuids) if not isinstance(uuids, range) else uuids   # type: ignore
        # default (range) uuids: the uuid -> row lookup is built on first use (10M+ rows: not up front)
        self._row_of = None if isinstance(uuids, range) else {u: i for i, u in enumerate(self._uuids)}
        self._adopted = adopted
        self._prepared = device.l2_prepare(self._x) if len(self._x) else None

    def _lookup(self) -> Dict[Hashable, int]:
        if self._row_of is None:
            self._row_of = {u: i for i, u in enumerate(self._uuids)}
        return self._row_of

    @staticmethod
    def _vectors(descriptors: Sequence[DescriptorElement]) -> numpy.ndarray:
        return numpy.asarray([d.vector() for d in descriptors], dtype=numpy.float32)

    # ------------------------------------------------------------------ build / update / remove
    def _build_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        if self.read_only:
            raise ReadOnlyError("Cannot modify read-only index.")
        desc_list = list(descriptors)
        with self._model_lock:
            self._descriptor_set.clear()
            self._descriptor_set.add_many_descriptors(desc_list)
            # a uuid given twice keeps its last vector, like the descriptor set
            last = {d.uuid(): d for d in desc_list}
            elems = list(last.values())
            self._set_matrix(self._vectors(elems), [d.uuid() for d in elems])

    def _update_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        import torch
        if self.read_only:
            raise ReadOnlyError("Cannot modify read-only index.")
        new = list(descriptors)
        with self._model_lock:
            self._descriptor_set.add_many_descriptors(new)
            last = {d.uuid(): d for d in new}
            if self._x is None or len(self._uuids) == 0:
                elems = list(last.values())
                self._set_matrix(self._vectors(elems), [d.uuid() for d in elems])
                return
            vec = torch.from_numpy(self._vectors(list(last.values()))).to(self._x.device)
            x = self._x
            uuids = list(self._uuids)
            row_of = self._lookup()
            append = []
            for i, u in enumerate(last):
                r = row_of.get(u)
                if r is None:
                    append.append(i)
                    uuids.append(u)
                else:
                    if self._adopted:                  # the matrix belongs to the caller: copy before the first write
                        x = x.clone()
                        self._adopted = False
                    x[r] = vec[i]                      # re-added uuid: overwrite in place
            if append:
                sel = torch.tensor(append, dtype=torch.int64, device=x.device)
                x = torch.cat([x, vec[sel]], dim=0)
            self._set_matrix(x, uuids)

    def _remove_from_index(self, uids: Iterable[Hashable]) -> None:
        import torch
        if self.read_only:
            raise ReadOnlyError("Cannot modify read-only index.")
        with self._model_lock:
            uids = list(uids)
            row_of = self._lookup()
            for uid in uids:                           # KeyError before any mutation (faiss.py:661-665)
                if uid not in row_of:
                    raise KeyError(uid)
            gone = {row_of[u] for u in uids}
            keep = [r for r in range(len(self._uuids)) if r not in gone]
            if self._descriptor_set.count():           # matrix-built indexes never filled the descriptor set
                self._descriptor_set.remove_many_descriptors([u for u in uids if self._descriptor_set.has_descriptor(u)])
            sel = torch.tensor(keep, dtype=torch.int64, device=self._x.device)
            self._set_matrix(self._x[sel], [self._uuids[r] for r in keep])

    def build_index_matrix(self, x, uuids: Optional[Sequence[Hashable]] = None) -> None:
        """Bulk build from a ``[N, D]`` matrix (numpy or float32 CUDA tensor, adopted without
        a copy -- the index never writes into it: a later ``update_index`` that overwrites a row
        copies first); ``uuids[r]`` names row ``r`` (default: the row number).  The
        ``descriptor_set`` collaborator is not populated: use ``nn_batch``."""
        if self.read_only:
            raise ReadOnlyError("Cannot modify read-only index.")
        if x is None or len(x) == 0:
            raise self._empty_iterable_exception()
        with self._model_lock:
            self._set_matrix(x, uuids if uuids is not None else range(len(x)), adopted=True)

    # ------------------------------------------------------------------ queries
    def nn_batch(self, queries, n: int = 1, return_device: bool = False):
        """``(rows int64[Q, n], dists float64[Q, n])`` in (distance, row) order; rows index
        :meth:`row_uuids`, -1 / NaN pad when the index holds fewer than ``n`` rows."""
        import torch
        from smqtk_indexing_b200 import device
        with self._model_lock:
            if self._x is None or len(self._x) == 0:
                raise ValueError("No index currently set to query from!")
            if isinstance(queries, torch.Tensor):
                q = queries.to(self._x.device, torch.float32)
            else:
                q = torch.from_numpy(numpy.ascontiguousarray(queries, dtype=numpy.float32)).to(self._x.device)
            if q.dim() == 1:
                q = q.unsqueeze(0)
            rows, dists = device.l2_topk(self._x, q.contiguous(), n, prepared=self._prepared)
        if return_device:
            return rows, dists
        return rows.cpu().numpy(), dists.cpu().numpy()

    def row_uuids(self) -> Sequence[Hashable]:
        """row -> uuid for the rows returned by :meth:`nn_batch`."""
        return self._uuids

    def _nn(self, d: DescriptorElement, n: int = 1) -> Tuple[Tuple[DescriptorElement, ...], Tuple[float, ...]]:
        rows, dists = self.nn_batch(numpy.asarray(d.vector(), dtype=numpy.float32)[None, :], n)
        keep = rows[0] >= 0
        with self._model_lock:
            uuids = [self._uuids[r] for r in rows[0][keep]]
            descriptors = tuple(self._descriptor_set.get_many_descriptors(uuids))
        return descriptors, tuple(float(v) for v in dists[0][keep])
