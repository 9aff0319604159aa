"""
B200 ``LSHNearestNeighborIndex``: locality-sensitive-hashing nearest-neighbour
index whose arithmetic (hashing, Hamming scan / top-k, candidate re-rank) runs on
the GPU through ``libsmqtk_b200``.

Drop-in for ``smqtk_indexing.impls.nn_index.lsh.LSHNearestNeighborIndex``
(reference: smqtk_indexing/impls/nn_index/lsh.py:39-519): same constructor and
config keys, same collaborators (``DescriptorSet``, ``KeyValueStore``,
``HashIndex``, ``LshFunctor``), same state after build / update / remove, same
exceptions.  Two query entry points:

* ``nn(d, n)`` -- the reference API, one ``DescriptorElement`` per call; walks the
  collaborator containers exactly like lsh.py:470-519 and uses the GPU for the
  three arithmetic stages.
* ``nn_batch(X, n)`` -- the throughput API: a ``[Q, D]`` matrix of queries against
  the device-resident mirror of the index (descriptor matrix, packed codes, unique
  code table, code -> rows CSR); nothing leaves HBM between stages.

Ordering contract (the reference leaves ties to set-iteration order): near codes
come in (Hamming distance, code value) order, candidates in (code rank, row)
order, results in (distance, candidate position) order.
"""
import logging
import threading
from typing import Any, Callable, Dict, Hashable, Iterable, Iterator, List, Optional, Sequence, Set, Tuple, Type, TypeVar

import numpy

from smqtk_core.configuration import from_config_dict, make_default_config, to_config_dict
from smqtk_core.dict import merge_dict
from smqtk_dataprovider import KeyValueStore
try:                                     # the real package keeps the sentinel next to the interface
    from smqtk_dataprovider.interfaces.key_value_store import NO_DEFAULT_VALUE
except ImportError:                      # stand-in package
    from smqtk_dataprovider import NO_DEFAULT_VALUE
from smqtk_dataprovider.exceptions import ReadOnlyError
from smqtk_descriptors import DescriptorElement, DescriptorSet

from smqtk_indexing_b200.engine import DeviceLshIndex
from smqtk_indexing_b200.interfaces import HashIndex, LshFunctor, NearestNeighborsIndex
from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
from smqtk_indexing_b200.utils import bits as bitutil
from smqtk_indexing_b200.utils import metrics

LOG = logging.getLogger(__name__)
T_LSH = TypeVar("T_LSH", bound="LSHNearestNeighborIndex")


class _DeviceMirror(DeviceLshIndex):
    """Device index + the row <-> uuid maps of the plugin layer."""

    def __init__(self) -> None:
        super().__init__()
        self.uuids: List[Hashable] = []     # row -> uuid (a ``range`` for matrix-built indexes without uuids)
        self.row_of: Optional[Dict[Hashable, int]] = {}   # uuid -> row; None = not built yet (matrix-built index)
        #: fed through the ``*_matrix`` methods / a snapshot: the collaborator containers are not its source of truth
        self.matrix_built = False

    def count(self) -> int:
        return len(self.uuids)


class LSHNearestNeighborIndex(NearestNeighborsIndex):
    """LSH nearest-neighbour index (hash functor + unique-code Hamming index +
    descriptor re-rank)."""

    @classmethod
    def is_usable(cls) -> bool:
        # library built and a CUDA device visible (no CPU implementation exists to fall back to)
        from smqtk_indexing_b200 import _lib
        return _lib.usable()

    @classmethod
    def get_default_config(cls) -> Dict[str, Any]:
        default = super(LSHNearestNeighborIndex, cls).get_default_config()
        default['lsh_functor'] = make_default_config(LshFunctor.get_impls())
        default['descriptor_set'] = make_default_config(DescriptorSet.get_impls())
        default['hash2uuids_kvstore'] = make_default_config(KeyValueStore.get_impls())
        default['hash_index'] = make_default_config(HashIndex.get_impls())
        return default

    @classmethod
    def from_config(cls: Type[T_LSH], config_dict: Dict, merge_default: bool = True) -> T_LSH:
        if merge_default:
            cfg = cls.get_default_config()
            merge_dict(cfg, config_dict)
        else:
            cfg = config_dict
        cfg['lsh_functor'] = from_config_dict(cfg['lsh_functor'], LshFunctor.get_impls())
        cfg['descriptor_set'] = from_config_dict(cfg['descriptor_set'], DescriptorSet.get_impls())
        cfg['hash2uuids_kvstore'] = from_config_dict(cfg['hash2uuids_kvstore'], KeyValueStore.get_impls())
        if cfg['hash_index'] and cfg['hash_index']['type']:
            cfg['hash_index'] = from_config_dict(cfg['hash_index'], HashIndex.get_impls())
        else:
            cfg['hash_index'] = None
        if 'hash_index_comment' in cfg:
            del cfg['hash_index_comment']
        return super(LSHNearestNeighborIndex, cls).from_config(cfg, False)

    def __init__(
        self,
        lsh_functor: LshFunctor,
        descriptor_set: DescriptorSet,
        hash2uuids_kvstore: KeyValueStore,
        hash_index: Optional[HashIndex] = None,
        distance_method: str = 'cosine',
        read_only: bool = False
    ):
        super(LSHNearestNeighborIndex, self).__init__()
        self.lsh_functor = lsh_functor
        self.descriptor_set = descriptor_set
        self.hash_index = hash_index
        self.hash2uuids_kvstore = hash2uuids_kvstore
        self.distance_method = distance_method
        self.read_only = read_only
        # In-process lock (the reference uses a multiprocessing.RLock because its
        # containers may live in other processes; device state cannot).
        self._model_lock = threading.RLock()
        self._distance_function = self._get_dist_func(self.distance_method)
        self._mirror = _DeviceMirror()

    @staticmethod
    def _get_dist_func(distance_method: str) -> Callable[[numpy.ndarray, numpy.ndarray], float]:
        """Distance function for a method label (reference lsh.py:236-255).

        :raises ValueError: unknown label."""
        if distance_method == "euclidean":
            return metrics.euclidean_distance
        elif distance_method == "cosine":
            return metrics.cosine_distance
        elif distance_method == 'hik':
            return metrics.histogram_intersection_distance_fast
        else:
            raise ValueError("Invalid distance method label. Must be one of "
                             "['euclidean' | 'cosine' | 'hik']")

    def get_config(self) -> Dict[str, Any]:
        hi_conf = None
        if self.hash_index is not None:
            hi_conf = to_config_dict(self.hash_index)
        return {
            "lsh_functor": to_config_dict(self.lsh_functor),
            "descriptor_set": to_config_dict(self.descriptor_set),
            "hash_index": hi_conf,
            "hash2uuids_kvstore": to_config_dict(self.hash2uuids_kvstore),
            "distance_method": self.distance_method,
            "read_only": self.read_only,
        }

    def count(self) -> int:
        """Number of descriptors reachable through the hash -> uuids map
        (reference lsh.py:271-281: sums the KVS value sets)."""
        with self._model_lock:
            c = 0
            for set_v in self.hash2uuids_kvstore.values():
                c += len(set_v)
            return c

    # ------------------------------------------------------------------ hashing helpers
    def _hash_matrix(self, x: numpy.ndarray):
        """[n, D] host matrix -> (packed codes on device int32[n, W], bit length b)."""
        f = self.lsh_functor
        if hasattr(f, "get_hash_packed"):
            codes = f.get_hash_packed(x)
            rot = numpy.asarray(f.rotation)
            b = rot.reshape(rot.shape[0], -1).shape[1]
            return codes, b
        # foreign functor: one host call per vector, pack + upload
        from smqtk_indexing_b200 import device
        hv = [numpy.asarray(f.get_hash(v)).astype(bool).ravel() for v in x]
        b = max(max(len(h) for h in hv), 1)          # lengths may differ (value semantics, like the reference's ints)
        ints = [bitutil.bit_vector_to_int_large(h) for h in hv]
        return device.codes_to_device(bitutil.ints_to_words(ints, bitutil.words_for_bits(b))), b

    @staticmethod
    def _vectors(descriptors: Sequence[DescriptorElement]) -> numpy.ndarray:
        return numpy.asarray([d.vector() for d in descriptors])

    def _push_hashes(self, op: str, codes_host: numpy.ndarray, bits: int, table=None) -> None:
        """Forward hash codes to the configured ``hash_index`` (reference
        lsh.py:326-329, 381-383, 444-447)."""
        hi = self.hash_index
        if hi is None:
            return
        if isinstance(hi, LinearHashIndex) and op == "build" and table is not None:
            hi.set_code_table(table)
            hi.save_cache()
            return
        vectors = bitutil.unpack_bits(codes_host, bits)
        getattr(hi, {"build": "build_index", "update": "update_index", "remove": "remove_from_index"}[op])(vectors)

    # ------------------------------------------------------------------ build / update / remove
    def _build_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        from smqtk_indexing_b200 import device
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            LOG.debug("Clearing and adding new descriptor elements")
            self.descriptor_set.clear()
            self.descriptor_set.add_many_descriptors(descriptors)

            LOG.debug("Generating hash codes")
            self.hash2uuids_kvstore.clear()
            elems = list(self.descriptor_set)
            self._mirror.clear()
            if not elems:
                return
            x = self._vectors(elems)
            codes, bits = self._hash_matrix(x)
            m = self._mirror
            import torch
            m.uuids = [d.uuid() for d in elems]
            m.row_of = {u: i for i, u in enumerate(m.uuids)}
            m.set_rows(torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32)).to(codes.device), codes)

            # hash -> uuids map, one add_many (reference lsh.py:313-324)
            table_ints = bitutil.words_to_ints(device.codes_to_host(m.table))
            row_code = m.row_code.cpu().numpy()
            kvstore_update: Dict[Hashable, Set[Hashable]] = {h: set() for h in table_ints}
            for r, c in enumerate(row_code):
                kvstore_update[table_ints[c]].add(m.uuids[r])
            self.hash2uuids_kvstore.add_many(kvstore_update)

            self._push_hashes("build", device.codes_to_host(m.table), bits, table=m.table)

    def _update_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        from smqtk_indexing_b200 import device
        import torch
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            new = list(descriptors)
            stale = not self._mirror_in_sync()                   # collaborators were filled elsewhere (from_config over stores)
            LOG.debug("Updating descriptor index.")
            self.descriptor_set.add_many_descriptors(new)

            LOG.debug("Generating hash codes for new descriptors")
            x = self._vectors(new)
            codes, bits = self._hash_matrix(x)
            codes_host = device.codes_to_host(codes)
            ints = bitutil.words_to_ints(codes_host)
            kvstore_update: Dict[Hashable, Set[Hashable]] = {}
            for d, h_int in zip(new, ints):
                if h_int not in kvstore_update:
                    kvstore_update[h_int] = self.hash2uuids_kvstore.get(h_int, set())
                kvstore_update[h_int] |= {d.uuid()}
            LOG.debug("Updating kv-store with new hash codes")
            self.hash2uuids_kvstore.add_many(kvstore_update)

            # device mirror: append new uuids, overwrite re-added ones
            m = self._mirror
            if stale:
                self._sync_mirror()                              # rebuilt from the descriptor set, new elements included
            else:
                xt = torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32)).to(codes.device)
                self._mirror_upsert(m, xt, codes, [d.uuid() for d in new])

            if self.hash_index is not None:
                LOG.debug("Updating hash index structure.")
                self._push_hashes("update", codes_host, bits)

    def _remove_from_index(self, uids: Iterable[Hashable]) -> None:
        from smqtk_indexing_b200 import device
        import torch
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            uids = list(uids)
            stale = not self._mirror_in_sync()
            # KeyError here (unknown uid) leaves everything untouched (lsh.py:407-416)
            descrs = list(self.descriptor_set.get_many_descriptors(uids))
            codes, bits = self._hash_matrix(self._vectors(descrs))
            codes_host = device.codes_to_host(codes)
            h_ints = bitutil.words_to_ints(codes_host)

            removal_rows: List[int] = []
            kvs_update: Dict[Hashable, Set[Hashable]] = {}
            kvs_remove = set()
            for i, (uid, h_int) in enumerate(zip(uids, h_ints)):
                if h_int not in kvs_update:
                    kvs_update[h_int] = self.hash2uuids_kvstore.get(h_int, set())
                kvs_update[h_int] -= {uid}
                if not kvs_update[h_int]:
                    del kvs_update[h_int]
                    kvs_remove.add(h_int)
                    removal_rows.append(i)
            self.hash2uuids_kvstore.add_many(kvs_update)
            self.hash2uuids_kvstore.remove_many(kvs_remove)

            if self.hash_index and removal_rows:
                self._push_hashes("remove", codes_host[removal_rows], bits)

            self.descriptor_set.remove_many_descriptors(uids)

            if stale:
                self._sync_mirror()
            else:
                self._mirror_remove([u for u in set(uids) if self._mirror_row(u) is not None])

    # ------------------------------------------------------------------ queries
    def _nn(self, d: DescriptorElement, n: int = 1) -> Tuple[Tuple[DescriptorElement, ...], Tuple[float, ...]]:
        """Reference-shaped query (lsh.py:452-519)."""
        from smqtk_indexing_b200 import device
        import torch
        LOG.debug("generating hash for descriptor")
        d_v = d.vector()
        d_h = self.lsh_functor.get_hash(d_v)

        with self._model_lock:
            LOG.debug("getting near hashes")
            hi = self.hash_index
            if hi is None:
                # on-the-fly linear index over the KVS keys (lsh.py:481-486)
                hi = LinearHashIndex()
                hi.index = self.hash2uuids_kvstore.keys()
            near_hashes, _ = hi.nn(d_h, n)

            LOG.debug("getting UUIDs of descriptors for nearby hashes")
            neighbor_uuids: List[Hashable] = []
            for h_int in map(bitutil.bit_vector_to_int_large, near_hashes):
                neighbor_uuids.extend(self.hash2uuids_kvstore.get(h_int, set()))
            LOG.debug("-- matched %d UUIDs", len(neighbor_uuids))
            neighbors = list(self.descriptor_set.get_many_descriptors(neighbor_uuids))

        # lock released: re-rank (lsh.py:503-519)
        if not neighbors:
            return (), ()
        dev = device.device()
        cand = torch.from_numpy(numpy.ascontiguousarray(self._vectors(neighbors), dtype=numpy.float32)).to(dev)
        q = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(d_v, dtype=numpy.float32))[None, :]).to(dev)
        m = len(neighbors)
        idx = torch.arange(m, dtype=torch.int64, device=dev)
        off = torch.tensor([0, m], dtype=torch.int64, device=dev)
        dist = device.rerank(cand, q, idx, off, self.distance_method)
        pos, od = device.rerank_select(dist, off, min(n, m))
        pos = pos[0].cpu().numpy()
        od = od[0].cpu().numpy()
        r_descrs = tuple(neighbors[p] for p in pos if p >= 0)
        r_dists = tuple(float(v) for v, p in zip(od, pos) if p >= 0)
        return r_descrs, r_dists

    # ------------------------------------------------------------------ device mirror upkeep
    def _mirror_in_sync(self) -> bool:
        """The reference keeps ALL index state in its collaborators (descriptor_set,
        hash2uuids_kvstore, hash_index), so an instance made by ``from_config`` over persisted stores
        is queryable at once (lsh.py:135-232).  The device mirror is private process state: it is in
        sync when it was fed through the matrix API (its own source of truth) or holds exactly the
        descriptor set's elements."""
        m = self._mirror
        return m.matrix_built or m.num_live == self.descriptor_set.count()

    def _sync_mirror(self) -> None:
        """Rebuild the device mirror from the descriptor set (vectors re-hashed on the device)."""
        import torch
        m = self._mirror
        m.clear()
        elems = list(self.descriptor_set)
        if not elems:
            return
        x = self._vectors(elems)
        codes, _ = self._hash_matrix(x)
        m.uuids = [d.uuid() for d in elems]
        m.row_of = {u: i for i, u in enumerate(m.uuids)}
        m.set_rows(torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32)).to(codes.device), codes)

    def _mirror_row(self, uuid: Hashable) -> Optional[int]:
        m = self._mirror
        if isinstance(m.uuids, range) and m.row_of is None:
            return int(uuid) if isinstance(uuid, (int, numpy.integer)) and 0 <= int(uuid) < m.num_rows else None
        if m.row_of is None:
            m.row_of = {u: i for i, u in enumerate(m.uuids)}
        return m.row_of.get(uuid)

    def _mirror_remove(self, uuids: Sequence[Hashable]) -> None:
        """Tombstone the rows of ``uuids`` (all present and live), re-index; compact once half the rows are dead."""
        import torch
        m = self._mirror
        if not uuids:
            return
        rows = [self._mirror_row(u) for u in uuids]
        sel = torch.tensor(rows, dtype=torch.int64, device=m.codes.device)
        m.remove_rows(sel)
        if m.row_of is not None:
            for u in uuids:
                m.row_of.pop(u, None)
        if m.num_dead * 2 > m.num_rows and not isinstance(m.uuids, range):
            remap = m.compact()                              # re-indexes
            if remap is not None:
                keep = torch.nonzero(remap >= 0).reshape(-1).cpu().tolist()
                m.uuids = [m.uuids[r] for r in keep]
                m.row_of = None
        else:
            m.reindex()

    def build_index_matrix(self, x, uuids: Optional[Sequence[Hashable]] = None) -> None:
        """Bulk build straight from a ``[N, D]`` matrix (numpy or float32 CUDA
        tensor, adopted without a copy): hash -> unique codes -> CSR, all on the
        device.  ``uuids[r]`` names row ``r`` (default: the row number).

        This is the ingest path for tables too large for per-element Python
        containers: the ``descriptor_set`` / ``hash2uuids_kvstore`` collaborators
        are NOT populated (use ``build_index`` for that), so only ``nn_batch`` /
        ``count_rows`` see these rows.  The configured ``hash_index`` receives the
        unique-code table.

        :raises ReadOnlyError: read-only index.  :raises ValueError: empty matrix.
        """
        import torch
        from smqtk_indexing_b200 import device
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            if x is None or len(x) == 0:
                raise self._empty_iterable_exception()
            if not isinstance(x, torch.Tensor):
                x = torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32))
            if not x.is_cuda:
                x = x.to(device.device())
            if x.dtype != torch.float32:
                x = x.to(torch.float32)
            m = self._mirror
            m.clear()
            m.matrix_built = True
            codes = self.lsh_functor.get_hash_packed(x)
            m.uuids = list(uuids) if uuids is not None else range(x.shape[0])  # type: ignore
            m.row_of = None                                      # uuid -> row lookup is built on first use
            m.set_rows(x, codes)
            if isinstance(self.hash_index, LinearHashIndex):
                self.hash_index.set_code_table(m.table)

    @staticmethod
    def _mirror_upsert(m: "_DeviceMirror", xt, codes, uuids: Sequence[Hashable]) -> None:
        """Device-mirror half of ``update_index`` (reference lsh.py:331-383): rows of new uuids are
        appended, a uuid seen before gets its vector and code overwritten (the descriptor set
        overwrites too); a uuid given twice in one batch keeps its last vector.  Re-indexes."""
        import torch
        if m.row_of is None:                                   # matrix-built mirror: build the lookup on first use
            m.row_of = {u: i for i, u in enumerate(m.uuids)}
        if not isinstance(m.uuids, list):
            m.uuids = list(m.uuids)
        last = {u: i for i, u in enumerate(uuids)}
        app_src, ow_src, ow_dst = [], [], []
        for u, i in last.items():
            r = m.row_of.get(u)
            if r is None:
                m.row_of[u] = len(m.uuids)
                m.uuids.append(u)
                app_src.append(i)
            else:
                ow_src.append(i)
                ow_dst.append(r)
        dev = codes.device
        if ow_src:
            src = torch.tensor(ow_src, dtype=torch.int64, device=dev)
            dst = torch.tensor(ow_dst, dtype=torch.int64, device=dev)
            m.overwrite_rows(dst, xt[src], codes[src])
            if m.alive is not None:                            # a removed uuid that comes back is live again
                m.alive[dst] = True
                m.num_dead = m.num_rows - int(m.alive.sum().item())
        if app_src:
            if len(app_src) == len(uuids):
                m.append_rows(xt, codes)
            else:
                sel = torch.tensor(app_src, dtype=torch.int64, device=dev)
                m.append_rows(xt[sel], codes[sel])
        m.reindex()

    def update_index_matrix(self, x, uuids: Optional[Sequence[Hashable]] = None) -> None:
        """``update_index`` (reference lsh.py:331-383) for a ``[N, D]`` matrix, at device speed: hash,
        append into the geometrically growing device buffers, re-index.  ``uuids[r]`` names row
        ``r``; a uuid already present is overwritten in place.  Default uuids continue the row
        numbering.  Like ``build_index_matrix`` this does not touch the ``descriptor_set`` /
        ``hash2uuids_kvstore`` collaborators -- ``hash2uuids_view()`` serves their readers.

        :raises ReadOnlyError: read-only index.  :raises ValueError: empty matrix."""
        import torch
        from smqtk_indexing_b200 import device
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            if x is None or len(x) == 0:
                raise self._empty_iterable_exception()
            m = self._mirror
            if m.num_rows == 0:
                return self.build_index_matrix(x, uuids)
            if not isinstance(x, torch.Tensor):
                x = torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32))
            x = x.to(m.codes.device, torch.float32)
            codes = self.lsh_functor.get_hash_packed(x)
            if uuids is None and isinstance(m.uuids, range) and m.row_of is None:
                m.append_rows(x, codes)                          # pure append, no per-row Python at all
                m.uuids = range(m.num_rows)
                m.reindex()
            else:
                if uuids is None:
                    uuids = range(m.num_rows, m.num_rows + x.shape[0])
                if len(uuids) != x.shape[0]:
                    raise ValueError("uuids must name every row of x")
                self._mirror_upsert(m, x, codes, uuids)
            if isinstance(self.hash_index, LinearHashIndex):
                self.hash_index.set_code_table(m.table)

    def remove_from_index_matrix(self, uuids: Sequence[Hashable]) -> None:
        """``remove_from_index`` (reference lsh.py:385-450) against the device-resident index: the rows
        become tombstones (row numbers of the others do not move) and the unique table / CSR are
        rebuilt over the live rows; buffers are compacted once half of them is dead.

        :raises KeyError: a uuid is not in the index -- nothing is modified (lsh.py:407-416)."""
        import torch
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            uuids = list(uuids)
            if not uuids:
                raise self._empty_iterable_exception()
            m = self._mirror
            rows = [self._mirror_row(u) for u in uuids]
            rows = [-1 if r is None else r for r in rows]
            if m.num_rows == 0 or min(rows) < 0:
                raise KeyError(uuids[rows.index(-1)] if -1 in rows else uuids[0])
            sel = torch.tensor(rows, dtype=torch.int64, device=m.codes.device)
            if m.alive is not None and not bool(m.alive[sel].all().item()):
                raise KeyError(uuids[int((~m.alive[sel]).nonzero()[0].item())])
            self._mirror_remove(uuids)
            if isinstance(self.hash_index, LinearHashIndex):
                self.hash_index.set_code_table(m.table)

    def hash2uuids_view(self) -> "DeviceHashToUuids":
        """Read-only ``KeyValueStore`` over the device index: ``{code int: set(uuid)}`` exactly as
        ``build_index`` would have filled ``hash2uuids_kvstore`` (reference lsh.py:316-323), for
        consumers of that store when the index was fed through the ``*_matrix`` methods."""
        return DeviceHashToUuids(self)

    def save_snapshot(self, path: str) -> None:
        """Write the device-resident index (descriptor rows, codes, tombstones, row -> uuid) to
        ``path`` (``torch.save``), so that a built index survives a restart without re-hashing:
        ``load_snapshot`` re-derives the unique table and the CSR on the device.
        The functor's model is NOT stored -- it has its own ``.npy`` caches (itq.py:212-237); its
        bit length, the descriptor dimension and the distance method are, and are checked on load.
        uuids that are ints / strings (or the default row numbers) load without unpickling."""
        import torch
        with self._model_lock:
            m = self._mirror
            state = m.state_dict()
            state["uuids"] = ("range", len(m.uuids)) if isinstance(m.uuids, range) else ("list", list(m.uuids))
            state["distance_method"] = self.distance_method
            state["bit_length"] = self._functor_bits()
            state["dim"] = None if m.x is None else int(m.x.shape[1])
            torch.save(state, path)

    def _functor_bits(self) -> Optional[int]:
        rot = getattr(self.lsh_functor, "rotation", None)
        if rot is None:
            rot = getattr(self.lsh_functor, "rps", None)
        if rot is None:
            return None
        rot = numpy.asarray(rot)
        return int(rot.reshape(rot.shape[0], -1).shape[1])

    def load_snapshot(self, path: str, trust_pickle: bool = False) -> None:
        """Replace the device-resident index by a snapshot written by ``save_snapshot``.

        The file is read with ``torch.load(weights_only=True)`` (tensors, numbers, strings, lists /
        tuples of those -- nothing is unpickled into code); a snapshot whose uuids are arbitrary
        Python objects needs ``trust_pickle=True`` and must come from a trusted source.

        :raises ReadOnlyError: read-only index.
        :raises ValueError: not a snapshot; or it does not fit this index: its code width differs
            from the functor's bit length, its descriptor dimension differs from the functor's, or it
            was written for another ``distance_method``.  Nothing is modified in that case."""
        import pickle
        import torch
        from smqtk_indexing_b200 import device
        with self._model_lock:
            if self.read_only:
                raise ReadOnlyError("Cannot modify container attributes due to being in read-only mode.")
            try:
                state = torch.load(path, map_location="cpu", weights_only=True)
            except pickle.UnpicklingError as e:
                if not trust_pickle:
                    raise ValueError("snapshot %r holds objects that need unpickling (non-primitive uuids?): "
                                     "pass trust_pickle=True if its source is trusted (%s)" % (path, str(e)[:80]))
                state = torch.load(path, map_location="cpu", weights_only=False)
            if not isinstance(state, dict) or state.get("format") != "smqtk_indexing_b200.DeviceLshIndex/1":
                raise ValueError("not a DeviceLshIndex snapshot: %r" % (path,))
            bits = self._functor_bits()
            if state.get("codes") is not None and bits is not None:
                if state.get("bit_length") not in (None, bits):
                    raise ValueError("snapshot was hashed to %s bits, the functor produces %d" % (state["bit_length"], bits))
                if int(state["codes"].shape[1]) != bitutil.words_for_bits(bits):
                    raise ValueError("snapshot codes are %d words wide, the functor's %d-bit codes need %d"
                                     % (state["codes"].shape[1], bits, bitutil.words_for_bits(bits)))
            mean = getattr(self.lsh_functor, "mean_vec", None)
            if state.get("x") is not None and mean is not None and int(state["x"].shape[1]) != int(numpy.asarray(mean).shape[0]):
                raise ValueError("snapshot descriptors have %d dimensions, the functor's model %d"
                                 % (state["x"].shape[1], numpy.asarray(mean).shape[0]))
            if state.get("distance_method") not in (None, self.distance_method):
                raise ValueError("snapshot was written for distance_method=%r, this index uses %r"
                                 % (state["distance_method"], self.distance_method))
            m = self._mirror
            m.load_state_dict(state, device.device())
            m.matrix_built = True
            kind, val = state["uuids"]
            m.uuids = range(val) if kind == "range" else list(val)     # type: ignore
            m.row_of = None
            if isinstance(self.hash_index, LinearHashIndex):
                self.hash_index.set_code_table(m.table)

    def count_rows(self) -> int:
        """Live descriptor rows in the device-resident index (see ``build_index_matrix``)."""
        return self._mirror.num_live

    def nn_batch(self, queries, n: int = 1, return_device: bool = False, graph: bool = True):
        """Batched LSH query against the device-resident index.

        :param queries: ``[Q, D]`` float array (host) or float32 CUDA tensor.
        :param n: neighbours per query (also the number of nearest unique codes
            whose descriptors form the candidate pool, as in the reference).
        :param graph: replay the pipeline as one CUDA graph once the same batch shape repeats
            (``False``: always launch kernel by kernel).
        :return: ``(rows int64[Q, n], dists float64[Q, n])`` -- rows index
            :meth:`mirror_uuids`; -1 / NaN pad queries with fewer than ``n``
            candidates.  Device tensors when ``return_device``.
        :raises ValueError: the index is empty.
        """
        import torch
        from smqtk_indexing_b200 import device
        with self._model_lock:
            m = self._mirror
            if not self._mirror_in_sync():
                # the collaborators were populated elsewhere (e.g. from_config over persisted stores) or
                # changed behind this object: the mirror is rebuilt from the descriptor set, never queried partial
                self._sync_mirror()
            if m.table is None:
                raise ValueError("No index currently set to query from!")
            dev = m.x.device
            use_graph = graph and hasattr(self.lsh_functor, "get_hash_packed")
            if isinstance(queries, torch.Tensor):
                q = queries if (use_graph and not queries.is_cuda and queries.dtype == torch.float32) else \
                    queries.to(dev, torch.float32)
            else:
                q = torch.from_numpy(numpy.ascontiguousarray(queries, dtype=numpy.float32))
                if not use_graph:
                    q = q.to(dev, non_blocking=True)
            if q.dim() == 1:
                q = q.unsqueeze(0)
            if use_graph:
                # host batches go straight into the captured pipeline's static input, results straight out of
                # its static output: one H2D and one D2H copy per call, nothing else on the host
                out = m.query_graphed(self.lsh_functor, q.contiguous(), n, self.distance_method,
                                      to_host=not return_device)
                if not return_device:
                    return out[0].numpy(), out[1].numpy()
                return out
            rows, dists = m.query(self.lsh_functor, q, n, self.distance_method)
        if return_device:
            return rows, dists
        return rows.cpu().numpy(), dists.cpu().numpy()

    def mirror_uuids(self) -> List[Hashable]:
        """row -> uuid for the rows returned by :meth:`nn_batch`."""
        return self._mirror.uuids


class DeviceHashToUuids(KeyValueStore):
    """Read-only ``KeyValueStore`` view ``{code int: set(uuid)}`` of a device-resident LSH index
    (the shape ``LSHNearestNeighborIndex.build_index`` gives ``hash2uuids_kvstore``, reference
    lsh.py:316-323).  Values are materialised from the code -> rows CSR on access; a snapshot of
    the index at construction time is NOT taken -- the view follows later updates."""

    def __init__(self, index: "LSHNearestNeighborIndex") -> None:
        # no super().__init__(): Pluggable's constructor only enforces is_usable(), which is False on
        # purpose (see below) -- instances are made by LSHNearestNeighborIndex.hash2uuids_view() only
        self._index = index

    @classmethod
    def is_usable(cls) -> bool:
        # a view bound to one index object, not something a config file can name: keep it out of
        # ``KeyValueStore.get_impls()`` / default-config generation
        return False

    def get_config(self) -> Dict[str, Any]:
        return {}

    def __repr__(self) -> str:
        return "<DeviceHashToUuids codes=%d>" % self.count()

    def _state(self):
        from smqtk_indexing_b200 import device
        m = self._index._mirror
        if m.table is None:
            return [], None, None, m
        ints = bitutil.words_to_ints(device.codes_to_host(m.table))
        return ints, m.csr_off.cpu().numpy(), m.csr_rows.cpu().numpy(), m

    def count(self) -> int:
        return self._index._mirror.num_codes

    def keys(self) -> Iterator[Hashable]:
        return iter(self._state()[0])

    def values(self) -> Iterator[Any]:
        ints, off, rows, m = self._state()
        for i in range(len(ints)):
            yield {m.uuids[r] for r in rows[off[i]:off[i + 1]]}

    def is_read_only(self) -> bool:
        return True

    def _find(self, key: Hashable):
        ints, off, rows, m = self._state()
        import bisect
        i = bisect.bisect_left(ints, key) if isinstance(key, int) else len(ints)
        if i < len(ints) and ints[i] == key:
            return {m.uuids[r] for r in rows[off[i]:off[i + 1]]}
        return None

    def has(self, key: Hashable) -> bool:
        return self._find(key) is not None

    def get(self, key: Hashable, default: Any = NO_DEFAULT_VALUE) -> Any:
        v = self._find(key)
        if v is None:
            if default is NO_DEFAULT_VALUE:
                raise KeyError(key)
            return default
        return v

    def add(self, key: Hashable, value: Any) -> "KeyValueStore":
        raise ReadOnlyError("Cannot add to read-only instance %s." % self)

    def add_many(self, d) -> "KeyValueStore":
        raise ReadOnlyError("Cannot add to read-only instance %s." % self)

    def remove(self, key: Hashable) -> "KeyValueStore":
        raise ReadOnlyError("Cannot remove from read-only instance %s." % self)

    def remove_many(self, keys) -> "KeyValueStore":
        raise ReadOnlyError("Cannot remove from read-only instance %s." % self)

    def clear(self) -> "KeyValueStore":
        raise ReadOnlyError("Cannot clear a read-only %s instance." % type(self).__name__)
