"""
B200 ``LinearHashIndex``: brute-force Hamming nearest neighbours over the set of
unique hash codes, resident in HBM as a sorted ``uint32[U, W]`` table and scanned
by the ``sb_hamming_scan`` kernel.

Drop-in for ``smqtk_indexing.impls.hash_index.linear.LinearHashIndex``
(reference: smqtk_indexing/impls/hash_index/linear.py:28-244): same constructor,
config keys, methods, return types and exceptions.  Differences, all documented
in DESIGN.md: (1) ties between equally distant codes are returned in ascending
code value instead of CPython set-iteration order (the reference leaves this
undefined); (2) the cache element stores the packed table for codes wider than
63 bits (the reference's ``numpy.save(tuple(ints))`` format is lossy from 64 bits
up); reference-format caches of narrower codes are read and written unchanged.
"""
import threading
from io import BytesIO
from typing import Any, Dict, Iterable, Optional, Set, Tuple, Type, TypeVar

import numpy

from smqtk_core.configuration import from_config_dict, make_default_config, to_config_dict
from smqtk_core.dict import merge_dict
from smqtk_dataprovider import DataElement

from smqtk_indexing_b200.interfaces import HashIndex
from smqtk_indexing_b200.utils import bits as bitutil

T = TypeVar("T", bound="LinearHashIndex")


class LinearHashIndex(HashIndex):
    """Linear (exhaustive) Hamming index on the GPU."""

    @classmethod
    def is_usable(cls) -> bool:
        # library built and a CUDA device visible (no CPU implementation exists to fall back to)
        from smqtk_indexing_b200 import _lib
        return _lib.usable()

    @classmethod
    def get_default_config(cls) -> Dict[str, Any]:
        c = super(LinearHashIndex, cls).get_default_config()
        c['cache_element'] = make_default_config(DataElement.get_impls())
        return c

    @classmethod
    def from_config(cls: Type[T], config_dict: Dict, merge_default: bool = True) -> T:
        if merge_default:
            config_dict = merge_dict(cls.get_default_config(), config_dict)
        cache_element = None
        if config_dict['cache_element'] and config_dict['cache_element']['type']:
            cache_element = from_config_dict(config_dict['cache_element'], DataElement.get_impls())
        config_dict['cache_element'] = cache_element
        return super(LinearHashIndex, cls).from_config(config_dict, False)

    def __init__(self, cache_element: Optional[DataElement] = None):
        super(LinearHashIndex, self).__init__()
        self.cache_element = cache_element
        self._model_lock = threading.RLock()
        #: sorted unique codes, int32[U, W] on the device (None = empty index)
        self._table = None
        #: host copy of the codes while no device is needed yet (set / cache load)
        self._pending_words: Optional[numpy.ndarray] = None
        self.load_cache()

    def get_config(self) -> Dict[str, Any]:
        c = self.get_default_config()
        if self.cache_element:
            c['cache_element'] = merge_dict(c['cache_element'], to_config_dict(self.cache_element))
        return c

    # ------------------------------------------------------------------ state
    def _device_table(self):
        """The code table on the device (uploads a pending host copy)."""
        if self._table is None and self._pending_words is not None:
            from smqtk_indexing_b200 import codes as codeops, device
            t = device.codes_to_device(self._pending_words)
            self._table = codeops.sort_unique(t)
            self._pending_words = None
        return self._table

    def _host_words(self) -> numpy.ndarray:
        if self._table is not None:
            from smqtk_indexing_b200 import device
            return device.codes_to_host(self._table)
        if self._pending_words is not None:
            return self._pending_words
        return numpy.zeros((0, 1), numpy.uint32)

    @property
    def index(self) -> Set[int]:
        """The indexed codes as a set of Python ints (reference attribute,
        linear.py:110); materialised from the device table on access."""
        with self._model_lock:
            return set(bitutil.words_to_ints(self._host_words()))

    @index.setter
    def index(self, values: Iterable[int]) -> None:
        """Replace the index content by integer codes (used by
        LSHNearestNeighborIndex's on-the-fly index, lsh.py:481-486)."""
        with self._model_lock:
            ints = [int(v) for v in values]
            self._table = None
            if not ints:
                self._pending_words = None
                return
            nbits = max(max(i.bit_length() for i in ints), 1)
            w = numpy.unique(bitutil.ints_to_words(ints, bitutil.words_for_bits(nbits)), axis=0)
            self._pending_words = w

    @property
    def code_table(self):
        """Sorted unique codes as an ``int32[U, W]`` CUDA tensor (None if empty)."""
        with self._model_lock:
            return self._device_table()

    def set_code_table(self, table) -> None:
        """Adopt an already sorted-unique device table (batch build path)."""
        with self._model_lock:
            self._table = table if (table is not None and table.shape[0]) else None
            self._pending_words = None

    # ------------------------------------------------------------------ cache
    def load_cache(self) -> None:
        """Load from the cache element: either the reference's 1-D integer
        ``.npy`` (linear.py:121-129) or this class's 2-D packed uint32 table."""
        with self._model_lock:
            if self.cache_element and not self.cache_element.is_empty():
                arr = numpy.load(BytesIO(self.cache_element.get_bytes()))
                if arr.ndim == 2 and arr.dtype == numpy.uint32:
                    self._table = None
                    self._pending_words = numpy.unique(arr, axis=0) if len(arr) else None
                else:
                    self.index = [int(v) for v in arr.ravel()]

    def save_cache(self) -> None:
        with self._model_lock:
            if self.cache_element and self.count():
                if self.cache_element.is_read_only():
                    raise ValueError("Cache element (%s) is read-only." % self.cache_element)
                words = self._host_words()
                ints = bitutil.words_to_ints(words)
                buff = BytesIO()
                if max(i.bit_length() for i in ints) <= 63:
                    # reference format: 1-D int64 array of the code integers
                    numpy.save(buff, numpy.asarray(ints, dtype=numpy.int64))
                else:
                    numpy.save(buff, words)
                self.cache_element.set_bytes(buff.getvalue())

    # ------------------------------------------------------------------ interface
    def count(self) -> int:
        with self._model_lock:
            if self._table is not None:
                return int(self._table.shape[0])
            if self._pending_words is not None:
                return int(self._pending_words.shape[0])
            return 0

    @staticmethod
    def _pack_iterable(hashes: Iterable[numpy.ndarray]) -> numpy.ndarray:
        """Iterable of equal-length bit vectors -> uint32[n, W] (host marshalling)."""
        mat = numpy.asarray(list(hashes))
        if mat.ndim != 2:
            raise ValueError("hash vectors must all have the same length")
        return bitutil.pack_bits(mat.astype(bool), bitutil.words_for_bits(max(mat.shape[1], 1)))

    def _merge_width(self, new_words: numpy.ndarray):
        """Bring the device table and ``new_words`` to a common word count."""
        from smqtk_indexing_b200 import codes as codeops, device
        new_t = device.codes_to_device(new_words)
        table = self._device_table()
        if table is None:
            return None, new_t
        w = max(table.shape[1], new_t.shape[1])
        return codeops.widen(table, w), codeops.widen(new_t, w)

    def _build_index(self, hashes: Iterable[numpy.ndarray]) -> None:
        from smqtk_indexing_b200 import codes as codeops, device
        words = self._pack_iterable(hashes)
        with self._model_lock:
            new_table = codeops.sort_unique(device.codes_to_device(words))
            self._table, self._pending_words = new_table, None
            self.save_cache()

    def _update_index(self, hashes: Iterable[numpy.ndarray]) -> None:
        from smqtk_indexing_b200 import codes as codeops
        words = self._pack_iterable(hashes)
        with self._model_lock:
            table, new_t = self._merge_width(words)
            merged = codeops.sort_unique(new_t) if table is None else codeops.union(table, new_t)
            self._table, self._pending_words = merged, None
            self.save_cache()

    def _remove_from_index(self, hashes: Iterable[numpy.ndarray]) -> None:
        from smqtk_indexing_b200 import codes as codeops, device
        words = self._pack_iterable(hashes)
        with self._model_lock:
            table, rem_t = self._merge_width(words)
            if table is None:
                raise KeyError(bitutil.words_to_ints(words[:1])[0])
            remaining, missing = codeops.difference(table, rem_t)
            if missing.shape[0]:
                # KeyError before any mutation (linear.py:199-203)
                raise KeyError(bitutil.words_to_ints(device.codes_to_host(missing[:1]))[0])
            self._table = remaining if remaining.shape[0] else None
            self._pending_words = None
            self.save_cache()

    def nn_packed(self, q_codes, n: int):
        """Batch query on device tensors: ``q_codes`` int32[Q, W'] ->
        (dist int32[Q, n], row int64[Q, n]) into :attr:`code_table`; -1 pads
        when the index holds fewer than ``n`` codes."""
        from smqtk_indexing_b200 import codes as codeops, device
        with self._model_lock:
            table = self._device_table()
            if table is None:
                raise ValueError("No index currently set to query from!")
            w = max(table.shape[1], q_codes.shape[1])
            if w != table.shape[1]:
                table = self._table = codeops.widen(table, w)
            q = codeops.widen(q_codes, w).contiguous()
            return device.hamming_topk(table, q, n)

    def _nn(self, h: numpy.ndarray, n: int = 1) -> Tuple[numpy.ndarray, Tuple[float, ...]]:
        from smqtk_indexing_b200 import device
        h = numpy.asarray(h).astype(bool).ravel()
        bits = len(h)
        with self._model_lock:
            q = device.codes_to_device(bitutil.pack_bits(h, bitutil.words_for_bits(max(bits, 1))))
            dist_t, rows_t = self.nn_packed(q, n)
            keep = rows_t[0] >= 0
            near = device.codes_to_host(self._table[rows_t[0][keep]])
            dist = dist_t[0][keep].cpu().numpy()
        return (
            bitutil.unpack_bits(near, bits),
            tuple(float(d) / float(bits) for d in dist),
        )
