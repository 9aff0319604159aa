"""
``SkLearnBallTreeHashIndex`` by name, on the GPU linear scan.

The reference's class (smqtk_indexing/impls/hash_index/sklearn_balltree.py:33-375) answers EXACT
Hamming nearest-neighbour queries over the set of unique hash codes with a scikit-learn ``BallTree``
(``metric='hamming'``, :228-230) and returns normalised distances ``d / b`` (:366-375) -- the same
contract as ``LinearHashIndex``; the tree only changes how the CPU finds the answer.  On a B200 the
exhaustive tensor-core / XOR-POPC scan is faster than any tree walk, so this class keeps the
reference's constructor, config keys, ``save_model`` / ``load_model`` names and exceptions and
serves them from ``LinearHashIndex``'s device table.  A config file that names
``...hash_index.sklearn_balltree.SkLearnBallTreeHashIndex`` therefore keeps working after the
module prefix is switched to ``smqtk_indexing_b200``.

Differences (documented, none changes a query result beyond tie order):
  * ``leaf_size`` and ``random_seed`` are accepted, stored and round-tripped through the config but
    have no effect (there is no tree);
  * ties are returned in ascending code value (the tree's order is unspecified);
  * the cache element is READ in the reference's ``.npz`` layout (only ``data_arr``, the 0/1 matrix of
    indexed codes, is used; the pickled tree internals are ignored) and WRITTEN in
    ``LinearHashIndex``'s layout.
"""
from io import BytesIO
from typing import Any, Dict, Optional, Tuple

import numpy

from smqtk_core.configuration import to_config_dict
from smqtk_core.dict import merge_dict
from smqtk_dataprovider import DataElement

from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
from smqtk_indexing_b200.utils import bits as bitutil


class SkLearnBallTreeHashIndex(LinearHashIndex):
    """Exact Hamming index under the reference's ball-tree class name (GPU linear scan inside)."""

    def __init__(self, cache_element: Optional[DataElement] = None, leaf_size: int = 40,
                 random_seed: Optional[int] = None):
        self.leaf_size = leaf_size
        self.random_seed = random_seed
        super(SkLearnBallTreeHashIndex, self).__init__(cache_element)

    def get_config(self) -> Dict[str, Any]:
        # reference sklearn_balltree.py:143-151
        c = merge_dict(self.get_default_config(), {
            'leaf_size': self.leaf_size,
            'random_seed': self.random_seed,
        })
        if self.cache_element:
            c['cache_element'] = merge_dict(c['cache_element'], to_config_dict(self.cache_element))
        return c

    # reference method names (sklearn_balltree.py:153-210)
    def save_model(self) -> None:
        """:raises ValueError: the cache element is read-only (sklearn_balltree.py:163-165)."""
        with self._model_lock:
            if self.cache_element:
                if self.cache_element.is_read_only():
                    raise ValueError("Configured cache element (%s) is read-only." % self.cache_element)
                if self.count():
                    self.save_cache()
                else:
                    self.cache_element.set_bytes(b'')          # no index: empty cache (:188-189)

    def load_model(self) -> None:
        self.load_cache()

    def load_cache(self) -> None:
        with self._model_lock:
            if self.cache_element and not self.cache_element.is_empty():
                raw = self.cache_element.get_bytes()
                if raw[:2] == b'PK':                           # the reference's np.savez archive (:179-184)
                    with numpy.load(BytesIO(raw), allow_pickle=False) as cache:
                        data = numpy.asarray(cache['data_arr'])
                    self._table = None
                    if data.ndim == 2 and data.shape[0]:
                        words = bitutil.pack_bits(data.astype(bool), bitutil.words_for_bits(max(data.shape[1], 1)))
                        self._pending_words = numpy.unique(words, axis=0)
                    else:
                        self._pending_words = None
                    return
            super(SkLearnBallTreeHashIndex, self).load_cache()

    def _remove_from_index(self, hashes) -> None:
        # an empty index raises KeyError for anything (sklearn_balltree.py:332-335); LinearHashIndex does too
        super(SkLearnBallTreeHashIndex, self)._remove_from_index(hashes)
        if not self.count():
            self.save_model()

    def _nn(self, h: numpy.ndarray, n: int = 1) -> Tuple[numpy.ndarray, Tuple[float, ...]]:
        if not self.count():
            raise RuntimeError("No index currently available to query from.")      # :363-365
        return super(SkLearnBallTreeHashIndex, self)._nn(h, n)
