"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

Literal restatement of the reference's QUERY path on the reference's own data
structures (a Python ``set`` of arbitrary-precision ints, a ``dict`` standing in
for the ``hash2uuids`` key-value store, a ``dict`` of descriptor vectors), kept
statement-for-statement close to the reference so that timing it on the GPU
box's host cores is timing what a user of SMQTK-Indexing runs today.  Used by
``bench.py`` (``cpu_baseline`` and ``--impl reference``) and by the tests that
pin it to the golden fixtures; nothing under ``smqtk_indexing_b200/`` imports it.

Citations are relative to the reference root (/root/reference):
  ItqFunctor.get_hash                 smqtk_indexing/impls/lsh_functor/itq.py:389-408
  bit_vector_to_int_large             smqtk_indexing/utils/bits.py:4-20
  LinearHashIndex._nn                 smqtk_indexing/impls/hash_index/linear.py:206-244
  hamming_distance                    smqtk_indexing/utils/metrics.py:140-155
  LSHNearestNeighborIndex._nn         smqtk_indexing/impls/nn_index/lsh.py:452-519

Parity status: PINNED -- ``tests/test_oracle.py::test_ref_port_golden`` replays
the golden LSH fixtures (outputs of the unmodified reference) through this class.
The reference is single-threaded pure Python on this path; so is this.
"""
import heapq
from typing import Dict, Hashable, Iterable, List, Sequence, Set, Tuple

import numpy as np

import np_oracle as O


class LiteralLshIndex:
    """State + query of ``LSHNearestNeighborIndex`` with ``ItqFunctor`` and a
    (possibly implicit) ``LinearHashIndex``."""

    def __init__(self, mean_vec: np.ndarray, rotation: np.ndarray, distance_method: str = "euclidean",
                 normalize=None) -> None:
        self.mean_vec = mean_vec
        self.rotation = rotation
        self.normalize = normalize
        self.distance_function = O.DISTANCE_FUNCTIONS[distance_method]
        self.index: Set[int] = set()                       # LinearHashIndex.index (linear.py:110)
        self.hash2uuids: Dict[int, Set[Hashable]] = {}     # hash2uuids_kvstore
        self.vectors: Dict[Hashable, np.ndarray] = {}      # descriptor_set

    # -- lsh_functor.get_hash (itq.py:404-408)
    def get_hash(self, descriptor: np.ndarray) -> np.ndarray:
        z = np.dot(O.norm_vector(descriptor, self.normalize) - self.mean_vec, self.rotation)
        b = np.zeros(z.shape, dtype=bool)
        b[z >= 0] = True
        return b

    # -- _build_index (lsh.py:316-329 + linear.py:163), one descriptor at a time
    def build(self, uuids: Sequence[Hashable], vectors: Iterable[np.ndarray]) -> None:
        self.index, self.hash2uuids, self.vectors = set(), {}, {}
        for u, v in zip(uuids, vectors):
            self.vectors[u] = v
            h_int = O.bit_vector_to_int(self.get_hash(v))
            self.hash2uuids.setdefault(h_int, set()).add(u)
            self.index.add(h_int)

    def adopt_codes(self, uuids: Sequence[Hashable], vectors, code_ints: Sequence[int]) -> None:
        """Install an index whose codes were computed elsewhere (bench set-up: the
        per-descriptor build loop above is not part of the query metric)."""
        self.index, self.hash2uuids = set(), {}
        self.vectors = vectors                             # anything indexable by uuid
        for u, h_int in zip(uuids, code_ints):
            self.hash2uuids.setdefault(h_int, set()).add(u)
            self.index.add(h_int)

    # -- LinearHashIndex._nn (linear.py:232-244)
    def hash_nn(self, h: np.ndarray, n: int) -> Tuple[np.ndarray, Tuple[float, ...]]:
        h_int = O.bit_vector_to_int(h)
        bits = len(h)
        near_codes = heapq.nsmallest(n, self.index, lambda e: O.hamming_distance_int(h_int, e))
        distances = [O.hamming_distance_int(h_int, c) for c in near_codes]
        hash_vectors = np.vstack([O.int_to_bit_vector(c, bits) for c in near_codes])
        return hash_vectors, tuple(np.array(distances, dtype=float) / bits)

    # -- LSHNearestNeighborIndex._nn (lsh.py:470-519)
    def nn(self, d_v: np.ndarray, n: int) -> Tuple[List[Hashable], List[float]]:
        d_h = self.get_hash(d_v)
        hashes, _ = self.hash_nn(d_h, n)
        neighbor_uuids: List[Hashable] = []
        for h_int in map(O.bit_vector_to_int, hashes):
            neighbor_uuids.extend(self.hash2uuids.get(h_int, set()))
        neighbor_vectors = np.asarray([self.vectors[u] for u in neighbor_uuids])
        distances = [self.distance_function(d_v, v) for v in neighbor_vectors]
        ordered = sorted(zip(neighbor_uuids, distances), key=lambda p: p[1])
        return [p[0] for p in ordered[:n]], [float(p[1]) for p in ordered[:n]]


def words_to_ints(words: np.ndarray) -> List[int]:
    """uint32[n, W] (word 0 most significant) -> Python ints."""
    be = np.ascontiguousarray(words, dtype=np.uint32).astype(">u4")
    w = be.shape[1] * 4
    raw = be.tobytes()
    return [int.from_bytes(raw[i:i + w], "big") for i in range(0, len(raw), w)]
