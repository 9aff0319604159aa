"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU (numpy / pure-Python) restatement of SMQTK-Indexing's LSH nearest-neighbour
path, used as the parity checker for the CUDA implementation.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module; nothing under ``smqtk_indexing_b200/``
does (the product path fails loudly without its CUDA library instead of falling
back to this code).

Parity status: PINNED.  Every function here is checked in ``tests/test_oracle.py``
against golden vectors produced by importing the unmodified reference
(``/root/reference``, v0.18.0) in the build container -- generator:
``oracle/gen_golden.py``, fixtures: ``tests/golden/*.npz`` -- and against the
known answers of the reference's own unit tests.  Unpinned by
the reference itself and therefore defined here: tie order (see
``hamming_topk``), ITQ ``fit`` beyond b=1 (pinned only against the reference's
output on this container's LAPACK, see ``itq_fit``).

All citations are relative to the reference root.
"""
import heapq
from math import pi
from typing import Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------
# bit vectors <-> ints  (smqtk_indexing/utils/bits.py:4-56)
# --------------------------------------------------------------------------
def bit_vector_to_int(v: Sequence) -> int:
    """Literal restatement of utils/bits.py:17-20 (index 0 = MSB)."""
    c = 0
    for b in v:
        c = (c << 1) + int(b)
    return c


def int_to_bit_vector(integer: int, bits: int = 0) -> np.ndarray:
    """Literal restatement of utils/bits.py:43-56."""
    size = len(bin(integer)) - 2
    if bits and (bits - size) < 0:
        raise ValueError("%d bits too small to represent integer value %d."
                         % (bits, integer))
    v = np.zeros(bits or size, np.bool_)
    for i in range(0, size):
        v[-(i + 1)] = integer & 1
        integer >>= 1
    return v


def pack_codes(bitmat: np.ndarray, words: int) -> np.ndarray:
    """bool[n,b] -> uint32[n,words]: each row is the integer value of
    utils/bits.py:17-20 written big-endian in 32-bit words (word 0 = most
    significant).  Independent of the product's own packer on purpose."""
    bitmat = np.atleast_2d(np.asarray(bitmat)).astype(bool)
    n, b = bitmat.shape
    out = np.zeros((n, words), np.uint32)
    for j in range(b):
        p = b - 1 - j                      # integer bit position of vector index j
        out[:, words - 1 - p // 32] |= bitmat[:, j].astype(np.uint32) << np.uint32(p % 32)
    return out


# --------------------------------------------------------------------------
# Hamming distance / linear scan  (utils/metrics.py:140-155, linear.py:206-244)
# --------------------------------------------------------------------------
def hamming_distance_int(i: int, j: int) -> int:
    """utils/metrics.py:155."""
    return bin(i ^ j).count('1')


def hamming_topk_literal(index_ints: Sequence[int], h_int: int, n: int):
    """Literal restatement of ``LinearHashIndex._nn`` (linear.py:232-244) over an
    *ordered* sequence of unique ints, with the canonical tie order made
    explicit: ``heapq.nsmallest`` is stable, so iterating the codes in table
    order breaks ties by table row.  Returns (rows, distances)."""
    order = heapq.nsmallest(
        n, range(len(index_ints)),
        key=lambda r: hamming_distance_int(h_int, index_ints[r]))
    return order, [hamming_distance_int(h_int, index_ints[r]) for r in order]


def hamming_distances(db_words: np.ndarray, q_words: np.ndarray) -> np.ndarray:
    """popcount(h xor e) for one query against every row (metrics.py:155),
    vectorised over uint32 words."""
    x = np.bitwise_xor(db_words, q_words[None, :])
    return np.bitwise_count(x).sum(axis=1, dtype=np.int64).astype(np.int32)


def hamming_topk(db_words: np.ndarray, q_words: np.ndarray, k: int,
                 idx_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """``k`` smallest Hamming distances per query over the rows of ``db_words``
    (linear.py:235-240), canonical order: ascending (distance, row).

    The reference's order among equal distances is CPython set-iteration order
    (linear.py:235 iterates a ``set``) and is asserted nowhere
    (tests/impls/hash_index/test_linear.py:150-153 compares tie *groups*), so
    the oracle fixes it: ties are broken by row index in the code table.

    :return: (dist int32[Q, k'], row int64[Q, k']) with k' = min(k, U).
    """
    db_words = np.ascontiguousarray(db_words, dtype=np.uint32)
    q_words = np.atleast_2d(np.ascontiguousarray(q_words, dtype=np.uint32))
    U = db_words.shape[0]
    kk = min(k, U)
    Q = q_words.shape[0]
    out_d = np.empty((Q, kk), np.int32)
    out_i = np.empty((Q, kk), np.int64)
    for qi in range(Q):
        d = hamming_distances(db_words, q_words[qi])
        if kk < U:
            # all rows with d <= (kk-th smallest) then exact canonical order
            thr = np.partition(d, kk - 1)[kk - 1]
            cand = np.flatnonzero(d <= thr)
        else:
            cand = np.arange(U)
        o = np.lexsort((cand, d[cand]))[:kk]
        out_i[qi] = cand[o] + idx_base
        out_d[qi] = d[cand[o]]
    return out_d, out_i


def normalised_hamming(dist: np.ndarray, bits: int) -> np.ndarray:
    """linear.py:243: ``d / float(bits)``."""
    return np.asarray(dist, np.float64) / float(bits)


def merge_topk(dists: Sequence[np.ndarray], idxs: Sequence[np.ndarray], k: int):
    """Merge per-shard top-k lists into the global top-k in canonical
    (distance, global row) order -- defines the multi-GPU result: identical to
    ``hamming_topk`` over the concatenated table."""
    d = np.concatenate(dists, axis=1)
    i = np.concatenate(idxs, axis=1)
    Q = d.shape[0]
    kk = min(k, d.shape[1])
    od = np.empty((Q, kk), d.dtype)
    oi = np.empty((Q, kk), i.dtype)
    for q in range(Q):
        o = np.lexsort((i[q], d[q]))[:kk]
        od[q], oi[q] = d[q][o], i[q][o]
    return od, oi


# --------------------------------------------------------------------------
# ITQ functor  (impls/lsh_functor/itq.py)
# --------------------------------------------------------------------------
def norm_vector(v: np.ndarray, normalize) -> np.ndarray:
    """itq.py:172-191."""
    if normalize is not None:
        n = np.linalg.norm(v, normalize, v.ndim - 1, keepdims=True)
        n[n == 0.] = 1.
        return v / n
    return v


def itq_project(x: np.ndarray, mean_vec: np.ndarray, rotation: np.ndarray,
                normalize=None) -> np.ndarray:
    """``z`` of itq.py:404-405: (norm(x) - mean) . R, in numpy's promoted dtype
    (float64 as soon as R is float64, which it always is after ``fit``)."""
    return np.dot(norm_vector(np.asarray(x), normalize) - mean_vec, rotation)


def itq_hash(x: np.ndarray, mean_vec: np.ndarray, rotation: np.ndarray,
             normalize=None) -> np.ndarray:
    """itq.py:404-408: bits = (z >= 0)."""
    return itq_project(x, mean_vec, rotation, normalize) >= 0


def find_itq_rotation(v: np.ndarray, n_iter: int, random_seed: Optional[int]):
    """itq.py:239-289, including the quirk that numpy's svd returns V^H so the
    update is R = V^H . U^T (itq.py:276-277)."""
    bit = v.shape[1]
    if random_seed is not None:
        np.random.seed(random_seed)
    r = np.random.randn(bit, bit)
    u11, _, _ = np.linalg.svd(r)
    r = u11[:, :bit]
    for _ in range(n_iter):
        z = np.dot(v, r)
        ux = np.where(z >= 0, 1.0, -1.0)
        c = np.dot(ux.transpose(), v)
        ub, _, ua = np.linalg.svd(c)
        r = np.dot(ua, ub.transpose())
    z = np.dot(v, r)
    return z >= 0, r


def itq_fit(x: np.ndarray, bit_length: int, itq_iterations: int = 50,
            normalize=None, random_seed: Optional[int] = None):
    """itq.py:338-383 on an already-assembled [N, D] matrix.

    :return: (codes bool[N,b], mean_vec[D], rotation[D,b])
    """
    x = np.array(x, copy=True)
    if x.shape[1] < bit_length:
        raise ValueError("Input descriptors have fewer features than "
                         "requested bit encoding.")
    x = norm_vector(x, normalize)
    mean_vec = np.mean(x, axis=0)
    x = x - mean_vec
    c = np.cov(x.transpose())
    c = np.atleast_2d(c)
    l, pc = np.linalg.eig(c)
    order = sorted(range(len(l)), key=lambda i_: l[i_], reverse=True)   # stable, like :368
    pc_top = np.array([pc[:, i_] for i_ in order[:bit_length]]).transpose()
    v = np.dot(x, pc_top)
    codes, r = find_itq_rotation(v, itq_iterations, random_seed)
    return codes, mean_vec, np.dot(pc_top, r)


# --------------------------------------------------------------------------
# Re-rank distances  (utils/metrics.py:7-137), 1-D query vs 2-D candidates
# --------------------------------------------------------------------------
def euclidean_distance(i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """metrics.py:83-86."""
    i = np.asarray(i, np.float64)
    j = np.asarray(j, np.float64)
    axis = 0 if (i.ndim == 1 and j.ndim == 1) else 1
    return np.sqrt(np.square(i - j).sum(axis))


def histogram_intersection_distance(i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """metrics.py:41-46 / :70: 1 - 0.5 * sum(i + j - |i - j|)."""
    i = np.asarray(i, np.float64)
    j = np.asarray(j, np.float64)
    axis = 0 if (i.ndim == 1 and j.ndim == 1) else 1
    return 1. - ((np.add(i, j) - np.abs(np.subtract(i, j))).sum(axis) * 0.5)


def cosine_distance(i: np.ndarray, j: np.ndarray, pos_vectors: bool = True) -> np.ndarray:
    """metrics.py:103-137.  The reference takes ``1 - cdist(i, j, 'cosine')``
    from scipy (metrics.py:111; scipy pinned 1.5.4 in poetry.lock:960-961),
    whose published definition is 1 - u.v / (|u|_2 |v|_2); restated directly so
    the oracle does not depend on scipy, and checked against the reference +
    scipy in the golden fixtures."""
    i = np.asarray(i, np.float64)
    j = np.atleast_2d(np.asarray(j, np.float64))
    uv = j @ i
    den = np.sqrt((i * i).sum()) * np.sqrt((j * j).sum(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        sim = uv / den
    sim = np.maximum(np.minimum(sim, 1), -1)
    out = (1 + bool(pos_vectors)) * np.arccos(sim) / pi
    return out[0] if out.size == 1 else out


def l2_topk(x_db: np.ndarray, q: np.ndarray, k: int):
    """Exact flat L2 k-nearest rows: what impls/nn_index/faiss.py:751-831 returns for an
    'IDMap,Flat' L2 index whose search is exact -- rows ordered by
    utils/metrics.py:73-86 euclidean_distance (float64), ties broken by row (FAISS and the
    reference's sorted() leave tie order unspecified).

    :return: (rows int64[Q, k'], dist float64[Q, k']), k' = min(k, N)
    """
    x_db = np.asarray(x_db, np.float64)
    q = np.atleast_2d(np.asarray(q, np.float64))
    kk = min(k, len(x_db))
    rows = np.empty((len(q), kk), np.int64)
    dist = np.empty((len(q), kk), np.float64)
    for i in range(len(q)):
        d = euclidean_distance(q[i], x_db) if len(x_db) else np.empty(0)
        o = np.lexsort((np.arange(len(d)), d))[:kk]
        rows[i], dist[i] = o, d[o]
    return rows, dist


DISTANCE_FUNCTIONS = {
    "euclidean": euclidean_distance,
    "cosine": cosine_distance,
    "hik": histogram_intersection_distance,
}


# --------------------------------------------------------------------------
# LSH index query  (impls/nn_index/lsh.py:452-519) on array form
# --------------------------------------------------------------------------
def unique_code_table(codes_words: np.ndarray):
    """The unique-code table + code->rows map that replaces the reference's
    ``set`` of ints (linear.py:163) and ``hash2uuids`` KVS (lsh.py:316-323).

    Table order is ascending integer value (= lexicographic word order), which
    is what the canonical tie-break of ``hamming_topk`` refers to.

    :return: (table uint32[U,W], row_of_code int64[N], csr_off int64[U+1],
              csr_rows int64[N])  with csr_rows listing descriptor rows of each
              code in ascending row order.
    """
    cw = np.ascontiguousarray(codes_words, dtype=np.uint32)
    keys = [cw[:, w] for w in range(cw.shape[1] - 1, -1, -1)]
    order = np.lexsort(keys)                       # stable: equal codes keep row order
    s = cw[order]
    new = np.ones(len(s), bool)
    if len(s) > 1:
        new[1:] = (s[1:] != s[:-1]).any(axis=1)
    table = s[new]
    code_id_sorted = np.cumsum(new) - 1
    inv = np.empty(len(s), np.int64)
    inv[order] = code_id_sorted
    off = np.concatenate([np.flatnonzero(new), [len(s)]]).astype(np.int64)
    return table, inv, off, order.astype(np.int64)


def lsh_nn(x_db: np.ndarray, codes_words: np.ndarray, q_vec: np.ndarray,
           q_words: np.ndarray, n: int, distance_method: str):
    """lsh.py:470-519 restated on arrays: n nearest *unique* codes (canonical
    order) -> all descriptor rows carrying those codes (code rank order, then
    row order) -> distances -> stable sort -> first n.

    :return: (rows int64[m], distances float64[m]), m <= n.
    """
    table, _, off, rows = unique_code_table(codes_words)
    _, near = hamming_topk(table, q_words, n)
    cand = np.concatenate([rows[off[c]:off[c + 1]] for c in near[0]])
    d = np.atleast_1d(DISTANCE_FUNCTIONS[distance_method](np.asarray(q_vec), x_db[cand]))
    o = np.argsort(d, kind="stable")[:n]
    return cand[o], d[o]
