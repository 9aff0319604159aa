#!/usr/bin/env python
"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Generates the golden fixtures under ``tests/golden/`` by importing the
UNMODIFIED reference (``/root/reference``, SMQTK-Indexing v0.18.0) and running
its own functions / classes.  The reference is pure Python, cannot travel to
the GPU box and needs three absent distributions (smqtk-core / -dataprovider /
-descriptors: plugin + container plumbing only, no arithmetic); they are
satisfied by the stand-ins in ``smqtk_indexing_b200/_compat/standins`` (which
pass the reference's own unit tests for this path: 108 passed, 4 skipped).

Run in the build container only:

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Inputs are produced from ``numpy.random.RandomState`` seeds recorded in each
fixture so tests can regenerate them bit-identically; outputs (and anything
that depends on LAPACK, e.g. fitted ITQ models) are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("SMQTK_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, os.path.join(ROOT, "smqtk_indexing_b200", "_compat", "standins"))
sys.path.insert(0, REFERENCE)

import warnings  # noqa: E402
warnings.filterwarnings("ignore")

from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore  # noqa: E402
from smqtk_descriptors.impls.descriptor_element.memory import DescriptorMemoryElement  # noqa: E402
from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet  # noqa: E402
from smqtk_indexing.impls.hash_index.linear import LinearHashIndex  # noqa: E402
from smqtk_indexing.impls.lsh_functor.itq import ItqFunctor  # noqa: E402
from smqtk_indexing.impls.nn_index.lsh import LSHNearestNeighborIndex  # noqa: E402
from smqtk_indexing.utils import bits as rbits  # noqa: E402
from smqtk_indexing.utils import metrics as rmetrics  # noqa: E402

sys.path.insert(0, HERE)
import golden_inputs as gi  # noqa: E402


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def int_to_hex(i):
    return format(i, "x")


# ---------------------------------------------------------------- bits
def gen_bits():
    rng = np.random.RandomState(101)
    lens, vecs, ints = [], [], []
    for b in [1, 2, 3, 5, 8, 31, 32, 33, 63, 64, 65, 100, 128, 255, 256, 257, 512, 1000, 1024]:
        for _ in range(4):
            v = rng.rand(b) > 0.5
            lens.append(b)
            pv = np.zeros(128, np.uint8)          # storage only; unpack with count=b
            pk = np.packbits(v)
            pv[:len(pk)] = pk
            vecs.append(pv)
            ints.append(int_to_hex(rbits.bit_vector_to_int_large(v)))
    # int -> vector, default width and padded width
    back_int, back_bits, back_vec = [], [], []
    for i, bits in [(0, 0), (0, 5), (1, 0), (5, 8), (2 ** 256 - 1, 0), (2 ** 256 - 1, 300),
                    (2 ** 512, 0), (2 ** 512, 513), (123456789123456789123456789, 100)]:
        back_int.append(int_to_hex(i))
        back_bits.append(bits)
        pv = np.zeros(128, np.uint8)
        pk = np.packbits(rbits.int_to_bit_vector_large(i, bits))
        pv[:len(pk)] = pk
        back_vec.append(pv)
    save("bits",
         lens=np.array(lens), vec_packed=np.array(vecs), ints_hex=np.array(ints),
         back_int_hex=np.array(back_int), back_bits=np.array(back_bits),
         back_len=np.array([len(rbits.int_to_bit_vector_large(int(h, 16), b))
                            for h, b in zip(back_int, back_bits)]),
         back_vec_packed=np.array(back_vec))


# ---------------------------------------------------------------- hamming distance + linear nn
def gen_hamming():
    rng = np.random.RandomState(202)
    out = {}
    for b in (64, 1024):
        a = [int.from_bytes(rng.bytes(b // 8), "big") for _ in range(200)]
        c = [int.from_bytes(rng.bytes(b // 8), "big") for _ in range(200)]
        out["a%d" % b] = np.array([int_to_hex(x) for x in a])
        out["b%d" % b] = np.array([int_to_hex(x) for x in c])
        out["d%d" % b] = np.array([rmetrics.hamming_distance(x, y) for x, y in zip(a, c)])
    save("hamming_distance", **out)


def gen_linear_nn():
    """LinearHashIndex.build_index + nn through the reference, on random codes.
    Stored: the db bit matrix seed, the reference's returned codes and
    distances for each (query, n)."""
    cases = []
    out = {}
    ci = 0
    for (b, U, seed) in gi.LINEAR_NN_CASES:
        db, qs = gi.linear_nn_inputs(b, U, seed)
        hi = LinearHashIndex()
        hi.build_index(db)
        for n in gi.linear_nn_ns(hi.count()):
            for qi in range(len(qs)):
                codes, dists = hi.nn(qs[qi], n)
                out["c%d_codes" % ci] = np.packbits(np.asarray(codes, bool), axis=1)
                out["c%d_dists" % ci] = np.asarray(dists, np.float64)
                cases.append((b, U, seed, qi, n, hi.count()))
                ci += 1
    out["cases"] = np.array(cases, np.int64)        # b, U, seed, query row, n, unique count
    save("linear_nn", **out)


# ---------------------------------------------------------------- ITQ
def _descr(mat, start=0):
    return [DescriptorMemoryElement(start + i).set_vector(mat[i]) for i in range(len(mat))]


def gen_itq():
    out = {}
    cases = []
    ci = 0
    # (N, D, bits, iterations, normalize, seed, dtype)
    for ci in range(len(gi.ITQ_CASES)):
        N, D, b, it, norm, seed, dt = gi.ITQ_CASES[ci]
        x, q = gi.itq_inputs(ci)
        f = ItqFunctor(bit_length=b, itq_iterations=it, normalize=norm, random_seed=seed)
        codes = f.fit(_descr(x))
        hq = np.array([f.get_hash(v) for v in q])
        zq = np.array([np.dot(f._norm_vector(v) - f.mean_vec, f.rotation) for v in q])
        out["c%d_mean" % ci] = f.mean_vec
        out["c%d_rot" % ci] = f.rotation
        out["c%d_fitcodes" % ci] = np.packbits(codes, axis=1)
        out["c%d_qbits" % ci] = np.packbits(hq, axis=1)
        out["c%d_qz" % ci] = zq
        cases.append((N, D, b, it, -1 if norm is None else norm, seed, 8 if dt == "f8" else 4))
    out["cases"] = np.array(cases, np.int64)
    save("itq", **out)


# ---------------------------------------------------------------- metrics
def gen_metrics():
    out = {}
    cases = []
    ci = 0
    for D in gi.METRIC_DIMS:
        q, c, qh, ch = gi.metrics_inputs(D)
        out["c%d_euclid" % ci] = np.array([rmetrics.euclidean_distance(q, r) for r in c])
        with np.errstate(all="ignore"):
            out["c%d_cosine" % ci] = np.array([rmetrics.cosine_distance(q, r) for r in c])
        out["c%d_hik" % ci] = np.array([rmetrics.histogram_intersection_distance_fast(qh, r) for r in ch])
        out["c%d_hik2d" % ci] = rmetrics.histogram_intersection_distance(qh, ch)
        cases.append(D)
        ci += 1
    out["cases"] = np.array(cases)
    save("metrics", **out)


# ---------------------------------------------------------------- LSH end to end
def gen_lsh():
    """tests/impls/nn_index/test_lsh.py:754-832 style: 1000x256 rand (seed 0),
    ITQ-32, each distance method, hash_index in {None, LinearHashIndex}."""
    out = {}
    cases = []
    ci = 0
    for ci in range(len(gi.LSH_CASES)):
        N, D, b, method, use_hi, seed = gi.LSH_CASES[ci]
        x, qs = gi.lsh_inputs(ci)
        td = _descr(x)
        f = ItqFunctor(bit_length=b, random_seed=seed)
        f.fit(td)
        idx = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                      hash_index=LinearHashIndex() if use_hi else None,
                                      distance_method=method)
        idx.build_index(td)
        for n in gi.LSH_NS:
            for qi in range(len(qs)):
                r, d = idx.nn(DescriptorMemoryElement(10 ** 6 + qi).set_vector(qs[qi]), n)
                out["c%d_q%d_n%d_uuids" % (ci, qi, n)] = np.array([e.uuid() for e in r], np.int64)
                out["c%d_q%d_n%d_dists" % (ci, qi, n)] = np.array(d, np.float64)
        out["c%d_mean" % ci] = f.mean_vec
        out["c%d_rot" % ci] = f.rotation
        out["c%d_count" % ci] = np.array(idx.count())
        out["c%d_nkeys" % ci] = np.array(idx.hash2uuids_kvstore.count())
        cases.append((N, D, b, {"euclidean": 0, "cosine": 1, "hik": 2}[method], use_hi, seed))
    out["cases"] = np.array(cases, np.int64)
    save("lsh_nn", **out)


# ---------------------------------------------------------------- LSH incremental ingest
def gen_lsh_ingest():
    """State the UNMODIFIED reference leaves after a build + update_index / remove_from_index
    sequence (lsh.py:283-450): the hash2uuids KVS, the LinearHashIndex content, count(), and nn()
    answers -- what the device-speed *_matrix ingest of the product must reproduce."""
    N, D, b, seed = gi.INGEST_SHAPE
    x, qs = gi.ingest_inputs()
    f = ItqFunctor(bit_length=b, random_seed=seed, itq_iterations=20)
    f.fit(_descr(x[:400]))
    # the fixture is only meaningful if no projection sits at rounding level (fp32 device hashing)
    z = np.dot(np.vstack([x, qs]).astype(np.float64) - f.mean_vec, f.rotation)
    assert np.abs(z).min() > 5e-5, "pick another seed: a projection is within 5e-5 of zero"
    idx = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                  hash_index=LinearHashIndex(), distance_method="euclidean")
    out = {"mean": f.mean_vec, "rot": f.rotation}
    live = set()
    for si, step in enumerate(gi.INGEST_STEPS):
        op, rows = gi.ingest_step_rows(step)
        if op == "add":
            (idx.build_index if si == 0 else idx.update_index)([DescriptorMemoryElement(r).set_vector(x[r]) for r in rows])
            live |= set(rows)
        else:
            idx.remove_from_index(rows)
            live -= set(rows)
        kvs = idx.hash2uuids_kvstore
        keys = sorted(kvs.keys())
        off, uu = [0], []
        for k in keys:
            uu.extend(sorted(kvs.get(k)))
            off.append(len(uu))
        assert sorted(uu) == sorted(live)
        out["s%d_keys_hex" % si] = np.array([int_to_hex(k) for k in keys])
        out["s%d_off" % si] = np.array(off, np.int64)
        out["s%d_uuids" % si] = np.array(uu, np.int64)
        out["s%d_hash_index_hex" % si] = np.array(sorted(int_to_hex(k) for k in idx.hash_index.index))
        out["s%d_count" % si] = np.array(idx.count())
        for qi in range(len(qs)):
            r, d = idx.nn(DescriptorMemoryElement(10 ** 6 + qi).set_vector(qs[qi]), gi.INGEST_N)
            out["s%d_q%d_uuids" % (si, qi)] = np.array([e.uuid() for e in r], np.int64)
            out["s%d_q%d_dists" % (si, qi)] = np.array(d, np.float64)
    save("lsh_ingest", **out)


if __name__ == "__main__":
    gen_bits()
    gen_hamming()
    gen_linear_nn()
    gen_itq()
    gen_metrics()
    gen_lsh()
    gen_lsh_ingest()
