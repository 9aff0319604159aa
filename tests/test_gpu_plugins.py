"""
GPU tests of the plugin classes through the reference-facing API, covering the
behaviours the reference's own tests pin (tests/impls/hash_index/test_linear.py,
tests/impls/lsh_functor/test_itq.py, tests/impls/nn_index/test_lsh.py) plus
golden comparisons against outputs of the unmodified reference.
"""
import random
from io import BytesIO
from math import sqrt

import numpy as np
import pytest

import golden_inputs as gi
import np_oracle as O

from smqtk_dataprovider.exceptions import ReadOnlyError
from smqtk_dataprovider.impls.data_element.memory import DataMemoryElement
from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
from smqtk_descriptors.impls.descriptor_element.memory import DescriptorMemoryElement
from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet

from smqtk_indexing_b200.interfaces import LshFunctor
from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex
from smqtk_indexing_b200.utils import bits as bitutil
from smqtk_indexing_b200.utils import metrics

pytestmark = pytest.mark.gpu

B4 = [[0, 1, 0], [1, 0, 0], [0, 1, 1], [0, 0, 1]]


def _descr(mat, start=0):
    return [DescriptorMemoryElement(start + i).set_vector(mat[i]) for i in range(len(mat))]


# =========================================================================== LinearHashIndex
def test_linear_build_update_remove():
    i = LinearHashIndex()
    i.build_index(B4)
    assert i.index == {1, 2, 3, 4} and i.count() == 4 and i.cache_element is None
    i.build_index([[0, 1, 0], [1, 0, 0]])                 # rebuild replaces
    assert i.index == {2, 4}
    i.update_index([[0, 1, 1], [0, 0, 1], [0, 1, 0]])     # union, duplicates collapse
    assert i.index == {1, 2, 3, 4}
    i.remove_from_index([[0, 1, 1], [1, 0, 0]])
    assert i.index == {1, 2}
    with pytest.raises(KeyError):                          # unknown code: nothing changes
        i.remove_from_index([[0, 0, 1], [1, 1, 1]])
    assert i.index == {1, 2}
    i.remove_from_index([[0, 0, 1], [0, 1, 0]])
    assert i.count() == 0 and i.index == set()
    e = LinearHashIndex()
    e.update_index(B4)                                     # update on an empty index builds it
    assert e.index == {1, 2, 3, 4}


def test_linear_nn_known_answer():
    # reference tests/impls/hash_index/test_linear.py:141-155
    i = LinearHashIndex()
    i.build_index([[0, 1, 0], [1, 1, 0], [0, 1, 1], [0, 0, 1]])
    near_codes, near_dists = i.nn([0, 0, 0], 4)
    assert near_codes.dtype == bool and near_codes.shape == (4, 3)
    assert set(map(tuple, near_codes[:2])) == {(0, 1, 0), (0, 0, 1)}
    assert set(map(tuple, near_codes[2:])) == {(1, 1, 0), (0, 1, 1)}
    np.testing.assert_array_almost_equal(near_dists, (1 / 3., 1 / 3., 2 / 3., 2 / 3.))
    # canonical tie order: ascending code value
    assert [tuple(c) for c in near_codes] == [(0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 1, 0)]
    codes, dists = i.nn([0, 0, 0], 10)                     # n > count
    assert len(codes) == 4 and len(dists) == 4


def test_linear_cache_save_load():
    ce = DataMemoryElement()
    i = LinearHashIndex(ce)
    assert ce.is_empty()
    i.build_index([[0, 1, 0], [1, 0, 0]])
    assert set(np.load(BytesIO(ce.get_bytes()))) == {2, 4}
    i.update_index([[0, 1, 1], [0, 0, 1]])
    assert set(np.load(BytesIO(ce.get_bytes()))) == {1, 2, 3, 4}
    i.remove_from_index([[0, 1, 1], [1, 0, 0]])
    assert set(np.load(BytesIO(ce.get_bytes()))) == {1, 2}
    j = LinearHashIndex(ce)                                # load on construction
    assert j.index == i.index
    ro = LinearHashIndex(DataMemoryElement(readonly=True))
    with pytest.raises(ValueError, match="is read-only"):
        ro.build_index(B4)
    with pytest.raises(ValueError, match="is read-only"):
        ro.update_index(B4)


def test_linear_nn_golden(golden):
    """against the reference's LinearHashIndex.nn outputs (tie groups, exact distances)"""
    g = golden("linear_nn")
    built = {}
    for ci, (b, U, seed, qi, n, ucount) in enumerate(g["cases"]):
        key = (int(b), int(U), int(seed))
        if key not in built:
            db, qs = gi.linear_nn_inputs(*key)
            hi = LinearHashIndex()
            hi.build_index(db)
            built[key] = (hi, qs)
        hi, qs = built[key]
        assert hi.count() == ucount
        codes, dists = hi.nn(qs[qi], int(n))
        ref_codes = np.unpackbits(g["c%d_codes" % ci], axis=1, count=int(b)).astype(bool)
        ref_d = g["c%d_dists" % ci]
        assert np.array_equal(np.asarray(dists), ref_d)
        boundary = ref_d[-1]
        mine_in = {tuple(c) for c, d in zip(codes, dists) if d < boundary}
        ref_in = {tuple(c) for c, d in zip(ref_codes, ref_d) if d < boundary}
        assert mine_in == ref_in
        assert len({tuple(c) for c in codes}) == len(codes)


def test_linear_wide_and_mixed_width_codes():
    rng = np.random.RandomState(0)
    db = rng.rand(500, 300) > 0.5
    hi = LinearHashIndex()
    hi.build_index(db)
    codes, dists = hi.nn(db[17], 3)
    assert np.array_equal(codes[0], db[17]) and dists[0] == 0.0
    W = 16
    t, _, _, _ = O.unique_code_table(O.pack_codes(db, W))
    od, oi = O.hamming_topk(t, O.pack_codes(db[17], W), 3)
    assert np.array_equal(np.asarray(dists) * 300, od[0])
    # ints set through the attribute (LSH on-the-fly path), queried with an explicit bit length
    h2 = LinearHashIndex()
    h2.index = {0b0001, 0b0110, 0b1111}
    c, d = h2.nn([0, 1, 1, 1], 2)
    assert [tuple(x) for x in c] == [(False, True, True, False), (True, True, True, True)]
    assert d == (0.25, 0.25)


# =========================================================================== ItqFunctor
def test_itq_get_hash_known_answer():
    # reference tests/impls/lsh_functor/test_itq.py:304-336
    itq = ItqFunctor(bit_length=1, random_seed=0)
    itq.mean_vec = np.array([0., 0.])
    itq.rotation = np.array([[1. / sqrt(2)], [1. / sqrt(2)]])
    for p, want in [([1, 1], True), ([-1, -1], False), ([-1, 1], True), ([-1.001, 1], False),
                    ([-1, 1.001], True), ([1, -1], True), ([1, -1.001], False), ([1.001, -1], True)]:
        h = itq.get_hash(np.array(p))
        assert h.dtype == bool and h.shape == (1,)
        np.testing.assert_array_equal(h, [want])
    assert np.array_equal(itq(np.array([1, 1])), [True])
    assert itq.get_hash(np.array([[1, 1], [-1, -1]])).tolist() == [[True], [False]]


def test_itq_fit_known_answer_and_cache():
    # reference tests/impls/lsh_functor/test_itq.py:255-302
    fit_descriptors = _descr([[-2. + i, -2. + i] for i in range(5)])
    itq = ItqFunctor(DataMemoryElement(), DataMemoryElement(), bit_length=1, random_seed=0)
    codes = itq.fit(fit_descriptors)
    assert codes.shape == (5, 1) and codes.dtype == bool
    np.testing.assert_array_almost_equal(itq.mean_vec, [0, 0])
    np.testing.assert_array_almost_equal(itq.rotation, [[1 / sqrt(2)], [1 / sqrt(2)]])
    np.testing.assert_array_almost_equal(np.load(BytesIO(itq.mean_vec_cache_elem.get_bytes())), [0, 0])
    np.testing.assert_array_almost_equal(np.load(BytesIO(itq.rotation_cache_elem.get_bytes())),
                                         [[1 / sqrt(2)], [1 / sqrt(2)]])
    with pytest.raises(RuntimeError):
        itq.fit(fit_descriptors)
    itq2 = ItqFunctor(bit_length=1, random_seed=0)
    itq2.fit(iter(fit_descriptors))                       # iterator input
    np.testing.assert_array_almost_equal(itq2.rotation, itq.rotation)


def _quant_loss(x, mean, rot, normalize):
    """ITQ objective |sign(z) - z|^2 / N for a model (lower is better)."""
    z = O.itq_project(x.astype(np.float64), mean, rot, normalize)
    return float(((np.where(z >= 0, 1.0, -1.0) - z) ** 2).sum() / len(x))


#: fit parity (GPU FP64 contractions + host LAPACK vs the reference on this
#: container's LAPACK).  ITQ's sign step amplifies 1e-16 summation-order
#: differences chaotically over many iterations, so beyond ~20 iterations the
#: rotation itself is not reproducible (SURVEY section 7 hard part 4); what is pinned:
#:   * mean: 1e-6 abs (f32 inputs: the reference averages in f32);
#:   * PCA stage: projector onto span(rotation) equals the reference's, 1e-6 abs;
#:   * rotation orthonormal to 1e-10;
#:   * <= 20 iterations and D <= 32: rotation 1e-5 abs, training codes >= 99.9 % identical;
#:   * (iteration numerics from a shared projection: test_itq_rotation_iterations_match_oracle);
#:   * otherwise: quantisation loss within 2 % of the reference model's.
def test_itq_fit_golden(golden):
    g = golden("itq")
    for ci, (N, D, b, it, norm, seed, isz) in enumerate(g["cases"]):
        x, q = gi.itq_inputs(ci)
        normalize = None if norm < 0 else int(norm)
        f = ItqFunctor(bit_length=int(b), itq_iterations=int(it), normalize=normalize, random_seed=int(seed))
        codes = f.fit(_descr(x))
        assert f.mean_vec.dtype == x.dtype and f.rotation.dtype == np.float64
        mean_ref, rot_ref = g["c%d_mean" % ci], np.real(g["c%d_rot" % ci])
        np.testing.assert_allclose(f.mean_vec, mean_ref, rtol=0, atol=1e-6)
        np.testing.assert_allclose(f.rotation @ f.rotation.T, rot_ref @ rot_ref.T, rtol=0, atol=1e-6)
        np.testing.assert_allclose(f.rotation.T @ f.rotation, np.eye(int(b)), rtol=0, atol=1e-10)
        ref_codes = np.unpackbits(g["c%d_fitcodes" % ci], axis=1, count=int(b)).astype(bool)
        if it <= 20 and D <= 32:
            # (well separated spectra; for a near-degenerate covariance even the
            # eigenvector basis inside the PCA subspace is LAPACK-build dependent)
            np.testing.assert_allclose(f.rotation, rot_ref, rtol=0, atol=1e-5)
            assert (codes != ref_codes).mean() <= 1e-3
            ref_bits = np.unpackbits(g["c%d_qbits" % ci], axis=1, count=int(b)).astype(bool)
            mine = np.array([f.get_hash(v) for v in q])
            far = np.abs(g["c%d_qz" % ci]) > 1e-4
            assert np.array_equal(mine[far], ref_bits[far])
        else:
            mine = _quant_loss(x, f.mean_vec.astype(np.float64), f.rotation, normalize)
            ref = _quant_loss(x, mean_ref.astype(np.float64), rot_ref, normalize)
            assert abs(mine - ref) <= 0.02 * ref, (mine, ref)


def test_itq_rotation_iterations_match_oracle():
    """The ITQ iteration itself (Z = V.R, sign, C = UX^T.V on the GPU, SVD on the
    host, R = Vh.U^T) from a SHARED projection V and seed: rotation after 1, 3 and
    10 iterations equals the oracle's (which is bit-pinned to the reference)."""
    import torch
    from smqtk_indexing_b200 import fit as fitops
    from smqtk_indexing_b200.utils.bits import unpack_bits
    rng = np.random.RandomState(5)
    for n, b in ((2000, 64), (777, 5), (4000, 32), (300, 100)):
        v = rng.randn(n, b) * (1 + np.arange(b))[None, :] * 0.1
        vd = torch.from_numpy(v).cuda()
        for iters in (1, 3, 10):
            codes, r = fitops.itq_rotation(vd, iters, 7)
            oc, orr = O.find_itq_rotation(v, iters, 7)
            np.testing.assert_allclose(r, orr, rtol=0, atol=1e-8)
            mine = unpack_bits(codes.cpu().numpy().view(np.uint32), b)
            z = v @ orr
            far = np.abs(z) > 1e-9
            assert np.array_equal(mine[far], oc[far])


def test_itq_fit_stages_match_oracle():
    """Stage-wise parity of the N-scaled contractions against numpy on the same
    data: row norms, column mean, covariance, projection (all FP64 on the GPU)."""
    import torch
    from smqtk_indexing_b200 import fit as fitops
    rng = np.random.RandomState(6)
    for dtype in (np.float64, np.float32):
        for normalize in (None, 2, 1, np.inf):
            x = rng.rand(1500, 48).astype(dtype)
            xd = torch.from_numpy(x).cuda()
            div = fitops.row_div(xd, normalize)
            xn = O.norm_vector(x.astype(np.float64), normalize)
            if div is not None:
                np.testing.assert_allclose(div.cpu().numpy(), np.linalg.norm(x.astype(np.float64), normalize, axis=1),
                                           rtol=1e-14)
            mean = fitops.col_mean(xd, div)
            np.testing.assert_allclose(mean.cpu().numpy(), xn.mean(0), rtol=1e-13)
            cov = fitops.gram(xd, xd, scale=1.0 / (len(x) - 1), a_div=div, a_mean=mean, b_div=div, b_mean=mean)
            np.testing.assert_allclose(cov.cpu().numpy(), np.cov((xn - xn.mean(0)).T), rtol=1e-11, atol=1e-15)
            pc = np.linalg.qr(rng.randn(48, 48))[0][:, :20]
            v, codes = fitops.project(xd, torch.from_numpy(pc).cuda(), a_div=div, a_mean=mean, want_codes=True)
            vref = (xn - xn.mean(0)) @ pc
            np.testing.assert_allclose(v.cpu().numpy(), vref, rtol=0, atol=1e-13)


def test_itq_get_hash_golden_models(golden):
    """reference-fitted model loaded into the functor: bits equal except |z| < eps"""
    g = golden("itq")
    for ci, (N, D, b, it, norm, seed, isz) in enumerate(g["cases"]):
        _, q = gi.itq_inputs(ci)
        f = ItqFunctor(bit_length=int(b), normalize=None if norm < 0 else int(norm))
        f.mean_vec, f.rotation = g["c%d_mean" % ci], g["c%d_rot" % ci]
        ref_bits = np.unpackbits(g["c%d_qbits" % ci], axis=1, count=int(b)).astype(bool)
        z = g["c%d_qz" % ci]
        a = O.norm_vector(q.astype(np.float64), f.normalize) - np.real(f.mean_vec)
        eps = 1e-5 * np.linalg.norm(a, axis=1)[:, None] * np.linalg.norm(np.real(f.rotation), axis=0)[None, :]
        mine = f.get_hash(q)
        bad = mine != ref_bits
        assert (np.abs(z[bad]) <= eps[bad]).all()


# =========================================================================== metrics
def test_metric_functions_known_answers():
    # reference tests/impls/nn_index/test_lsh.py:102-136 and tests/utils/test_metrics.py
    e = LSHNearestNeighborIndex._get_dist_func('euclidean')
    c = LSHNearestNeighborIndex._get_dist_func('cosine')
    h = LSHNearestNeighborIndex._get_dist_func('hik')
    a = np.array
    assert abs(e(a([0, 0]), a([0, 1])) - 1.0) < 1e-12
    assert abs(c(a([1, 0]), a([0, 1])) - 1.0) < 1e-7
    assert abs(c(a([1, 0]), a([1, 1])) - 0.5) < 1e-7
    assert h(a([0, 0]), a([0, 1])) == 1.0 and h(a([1, 0]), a([0, 1])) == 1.0 and h(a([1, 1]), a([0, 1])) == 0.0
    hd = metrics.histogram_intersection_distance
    assert hd(a([.5, .5]), a([.5, .5])) == 0.0
    np.testing.assert_array_equal(hd(a([.5, .5]), a([[0, 1], [1, 0], [.5, .5], [0, 0]])), [.5, .5, 0., 1.])
    np.testing.assert_array_equal(hd(a([[0, 1], [1, 0], [.5, .5], [0, 0]]), a([.5, .5])), [.5, .5, 0., 1.])
    np.testing.assert_array_equal(hd(a([[0, 1], [1, 0]]), a([[0, 1], [0, 1]])), [0., 1.])
    with pytest.raises(ValueError):
        hd(a([[0, 1], [1, 0]]), a([[0, 1]]))
    rng = random.Random(0)
    for bits in (64, 1024):
        for _ in range(20):
            x, y = rng.getrandbits(bits), rng.getrandbits(bits)
            assert metrics.hamming_distance(x, y) == bin(x ^ y).count('1')


# =========================================================================== LSH index
class DummyHashFunctor(LshFunctor):
    """bits of int(sum(v)) -- predictable codes for state tests (a foreign,
    host-side functor: also exercises the non-batch hashing path)."""

    @classmethod
    def is_usable(cls):
        return True

    def get_config(self):
        return {}

    def get_hash(self, descriptor):
        return np.asarray([int(c) for c in bin(int(descriptor.sum()))[2:]], bool)


def _dummy_index(**kw):
    return LSHNearestNeighborIndex(DummyHashFunctor(), MemoryDescriptorSet(), MemoryKeyValueStore(), **kw)


@pytest.mark.parametrize("hi", [None, "linear"])
def test_lsh_build_update_remove_state(hi):
    idx = _dummy_index(hash_index=LinearHashIndex() if hi else None)
    d = _descr([[float(i)] * 2 for i in range(1, 6)])       # sums 2,4,6,8,10
    d.append(DescriptorMemoryElement(5).set_vector([1., 1.]))  # same hash as uuid 0
    idx.build_index(d)
    assert idx.count() == 6 and idx.descriptor_set.count() == 6
    assert set(idx.hash2uuids_kvstore.keys()) == {2, 4, 6, 8, 10}
    assert idx.hash2uuids_kvstore.get(2) == {0, 5}
    if hi:
        assert idx.hash_index.index == {2, 4, 6, 8, 10}
    # rebuild replaces everything
    idx.build_index(d[:2])
    assert idx.count() == 2 and set(idx.hash2uuids_kvstore.keys()) == {2, 4}
    if hi:
        assert idx.hash_index.index == {2, 4}
    # additive update, duplicate uuids are idempotent
    idx.update_index(d[2:])
    idx.update_index(d[2:4])
    assert idx.count() == 6 and idx.hash2uuids_kvstore.get(2) == {0, 5}
    # removing one of two uuids sharing a hash keeps the code indexed
    idx.remove_from_index([5])
    assert idx.hash2uuids_kvstore.get(2) == {0} and idx.descriptor_set.count() == 5
    if hi:
        assert 2 in idx.hash_index.index
    idx.remove_from_index([0, 1])
    assert set(idx.hash2uuids_kvstore.keys()) == {6, 8, 10}
    if hi:
        assert idx.hash_index.index == {6, 8, 10}
    # unknown uid: KeyError and nothing changes
    before = (idx.count(), set(idx.hash2uuids_kvstore.keys()), idx.descriptor_set.count())
    with pytest.raises(KeyError):
        idx.remove_from_index([2, 99])
    assert before == (idx.count(), set(idx.hash2uuids_kvstore.keys()), idx.descriptor_set.count())
    # update on an empty index builds it
    e = _dummy_index(hash_index=LinearHashIndex() if hi else None)
    e.update_index(d[:3])
    assert e.count() == 3


def test_lsh_read_only():
    idx = _dummy_index(read_only=True)
    d = _descr([[1., 1.]])
    for call, arg in ((idx.build_index, d), (idx.update_index, d), (idx.remove_from_index, [0])):
        with pytest.raises(ReadOnlyError):
            call(arg)


def _itq(bits, seed=0):
    return ItqFunctor(bit_length=bits, random_seed=seed)


@pytest.mark.parametrize("use_hi", [False, True])
def test_lsh_random_euclidean(use_hi):
    # reference tests/impls/nn_index/test_lsh.py:754-832
    n, dim = 1000, 256
    np.random.seed(0)
    td = _descr([np.random.rand(dim) for _ in range(n)])
    f = _itq(32)
    f.fit(td)
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                    hash_index=LinearHashIndex() if use_hi else None,
                                    distance_method='euclidean')
    index.build_index(td)
    q = td[255]
    r, dists = index.nn(q, 1)
    assert r[0] == q and r[0].uuid() == 255
    v = td[0].vector().copy()
    v_min = max(v.min(), 0.1)
    v[0] += v_min
    v[dim - 1] -= v_min
    r, dists = index.nn(DescriptorMemoryElement(n).set_vector(v), 1)
    assert r[0] == td[0]
    qv = DescriptorMemoryElement(n + 1).set_vector(np.random.rand(dim))
    for k in (10, n):
        r, dists = index.nn(qv, k)
        assert len(r) == len(dists) <= k
        assert all(dists[j] > dists[j - 1] for j in range(1, len(dists)))


@pytest.mark.parametrize("use_hi", [False, True])
@pytest.mark.parametrize("method", ["euclidean", "hik"])
def test_lsh_known_unit(use_hi, method):
    # reference tests/impls/nn_index/test_lsh.py:837-919
    dim = 5
    td = _descr(np.eye(dim))
    f = _itq(5)
    f.fit(td)
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                    hash_index=LinearHashIndex() if use_hi else None, distance_method=method)
    index.build_index(td)
    r, dists = index.nn(DescriptorMemoryElement(0).set_vector(np.zeros(dim, float)), dim)
    assert len(dists) == dim and all(d == 1. for d in dists)
    r, dists = index.nn(td[3], 1)
    assert r[0] == td[3] and dists[0] == 0.
    r, dists = index.nn(td[3], dim)
    assert r[0] == td[3] and dists[0] == 0.


@pytest.mark.parametrize("use_hi", [False, True])
def test_lsh_known_ordered_euclidean(use_hi):
    # reference tests/impls/nn_index/test_lsh.py:924-979 (1-bit codes: heavy collisions)
    n = 1000
    td = _descr([np.array([j, j * 2], float) for j in range(n)])
    random.Random(1).shuffle(td)
    f = _itq(1)
    f.fit(td)
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                    hash_index=LinearHashIndex() if use_hi else None,
                                    distance_method='euclidean')
    index.build_index(td)
    q = DescriptorMemoryElement(n).set_vector(np.array([0, 0], float))
    r, dists = index.nn(q, 5)
    assert [d.uuid() for d in r] == [0, 1, 2, 3, 4]
    r, dists = index.nn(q, n)
    assert [d.uuid() for d in r] == list(range(n))
    rows, bd = index.nn_batch(np.array([[0., 0.]]), n)
    uu = index.mirror_uuids()
    assert [uu[i] for i in rows[0]] == list(range(n))
    np.testing.assert_allclose(bd[0], dists, rtol=1e-12)


def test_lsh_nn_golden(golden):
    """nn() and nn_batch() with the reference-fitted model against the
    reference's own LSHNearestNeighborIndex.nn outputs."""
    g = golden("lsh_nn")
    methods = ["euclidean", "cosine", "hik"]
    for ci, (N, D, b, mi, use_hi, seed) in enumerate(g["cases"]):
        x, qs = gi.lsh_inputs(ci)
        f = ItqFunctor(bit_length=int(b))
        f.mean_vec, f.rotation = g["c%d_mean" % ci], g["c%d_rot" % ci]
        index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(),
                                        hash_index=LinearHashIndex() if use_hi else None,
                                        distance_method=methods[mi])
        index.build_index(_descr(x))
        assert index.count() == g["c%d_count" % ci]
        assert index.hash2uuids_kvstore.count() == g["c%d_nkeys" % ci]
        W = [w for w in (1, 2, 4, 8) if w * 32 >= b][0]
        table, _, _, _ = O.unique_code_table(O.pack_codes(O.itq_hash(x, np.real(f.mean_vec), np.real(f.rotation)), W))
        uu = index.mirror_uuids()
        for n in gi.LSH_NS:
            rows_b, dists_b = index.nn_batch(qs, n)
            for qi in range(len(qs)):
                ref_u = g["c%d_q%d_n%d_uuids" % (ci, qi, n)]
                ref_d = g["c%d_q%d_n%d_dists" % (ci, qi, n)]
                qw = O.pack_codes(O.itq_hash(qs[qi], np.real(f.mean_vec), np.real(f.rotation)), W)
                srt = np.sort(O.hamming_distances(table, qw[0]))
                if n < len(table) and srt[n - 1] == srt[n]:
                    continue            # candidate pool depends on the reference's unspecified tie order
                r, d = index.nn(DescriptorMemoryElement(10 ** 6).set_vector(qs[qi]), n)
                # descriptors are fp32 on the device: 1e-5 relative (+ cosine conditioning floor)
                np.testing.assert_allclose(d, ref_d, rtol=1e-5, atol=5e-8)
                keep = rows_b[qi] >= 0
                np.testing.assert_allclose(dists_b[qi][keep], ref_d, rtol=1e-5, atol=5e-8)
                gaps = np.diff(ref_d)
                if len(ref_d) == 1 or gaps.min() > 1e-4 * ref_d.max():
                    assert [e.uuid() for e in r] == list(ref_u)
                    assert [uu[i] for i in rows_b[qi][keep]] == list(ref_u)


def test_lsh_nn_batch_matches_oracle_pipeline():
    """device pipeline vs the oracle's array-form LSH query (canonical ties)."""
    rng = np.random.RandomState(3)
    N, D, b = 3000, 64, 12
    x = rng.rand(N, D).astype(np.float32)
    f = ItqFunctor(bit_length=b, random_seed=1, itq_iterations=10)
    f.fit_matrix(x)
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), 'euclidean')
    index.build_index(_descr(x))
    qs = rng.rand(20, D).astype(np.float32)
    rows, dists = index.nn_batch(qs, 15)
    uu = index.mirror_uuids()
    xs = np.array([index.descriptor_set.get_descriptor(u).vector() for u in uu])
    codes = O.pack_codes(f.get_hash(xs), 1)
    for qi in range(len(qs)):
        qw = O.pack_codes(f.get_hash(qs[qi]), 1)
        orows, od = O.lsh_nn(xs.astype(np.float64), codes, qs[qi].astype(np.float64), qw, 15, "euclidean")
        np.testing.assert_allclose(dists[qi][:len(od)], od, rtol=1e-5, atol=1e-9)
        if np.diff(od).min() > 1e-6:
            assert list(rows[qi][:len(od)]) == list(orows)


# --------------------------------------------------------------------- flat L2 index (row N1)
def test_flat_l2_index_build_update_remove_nn():
    """State semantics of the reference's flat index (faiss.py:486-679) + exact neighbours."""
    from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
    rng = np.random.RandomState(11)
    x = rng.rand(600, 32)
    idx = FlatL2NearestNeighborsIndex(MemoryDescriptorSet())
    idx.build_index(_descr(x[:400]))
    assert idx.count() == 400 and idx._descriptor_set.count() == 400
    # self query, known order
    r, d = idx.nn(DescriptorMemoryElement(9000).set_vector(x[5]), 3)
    assert r[0].uuid() == 5 and d[0] == 0.0 and list(d) == sorted(d)
    orow, od = O.l2_topk(x[:400].astype(np.float32), x[5].astype(np.float32), 3)
    assert [e.uuid() for e in r] == list(orow[0])
    np.testing.assert_allclose(d, od[0], rtol=1e-5)
    # update: new rows appended, an existing uuid overwritten in place
    changed = DescriptorMemoryElement(7).set_vector(x[500])
    idx.update_index(_descr(x[400:], start=400) + [changed])
    assert idx.count() == 600
    truth = x.copy()
    truth[7] = x[500]
    qs = rng.rand(5, 32).astype(np.float32)
    rows, dists = idx.nn_batch(qs, 10)
    orow, od = O.l2_topk(truth.astype(np.float32), qs, 10)
    uu = idx.row_uuids()
    assert [[uu[i] for i in row] for row in rows] == [list(o) for o in orow]
    np.testing.assert_allclose(dists, od, rtol=1e-5)
    # remove: KeyError is non-mutating, then rows vanish from results
    with pytest.raises(KeyError):
        idx.remove_from_index([3, 99999])
    assert idx.count() == 600
    near = [int(u) for u in orow[0][:4]]
    idx.remove_from_index(near)
    assert idx.count() == 596 and idx._descriptor_set.count() == 596
    r, d = idx.nn(DescriptorMemoryElement(9001).set_vector(qs[0]), 6)
    assert [e.uuid() for e in r] == list(orow[0][4:10])
    # n larger than the index
    r, d = idx.nn(DescriptorMemoryElement(9002).set_vector(qs[1]), 1000)
    assert len(r) == 596 and list(d) == sorted(d)
    idx.remove_from_index([u for u in idx.row_uuids()])
    assert idx.count() == 0
    with pytest.raises(ValueError):
        idx.nn(DescriptorMemoryElement(9003).set_vector(qs[1]), 1)


def test_flat_l2_index_matrix_build_and_batch():
    from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
    rng = np.random.RandomState(12)
    x = rng.rand(30000, 128).astype(np.float32)
    idx = FlatL2NearestNeighborsIndex(MemoryDescriptorSet())
    idx.build_index_matrix(x)
    qs = rng.rand(40, 128).astype(np.float32)
    rows, dists = idx.nn_batch(qs, 100)
    orow, od = O.l2_topk(x, qs, 100)
    assert np.array_equal(rows, orow)
    np.testing.assert_allclose(dists, od, rtol=1e-9)


def test_flat_l2_matrix_built_index_supports_update_remove_without_touching_the_callers_tensor():
    """A matrix-built index (default uuids = row numbers) takes the element API afterwards: remove by
    row-number uuid, re-adding a uuid overwrites (no duplicate row), KeyError before mutation -- and the
    caller's CUDA tensor, adopted without a copy, is never written (copy-on-write)."""
    import torch
    from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
    rng = np.random.RandomState(13)
    x = rng.rand(5000, 32).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    keep_copy = xt.clone()
    idx = FlatL2NearestNeighborsIndex(MemoryDescriptorSet())
    idx.build_index_matrix(xt)
    new_vec = rng.rand(32)
    idx.update_index([DescriptorMemoryElement(7).set_vector(new_vec), DescriptorMemoryElement(5000).set_vector(x[0] + 1.0)])
    assert idx.count() == 5001                                   # uuid 7 overwritten, uuid 5000 appended
    assert torch.equal(xt, keep_copy)                            # the caller's matrix is untouched
    rows, dists = idx.nn_batch(new_vec.astype(np.float32)[None, :], 1)
    assert idx.row_uuids()[rows[0][0]] == 7 and dists[0][0] < 1e-6
    idx.remove_from_index([3, 7])
    assert idx.count() == 4999 and 7 not in idx.row_uuids() and 3 not in idx.row_uuids()
    with pytest.raises(KeyError):
        idx.remove_from_index([10, 3])
    assert idx.count() == 4999
    want = np.delete(np.vstack([x, x[:1] + 1.0]), [3, 7], axis=0)
    qs = rng.rand(9, 32).astype(np.float32)
    rows, dists = idx.nn_batch(qs, 5)
    orow, od = O.l2_topk(want, qs, 5)
    assert np.array_equal(rows, orow)
    np.testing.assert_allclose(dists, od, rtol=1e-9)


def test_flat_l2_unaligned_dimension_large_table():
    """D % 16 != 0 takes the brute-force path (re-rank kernel over every row + selection);
    at 1e5+ rows the selection must not be the O(m^2) rank count."""
    from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
    rng = np.random.RandomState(14)
    x = rng.rand(120_000, 30).astype(np.float32)
    idx = FlatL2NearestNeighborsIndex(MemoryDescriptorSet())
    idx.build_index_matrix(x)
    qs = rng.rand(6, 30).astype(np.float32)
    qs[0] = x[77]
    rows, dists = idx.nn_batch(qs, 20)
    orow, od = O.l2_topk(x, qs, 20)
    assert np.array_equal(rows, orow)
    np.testing.assert_allclose(dists, od, rtol=1e-9)


def test_simple_rp_functor_bits_and_lsh_use():
    """N4: random projections through the hash kernels; bits = (v - mean) . rps >= 0 for any
    `normalize` (simple_rp.py:52-59 divides rows by a positive scalar)."""
    from smqtk_indexing_b200.impls.lsh_functor.simple_rp import SimpleRPFunctor
    rng = np.random.RandomState(6)
    x = rng.rand(700, 48)
    for normalize in (None, 2):
        f = SimpleRPFunctor(bit_length=40, normalize=normalize, random_seed=5)
        codes = f.fit(_descr(x))
        np.random.seed(5)
        rps = np.random.randn(48, 40)
        assert np.array_equal(f.rps, rps)
        z = (x - x.mean(0)).dot(rps)
        far = np.abs(z) > 1e-4
        assert codes.shape == (700, 40) and np.array_equal(codes[far], (z >= 0)[far])
        with pytest.raises(RuntimeError):
            f.fit(_descr(x))
        q = rng.rand(48)
        zq = (q - x.mean(0)).dot(rps)
        assert np.array_equal(f.get_hash(q)[np.abs(zq) > 1e-4], (zq >= 0)[np.abs(zq) > 1e-4])
    # drives the LSH index like ItqFunctor does
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), 'euclidean')
    index.build_index(_descr(x))
    r, d = index.nn(DescriptorMemoryElement(9999).set_vector(x[10]), 3)
    assert r[0].uuid() == 10 and d[0] == 0.0


def _matrix_index(f):
    return LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), 'euclidean')


def _same_neighbours(a, ua, b, ub, q, n):
    """nn_batch of two indexes agrees once rows are translated to uuids (exact distances, same order)."""
    ra, da = a.nn_batch(q, n)
    rb, db = b.nn_batch(q, n)
    np.testing.assert_array_equal(da, db)
    ta = np.where(ra >= 0, np.asarray(ua, dtype=object)[np.clip(ra, 0, None)], None)
    tb = np.where(rb >= 0, np.asarray(ub, dtype=object)[np.clip(rb, 0, None)], None)
    assert (ta == tb).all()


def test_matrix_ingest_matches_reference_state_and_oracle(golden):
    """SURVEY 8f N3, pinned OUTSIDE the product: the same build / update / remove sequence the
    UNMODIFIED reference ran through LSHNearestNeighborIndex.update_index / remove_from_index
    (lsh.py:331-450; fixture tests/golden/lsh_ingest.npz, generator oracle/gen_golden.py::gen_lsh_ingest)
    goes through update_index_matrix / remove_from_index_matrix.  After EVERY step:
      * hash2uuids_view() == the reference's hash2uuids KVS (keys and uuid sets),
      * the LinearHashIndex content == the reference's hash_index.index, count == reference count(),
      * nn_batch == the reference's nn() answers (rtol 1e-5; order where the reference's tie order is defined),
      * nn_batch == the oracle's LSH query over the live rows (canonical ties: exact rows)."""
    g = golden("lsh_ingest")
    N, D, b, seed = gi.INGEST_SHAPE
    x, qs = gi.ingest_inputs()
    f = ItqFunctor(bit_length=b)
    f.mean_vec, f.rotation = g["mean"], g["rot"]
    index = _matrix_index(f)
    n = gi.INGEST_N
    x64 = x.astype(np.float64)
    mean, rot = np.real(g["mean"]), np.real(g["rot"])
    all_codes = O.pack_codes(O.itq_hash(x64, mean, rot), 1)
    q_codes = O.pack_codes(O.itq_hash(qs.astype(np.float64), mean, rot), 1)
    live = set()
    for si, step in enumerate(gi.INGEST_STEPS):
        op, rows = gi.ingest_step_rows(step)
        if op == "add":
            index.update_index_matrix(x[rows], uuids=rows)
            live |= set(rows)
        else:
            index.remove_from_index_matrix(rows)
            live -= set(rows)
        # --- state vs the reference's containers
        keys = [int(h, 16) for h in g["s%d_keys_hex" % si]]
        off, uu = g["s%d_off" % si], g["s%d_uuids" % si]
        want = {k: set(int(u) for u in uu[off[i]:off[i + 1]]) for i, k in enumerate(keys)}
        view = index.hash2uuids_view()
        assert {k: view.get(k) for k in view.keys()} == want, "step %d" % si
        assert sum(len(v) for v in view.values()) == int(g["s%d_count" % si]) == index.count_rows()
        assert index.hash_index.index == {int(h, 16) for h in g["s%d_hash_index_hex" % si]}
        # --- queries vs the reference's nn() and vs the oracle
        rows_b, dists_b = index.nn_batch(qs, n)
        uuids = index.mirror_uuids()
        live_rows = np.array(sorted(live))
        table = O.unique_code_table(all_codes[live_rows])[0]
        ordered = 0
        for qi in range(len(qs)):
            keep = rows_b[qi] >= 0
            got_u = [uuids[r] for r in rows_b[qi][keep]]
            orow, od = O.lsh_nn(x64[live_rows], all_codes[live_rows], qs[qi].astype(np.float64), q_codes[qi:qi + 1], n, "euclidean")
            np.testing.assert_allclose(dists_b[qi][keep], od, rtol=1e-5, atol=1e-9)
            if len(od) < 2 or np.diff(od).min() > 1e-6 * od.max():
                assert got_u == [int(live_rows[r]) for r in orow], "step %d query %d" % (si, qi)
            srt = np.sort(O.hamming_distances(table, q_codes[qi]))
            if n < len(table) and srt[n - 1] == srt[n]:
                continue            # the reference's candidate pool depends on its unspecified tie order
            ref_u, ref_d = g["s%d_q%d_uuids" % (si, qi)], g["s%d_q%d_dists" % (si, qi)]
            np.testing.assert_allclose(dists_b[qi][keep], ref_d, rtol=1e-5, atol=5e-8)
            if len(ref_d) < 2 or np.diff(ref_d).min() > 1e-4 * ref_d.max():
                assert got_u == list(ref_u), "step %d query %d" % (si, qi)
                ordered += 1
        assert ordered >= 3, "step %d: too few queries with a defined order" % si
    # KeyError before any mutation (lsh.py:407-416): uuid 3 is live again, 7 is not
    with pytest.raises(KeyError):
        index.remove_from_index_matrix([3, 7])
    assert index.count_rows() == int(g["s%d_count" % (len(gi.INGEST_STEPS) - 1)])


def test_index_over_prepopulated_collaborators_is_queryable_and_never_partial():
    """all reference state lives in the collaborators (lsh.py:160-232), so a second index
    object over the SAME stores (what from_config over persisted stores gives) must answer nn_batch at
    once -- and after one update_index it must see every row, not only the new one."""
    rng = np.random.RandomState(41)
    x = rng.rand(700, 24)
    f = ItqFunctor(bit_length=12, itq_iterations=5, random_seed=1)
    f.fit_matrix(x.astype(np.float32))
    ds, kvs, hi = MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex()
    a = LSHNearestNeighborIndex(f, ds, kvs, hi, 'euclidean')
    a.build_index(_descr(x[:600]))
    qs = rng.rand(15, 24).astype(np.float32)
    ra, da = a.nn_batch(qs, 5)
    b = LSHNearestNeighborIndex(f, ds, kvs, hi, 'euclidean')     # empty mirror, populated collaborators
    assert b.count() == 600 and b.count_rows() == 0
    rb, db = b.nn_batch(qs, 5)                                    # mirror rebuilt from the descriptor set
    np.testing.assert_array_equal(da, db)
    ua, ub = a.mirror_uuids(), b.mirror_uuids()
    assert [[ua[r] for r in row] for row in ra] == [[ub[r] for r in row] for row in rb]
    c = LSHNearestNeighborIndex(f, ds, kvs, hi, 'euclidean')
    c.update_index(_descr(x[600:], 600))                         # first call on a stale mirror is an update
    assert c.count_rows() == 700 == c.count()
    fresh = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), 'euclidean')
    fresh.build_index(_descr(x))
    rc, dc = c.nn_batch(qs, 5)
    rf, df = fresh.nn_batch(qs, 5)
    np.testing.assert_array_equal(dc, df)
    # ... and a now sees a descriptor set that changed behind its back: never a partial answer
    ra2, da2 = a.nn_batch(qs, 5)
    np.testing.assert_array_equal(da2, df)
    # element removal keeps row numbers (tombstones), KeyError leaves everything alone
    c.remove_from_index([0, 1, 2, 650])
    assert c.count_rows() == 696 == c.count()
    with pytest.raises(KeyError):
        c.remove_from_index([5, 0])
    assert c.count_rows() == 696
    keep = np.setdiff1d(np.arange(700), [0, 1, 2, 650])
    fresh.build_index([DescriptorMemoryElement(int(i)).set_vector(x[i]) for i in keep])
    rc, dc = c.nn_batch(qs, 5)
    rf, df = fresh.nn_batch(qs, 5)
    np.testing.assert_array_equal(dc, df)
    uc, uf = c.mirror_uuids(), fresh.mirror_uuids()
    assert [[uc[r] for r in row if r >= 0] for row in rc] == [[uf[r] for r in row if r >= 0] for row in rf]


def test_any_n_through_the_plugin_api_and_heavy_collisions():
    """n beyond the scan kernels' 2048-entry lists (LinearHashIndex.nn, LSH nn / nn_batch) and
    the candidate pool of short codes (default bit_length=8: a handful of codes, thousands of rows each)."""
    rng = np.random.RandomState(43)
    bits = rng.rand(6000, 40) > 0.5
    hi = LinearHashIndex()
    hi.build_index(bits)
    codes, dists = hi.nn(bits[0], 3000)
    table = O.unique_code_table(O.pack_codes(bits, 2))[0]
    od, oi = O.hamming_topk(table, O.pack_codes(bits[:1], 2), 3000)
    assert codes.shape == (3000, 40) and np.array_equal(O.pack_codes(codes, 2), table[oi[0]])
    np.testing.assert_allclose(dists, od[0] / 40.0)
    # heavy collisions: 8-bit codes over 20k rows, n = 2500 -> every candidate list holds thousands of rows
    x = rng.rand(20_000, 16).astype(np.float32)
    f = ItqFunctor(bit_length=8, itq_iterations=5, random_seed=2)
    f.fit_matrix(x)
    index = _matrix_index(f)
    index.build_index_matrix(x)
    qs = rng.rand(3, 16).astype(np.float32)
    n = 2500
    rows, d = index.nn_batch(qs, n)
    codes_dev = index._mirror.codes.cpu().numpy().view(np.uint32)
    qc = f.get_hash_packed(qs).cpu().numpy().view(np.uint32)
    x64 = x.astype(np.float64)
    for qi in range(3):
        orow, od = O.lsh_nn(x64, codes_dev, qs[qi].astype(np.float64), qc[qi:qi + 1], n, "euclidean")
        np.testing.assert_allclose(d[qi][:len(od)], od, rtol=1e-5)
        gaps = np.diff(od) > 1e-6 * od.max()
        same = rows[qi][:len(od)] == orow
        assert same[np.concatenate([[True], gaps]) & np.concatenate([gaps, [True]])].all()


def test_snapshot_validation_and_safe_loading(tmp_path):
    """load_snapshot reads with weights_only=True and refuses a snapshot that does not fit
    the index (code width / dimension / distance method) without modifying anything."""
    rng = np.random.RandomState(47)
    x = rng.rand(500, 32).astype(np.float32)
    f16 = ItqFunctor(bit_length=16, itq_iterations=3, random_seed=0)
    f16.fit_matrix(x)
    a = _matrix_index(f16)
    a.build_index_matrix(x, uuids=["u%d" % i for i in range(500)])
    p = str(tmp_path / "snap.pt")
    a.save_snapshot(p)
    ok = _matrix_index(f16)
    ok.load_snapshot(p)
    assert ok.count_rows() == 500 and ok.mirror_uuids()[7] == "u7"
    f40 = ItqFunctor(bit_length=40, itq_iterations=3, random_seed=0)   # 2 words: does not fit 1-word codes
    f40.fit_matrix(rng.rand(500, 64).astype(np.float32))
    bad = _matrix_index(f40)
    with pytest.raises(ValueError):
        bad.load_snapshot(p)
    assert bad.count_rows() == 0
    other = LSHNearestNeighborIndex(f16, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), 'hik')
    with pytest.raises(ValueError):
        other.load_snapshot(p)
    # uuids that need unpickling are refused unless the caller vouches for the file
    b = _matrix_index(f16)
    b.build_index_matrix(x[:10], uuids=[("t", i) if i else frozenset([1]) for i in range(10)])
    p2 = str(tmp_path / "snap2.pt")
    b.save_snapshot(p2)
    c = _matrix_index(f16)
    with pytest.raises(ValueError):
        c.load_snapshot(p2)
    c.load_snapshot(p2, trust_pickle=True)
    assert c.count_rows() == 10


def test_matrix_update_remove_equal_fresh_build():
    """update_index_matrix / remove_from_index_matrix (device-speed ingest: appended
    buffers, tombstones, re-index over the live rows) leave the index in the state a fresh
    build_index_matrix over the same rows gives -- reference semantics of lsh.py:331-450."""
    import torch
    rng = np.random.RandomState(31)
    centres = rng.rand(60, 32)
    x = (centres[rng.randint(0, 60, 6000)] + 0.05 * rng.randn(6000, 32)).astype(np.float32)
    q = (centres[rng.randint(0, 60, 150)] + 0.05 * rng.randn(150, 32)).astype(np.float32)
    f = ItqFunctor(bit_length=16, itq_iterations=10, random_seed=0)
    f.fit_matrix(x)
    x0, x1, x2 = x[:3000], x[3000:4500], x[4500:]

    # pure appends with default uuids (= row numbers), two batches
    a = _matrix_index(f)
    a.update_index_matrix(x0)                       # empty index: same as build
    a.update_index_matrix(x1)
    a.update_index_matrix(x2)
    fresh = _matrix_index(f)
    fresh.build_index_matrix(x)
    assert a.count_rows() == 6000 and a._mirror.num_codes == fresh._mirror.num_codes
    _same_neighbours(a, range(6000), fresh, range(6000), q, 7)
    assert torch.equal(a.hash_index.code_table, fresh.hash_index.code_table)

    # removal: tombstones, rows keep their numbers; unknown / already removed uuid -> KeyError, nothing changes
    gone = rng.choice(6000, 1700, replace=False)
    a.remove_from_index_matrix([int(g) for g in gone])
    keep = np.setdiff1d(np.arange(6000), gone)
    fresh = _matrix_index(f)
    fresh.build_index_matrix(x[keep], uuids=[int(k) for k in keep])
    assert a.count_rows() == len(keep) and a._mirror.num_rows == 6000
    _same_neighbours(a, range(6000), fresh, fresh.mirror_uuids(), q, 7)
    for bad in ([int(gone[0])], [6000], [int(keep[0]), -5]):
        with pytest.raises(KeyError):
            a.remove_from_index_matrix(bad)
    assert a.count_rows() == len(keep)
    _same_neighbours(a, range(6000), fresh, fresh.mirror_uuids(), q, 7)
    with pytest.raises(ValueError):
        a.update_index_matrix(np.zeros((0, 32), np.float32))

    # named uuids: overwrite in place, append the new ones, compaction once half the rows are dead
    names = ["u%d" % i for i in range(6000)]
    b = _matrix_index(f)
    b.build_index_matrix(x0, uuids=names[:3000])
    x_over = x[5000:5010]
    b.update_index_matrix(np.concatenate([x_over, x1]), uuids=names[5:15] + names[3000:4500])
    want = np.concatenate([x0, x1])
    want[5:15] = x_over
    fresh = _matrix_index(f)
    fresh.build_index_matrix(want, uuids=names[:4500])
    assert b.count_rows() == 4500
    _same_neighbours(b, b.mirror_uuids(), fresh, names[:4500], q, 5)
    b.remove_from_index_matrix(names[:2600])                   # > half dead: compacted
    assert b._mirror.num_rows == b.count_rows() == 1900
    fresh = _matrix_index(f)
    fresh.build_index_matrix(want[2600:], uuids=names[2600:4500])
    _same_neighbours(b, b.mirror_uuids(), fresh, names[2600:4500], q, 5)
    b.update_index_matrix(x[:3], uuids=names[:3])              # removed uuids may come back
    assert b.count_rows() == 1903

    # snapshot: a restarted process reloads rows, codes, tombstones and uuids, re-derives table + CSR
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        b.save_snapshot(os.path.join(td, "lsh.pt"))
        c = _matrix_index(f)
        c.load_snapshot(os.path.join(td, "lsh.pt"))
    assert c.count_rows() == b.count_rows() and c.mirror_uuids() == b.mirror_uuids()
    _same_neighbours(b, b.mirror_uuids(), c, c.mirror_uuids(), q, 5)
    assert torch.equal(c.hash_index.code_table, b.hash_index.code_table)

    # the KeyValueStore view is what build_index would have put into hash2uuids_kvstore
    view = b.hash2uuids_view()
    ref = _matrix_index(f)
    uu = b.mirror_uuids()
    live = [r for r in range(b._mirror.num_rows)]
    codes = f.get_hash_packed(b._mirror.x).cpu().numpy().view(np.uint32)
    want_map = {}
    for r in live:
        want_map.setdefault(int(bitutil.words_to_ints(codes[r:r + 1])[0]), set()).add(uu[r])
    assert view.count() == len(want_map) and set(view.keys()) == set(want_map)
    assert {k: view.get(k) for k in view.keys()} == want_map
    some = next(iter(want_map))
    assert view.has(some) and not view.has(-1) and view.get(-1, None) is None
    assert sum(len(v) for v in view.values()) == b.count_rows()
    with pytest.raises(KeyError):
        view.get(-1)
    with pytest.raises(ReadOnlyError):
        view.add(1, {2})
    assert view.is_read_only() and ref.count_rows() == 0


def test_itq_fit_streaming_equals_materialised():
    """The V-free ITQ iteration (fit.itq_rotation_streaming) is the same algorithm up to FP64
    re-association: same rotation to ~1e-9, same training codes except rounding-level projections."""
    from smqtk_indexing_b200 import fit as fitops
    rng = np.random.RandomState(8)
    x = rng.rand(4000, 48).astype(np.float32)
    for normalize in (None, 2):
        c0, m0, r0 = fitops.itq_fit(x, 16, itq_iterations=8, normalize=normalize, random_seed=2, streaming=False)
        c1, m1, r1 = fitops.itq_fit(x, 16, itq_iterations=8, normalize=normalize, random_seed=2, streaming=True)
        np.testing.assert_array_equal(m0, m1)
        np.testing.assert_allclose(r1, r0, rtol=0, atol=1e-8)
        assert (c0 != c1).mean() < 1e-3


def test_itq_fit_streaming_on_tensor_core_hash():
    """float32 data of an aligned shape: the streaming fit takes its sign bits from the production
    tensor-core hash kernel; the model agrees with the FP64 fit up to rounding-level bit flips."""
    from smqtk_indexing_b200 import fit as fitops
    rng = np.random.RandomState(18)
    x = rng.rand(6000, 64).astype(np.float32)
    c0, m0, r0 = fitops.itq_fit(x, 32, itq_iterations=6, random_seed=4, streaming=False)
    c1, m1, r1 = fitops.itq_fit(x, 32, itq_iterations=6, random_seed=4, streaming=True)
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_allclose(r1, r0, rtol=0, atol=1e-4)
    np.testing.assert_allclose(r1.T @ r1, np.eye(32), atol=1e-9)
    assert (c0 != c1).mean() < 1e-3


def test_itq_fit_streaming_on_tensor_core_gram():
    """>= 65536 aligned float32 rows: the streaming fit also takes UX^T.X from the tensor cores."""
    from smqtk_indexing_b200 import fit as fitops
    rng = np.random.RandomState(19)
    n = fitops.TC_GRAM_MIN_ROWS + 4321
    # clustered descriptors with a decaying spectrum (on i.i.d. data every rotation is as good as any
    # other and the ITQ iteration amplifies rounding-level differences into a different model)
    centres = rng.randn(24, 64) * np.linspace(2.0, 0.2, 64)
    x = (centres[rng.randint(0, 24, n)] + 0.15 * rng.randn(n, 64)).astype(np.float32)
    c0, m0, r0 = fitops.itq_fit(x, 32, itq_iterations=5, random_seed=4, streaming=False)
    c1, m1, r1 = fitops.itq_fit(x, 32, itq_iterations=5, random_seed=4, streaming=True)
    c2, m2, r2 = fitops.itq_fit(x, 32, itq_iterations=5, random_seed=4, streaming=True, tensor_cores=False)
    np.testing.assert_array_equal(m0, m1)
    # FP64 streaming == FP64 with v materialised
    np.testing.assert_allclose(r2, r0, rtol=0, atol=1e-8)
    # tensor-core Gram: the Gram itself is good to 1e-8 (next test), but sign() makes the ITQ iteration
    # discontinuous -- a handful of flipped rows per iteration moves the rotation at the 1e-3 level.
    # Same model up to that: orthonormal, same quantisation loss, 99 % of the bits.
    np.testing.assert_allclose(r1.T @ r1, np.eye(32), atol=1e-9)
    np.testing.assert_allclose(r1, r0, rtol=0, atol=5e-2)
    assert (c0 != c1).mean() < 1e-2
    xc = x[:20000].astype(np.float64) - m0.astype(np.float64)
    loss = [np.square(np.sign(xc @ r) - xc @ r).sum() for r in (r0, r1)]
    assert abs(loss[1] - loss[0]) < 1e-3 * loss[0], loss


@pytest.mark.parametrize("n,D,b,normalize", [(5000, 64, 32, None), (70000, 256, 256, None), (9000, 512, 128, 2),
                                             (4100, 96, 64, None), (20000, 128, 100, None)])
def test_gram_bits_tensor_core_matches_fp64(n, D, b, normalize):
    """sb_fit_gram_bits_tc (MN-major TF32 operands on fixed-point grids, exact FP32 accumulation,
    FP64 across 8192-row windows) vs the FP64 DFMA Gram on the same packed sign bits: the only
    error is the rounding of each centred value to 2^-22 of the bound -- a random walk of
    sqrt(n) * 2^-22 * bound, asserted here as 1e-7 of sum |x - mean|."""
    import torch
    from smqtk_indexing_b200 import device as dev, fit as fitops
    from smqtk_indexing_b200.utils.bits import words_for_bits
    rng = np.random.RandomState(n + b)
    x = torch.from_numpy(rng.rand(n, D).astype(np.float32)).cuda()
    W = words_for_bits(b)
    bits = rng.rand(n, b) > 0.5
    codes = dev.codes_to_device(O.pack_codes(bits, W))
    div = fitops.row_div(x, normalize)
    mean = fitops.col_mean(x, div)
    ref = fitops.gram(codes, x, a_bits=b, b_div=div, b_mean=mean)
    got = fitops.gram_bits_tc(codes, b, x, mean.to(torch.float32), None if div is None else div.to(torch.float32))
    torch.cuda.synchronize()
    xc = x.double() / (div[:, None] if div is not None else 1.0) - mean[None, :]
    scale = xc.abs().sum(0)[None, :]                        # worst-case magnitude of each column sum
    err = ((got - ref).abs() / scale).max().item()
    assert err < 1e-7, err
    # and against numpy on a small slice of the output
    ux = np.where(bits[:, :4], 1.0, -1.0)
    np.testing.assert_allclose(ref[:4].cpu().numpy(), ux.T @ xc.cpu().numpy(), rtol=1e-9, atol=1e-7)
