"""
GPU parity at BASELINE.json's sizes and shapes.

* C2 (10M x 256-bit codes, 4096 queries, k=10) is far beyond what the oracle can
  brute-force per test run, so the full-size checks are size-independent
  properties of the result (every returned distance re-derived from the table,
  canonical ordering, exact agreement with an independent torch brute force on a
  sample of queries, shard-and-merge == single scan, few-queries kernel ==
  batched kernel) -- plus the oracle on the rows it can afford.
* C1 (100k x 128-d, ITQ-64, euclidean, k=10) and C4 (hik, 1M x 4096-d, ITQ-256, k=50; plus a
  reduced 20k-row variant) run the whole plugin pipeline against the oracle's array-form LSH query.
* C5 (fit + build) runs at 10M x 256-d, where the streaming tensor-core fit engages.
* The C2 scan tests FAIL unless the tcgen05 kernel produced the result (no silent fallback).
"""
import numpy as np
import pytest
import torch

import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from smqtk_indexing_b200 import device
    device.require_cuda()
    return device


_LUT = None


def _popcount_rows(x_i32: torch.Tensor) -> torch.Tensor:
    """popcount per row of an int32[n, W] tensor with a byte LUT (independent of the kernels)."""
    global _LUT
    if _LUT is None or _LUT.device != x_i32.device:
        _LUT = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=x_i32.device)
    b = x_i32.contiguous().view(torch.uint8)
    return _LUT[b.long()].sum(dim=1, dtype=torch.int32)


def _topk_on_tensor_cores(dev, db, q, k):
    """hamming_topk that FAILS unless the tcgen05 scan produced the result: an overflowed batch is
    redone by the predicated XOR/POPC scan (device.hamming_scan_keys), which would let a
    full-size test pass without the headline kernel."""
    from smqtk_indexing_b200 import _lib
    overflows0 = dev.tc_scan_overflows()
    _lib.profile_fetch()
    _lib.profile_enable(True)
    try:
        dist, idx = dev.hamming_topk(db, q, k)
        torch.cuda.synchronize()
    finally:
        _lib.profile_enable(False)
    kernels = {name for name, _ in _lib.profile_fetch()}
    assert "ham_filter_tc_kernel" in kernels, "the tensor-core scan did not run: %s" % sorted(kernels)
    # (the predicated fallback kernels are always launched and return at once: their names prove nothing)
    assert dev.tc_scan_overflows() == overflows0, "the tensor-core scan overflowed a candidate buffer: fallback result"
    return dist, idx


def _check_topk_properties(db, q, k, dist, idx, sample):
    """Size-independent checks of a top-k result + exact agreement with a torch brute force on `sample`."""
    Q, U = q.shape[0], db.shape[0]
    assert int(idx.min()) >= 0 and int(idx.max()) < U
    rows = db[idx.reshape(-1)].reshape(Q, k, -1)
    x = (rows ^ q[:, None, :]).reshape(Q * k, -1)
    assert torch.equal(_popcount_rows(x).reshape(Q, k), dist)
    key = dist.to(torch.int64) * (1 << 40) + idx
    assert bool((key[:, 1:] > key[:, :-1]).all())
    for qi in sample:
        d_all = torch.zeros(U, dtype=torch.int32, device=db.device)
        for s0 in range(0, U, 2_000_000):                      # bounded temporaries
            d_all[s0:s0 + 2_000_000] = _popcount_rows(db[s0:s0 + 2_000_000] ^ q[qi][None, :])
        kk = d_all.to(torch.int64) * (1 << 24) + torch.arange(U, device=db.device)
        best = torch.topk(kk, k, largest=False, sorted=True).values
        assert torch.equal(best >> 24, dist[qi].to(torch.int64))
        assert torch.equal(best & ((1 << 24) - 1), idx[qi])


@pytest.fixture(scope="module")
def c2_scan(dev):
    U, W, Q, k = 10_000_000, 8, 4096, 10
    g = torch.Generator(device="cuda").manual_seed(42)
    db = torch.randint(-2 ** 31, 2 ** 31 - 1, (U, W), dtype=torch.int32, device="cuda", generator=g)
    q = torch.randint(-2 ** 31, 2 ** 31 - 1, (Q, W), dtype=torch.int32, device="cuda", generator=g)
    # plant duplicates and near-duplicates of some queries so that small distances and exact ties occur
    db[123456] = q[0]
    db[9_999_999] = q[0]
    near = q[1].clone()
    near[7] ^= 1                                                # one bit flipped
    db[5_000_000] = near
    near = q[1].clone()
    near[0] ^= 2
    db[5_000_001] = near
    dist, idx = _topk_on_tensor_cores(dev, db, q, k)
    return db, q, k, dist, idx


def test_c2_scan_result_is_consistent_with_the_table(c2_scan):
    db, q, k, dist, idx = c2_scan
    Q = q.shape[0]
    assert int(idx.min()) >= 0 and int(idx.max()) < db.shape[0]
    # every reported distance is the popcount of (row xor query)
    rows = db[idx.reshape(-1)].reshape(Q, k, -1)
    x = (rows ^ q[:, None, :]).reshape(Q * k, -1)
    assert torch.equal(_popcount_rows(x).reshape(Q, k), dist)
    # canonical order: ascending (distance, row), rows distinct
    key = dist.to(torch.int64) * (1 << 40) + idx
    assert bool((key[:, 1:] > key[:, :-1]).all())
    # planted rows are found, ties broken by row
    assert idx[0, 0].item() == 123456 and idx[0, 1].item() == 9_999_999 and dist[0, :2].tolist() == [0, 0]
    assert idx[1, :2].tolist() == [5_000_000, 5_000_001] and dist[1, :2].tolist() == [1, 1]


def test_c2_scan_equals_torch_brute_force_on_sampled_queries(c2_scan):
    db, q, k, dist, idx = c2_scan
    U = db.shape[0]
    for qi in (0, 1, 2, 777, 4095):
        d_all = torch.zeros(U, dtype=torch.int32, device=db.device)
        for s in range(0, U, 2_000_000):                      # bounded temporaries
            d_all[s:s + 2_000_000] = _popcount_rows(db[s:s + 2_000_000] ^ q[qi][None, :])
        key = d_all.to(torch.int64) * (1 << 24) + torch.arange(U, device=db.device)
        best = torch.topk(key, k, largest=False, sorted=True).values
        assert torch.equal(best >> 24, dist[qi].to(torch.int64))
        assert torch.equal(best & ((1 << 24) - 1), idx[qi])


def test_c2_scan_matches_oracle_on_a_table_prefix(dev, c2_scan):
    """The oracle (numpy restatement of linear.py:232-244) on the first 200k rows."""
    db, q, k, _, _ = c2_scan
    n = 200_000
    d, i = dev.hamming_topk(db[:n].contiguous(), q[:8].contiguous(), k)
    od, oi = O.hamming_topk(db[:n].cpu().numpy().view(np.uint32), q[:8].cpu().numpy().view(np.uint32), k)
    assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(i.cpu().numpy(), oi)


def test_c2_sharded_scan_and_merge_equals_single_scan(dev, c2_scan):
    db, q, k, dist, idx = c2_scan
    U = db.shape[0]
    cuts = [U * r // 8 // 4 * 4 for r in range(8)] + [U]
    parts = [dev.hamming_scan_keys(db[cuts[r]:cuts[r + 1]], q, k, idx_base=cuts[r]) for r in range(8)]
    md, mi = dev.topk_merge(torch.stack(parts).contiguous())
    assert torch.equal(md, dist) and torch.equal(mi, idx)


def test_c2_few_queries_kernel_equals_batched_kernel(dev, c2_scan):
    db, q, k, dist, idx = c2_scan
    for nq in (1, 3, 4):
        d, i = dev.hamming_topk(db, q[:nq].contiguous(), k)
        assert torch.equal(d, dist[:nq]) and torch.equal(i, idx[:nq])


def test_c2_hash_is_row_independent_and_matches_ffma(dev):
    """ITQ-256 over 512-d rows at scale: the tensor-core kernel's code of a row does not
    depend on where the row sits in the matrix, and agrees with the FFMA kernel except
    for projections at rounding level."""
    n, D, b = 1_000_000, 512, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.rand((n, D), generator=g, device="cuda")
    mean = torch.full((D,), 0.5, device="cuda")
    R = torch.from_numpy(np.linalg.qr(np.random.RandomState(7).randn(D, D))[0][:, :b].astype(np.float32)).cuda()
    img = dev.itq_rotation_image(R)
    codes = dev.itq_hash(X, mean, R, variant=2, r_image=img)
    pick = torch.randperm(n, generator=g, device="cuda")[:70_001]
    again = dev.itq_hash(X[pick].contiguous(), mean, R, variant=2, r_image=img)
    assert torch.equal(codes[pick], again)
    ffma, z = dev.itq_hash(X[:100_000], mean, R, variant=1, want_z=True)
    diff = _popcount_rows(codes[:100_000] ^ ffma)
    assert int(diff.sum()) <= 20                                # of 25.6M bits
    from smqtk_indexing_b200.utils.bits import unpack_bits
    bad = unpack_bits(dev.codes_to_host(codes[:100_000]), b) != unpack_bits(dev.codes_to_host(ffma), b)
    assert (np.abs(z.cpu().numpy()[bad]) < 2e-5).all()


def test_c2_scan_over_a_sorted_clustered_itq_table_10m(dev):
    """C2 on the table REAL ITQ codes produce: 10M x 256-bit codes of clustered descriptors, hashed by a
    fitted ITQ-256 model, sort-uniqued (near-duplicate codes are contiguous -- the case the golden-ratio
    visiting order exists for).  Queries come from the same clusters, so every query has thousands of
    codes at small distances.  The tcgen05 kernel must produce the result (no overflow fallback)."""
    from smqtk_indexing_b200 import codes as codeops
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    N, D, b, Q, k, C = 10_000_000, 256, 256, 4096, 10, 3000
    g = torch.Generator(device="cuda").manual_seed(77)
    centres = torch.rand((C, D), generator=g, device="cuda")

    def draw(n):
        c = torch.randint(0, C, (n,), generator=g, device="cuda")
        return centres[c] + 0.08 * torch.randn((n, D), generator=g, device="cuda")

    f = ItqFunctor(bit_length=b, itq_iterations=5, random_seed=0)
    f.fit_matrix(draw(200_000))
    parts = [f.get_hash_packed(draw(1_000_000)) for _ in range(N // 1_000_000)]   # descriptors are not kept
    table = codeops.sort_unique(torch.cat(parts))
    del parts
    U = table.shape[0]
    assert U > 0.9 * N                                          # noise keeps most codes distinct
    q = f.get_hash_packed(draw(Q))
    dist, idx = _topk_on_tensor_cores(dev, table, q, k)
    # clustered: the k-th neighbour is far below the 128 bits of random codes
    assert float(dist[:, -1].float().mean()) < 90
    _check_topk_properties(table, q, k, dist, idx, sample=(0, 1, 1234, 4095))
    # the sorted table really is clustered: adjacent rows are much closer than random pairs
    adj = _popcount_rows(table[1:200_001] ^ table[:200_000]).float().mean().item()
    assert adj < 100


# ------------------------------------------------------------------ config-shaped pipelines
def _oracle_lsh(x64, table, off, rows, q_vec, q_words, n, method):
    """lsh.py:470-519 restated with a prebuilt unique-code table (np_oracle.lsh_nn rebuilds it per call)."""
    _, near = O.hamming_topk(table, q_words, n)
    cand = np.concatenate([rows[off[c]:off[c + 1]] for c in near[0]])
    d = np.atleast_1d(O.DISTANCE_FUNCTIONS[method](q_vec, x64[cand]))
    o = np.argsort(d, kind="stable")[:n]
    return cand[o], d[o]


def _pipeline_case(N, D, b, k, method, n_queries, n_check, seed, normalise_rows=False, fit_rows=20_000):
    from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
    from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
    from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex
    rng = np.random.RandomState(seed)
    x = rng.rand(N, D).astype(np.float32)
    qs = rng.rand(n_queries, D).astype(np.float32)
    qs[0] = x[17]                                               # an indexed vector
    qs[1] = x[29] * (1 + 1e-3 * rng.rand(D).astype(np.float32))  # a near duplicate
    if normalise_rows:                                          # L1-normalised histograms
        x /= x.sum(axis=1, keepdims=True)
        qs /= qs.sum(axis=1, keepdims=True)
    f = ItqFunctor(bit_length=b, itq_iterations=10, random_seed=0)
    f.fit_matrix(x[:fit_rows])
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), method)
    index.build_index_matrix(x)
    rows, dists = index.nn_batch(qs, k)
    # oracle on the DEVICE codes (hash parity has its own tests) and float64 arithmetic
    W = index._mirror.codes.shape[1]
    codes = index._mirror.codes.cpu().numpy().view(np.uint32)
    table, _, off, crow = O.unique_code_table(codes)
    assert len(table) == index._mirror.num_codes
    x64 = x.astype(np.float64)
    qcodes = f.get_hash_packed(qs).cpu().numpy().view(np.uint32)
    checked_order = 0
    for qi in range(n_check):
        orow, od = _oracle_lsh(x64, table, off, crow, qs[qi].astype(np.float64), qcodes[qi:qi + 1], k, method)
        m = len(od)
        np.testing.assert_allclose(dists[qi][:m], od, rtol=1e-5, atol=1e-9)
        assert (rows[qi][m:] == -1).all()
        if m > 1 and np.diff(od).min() > 1e-6 * max(od.max(), 1e-30):
            assert list(rows[qi][:m]) == list(orow)
            checked_order += 1
    assert checked_order >= n_check // 2
    return rows, dists


def test_c1_pipeline_100k_x_128_itq64_euclidean_k10():
    rows, dists = _pipeline_case(100_000, 128, 64, 10, "euclidean", n_queries=1000, n_check=40, seed=0)
    assert rows[0][0] == 17 and dists[0][0] == 0.0
    assert rows[1][0] == 29


def test_c4_pipeline_hik_4096d_itq256_k50():
    rows, dists = _pipeline_case(20_000, 4096, 256, 50, "hik", n_queries=64, n_check=12, seed=4,
                                 normalise_rows=True, fit_rows=5_000)
    assert rows[0][0] == 17 and abs(dists[0][0]) < 1e-6      # 1 - sum(fp32 row): not exactly 0


def test_c4_full_size_1m_x_4096_hik_itq256_k50(dev):
    """BASELINE configs[3] at its stated size: 1M x 4096-d L1-normalised histograms (16.4 GB fp32 on the
    device), ITQ-256, histogram-intersection re-rank, k=50.  Descriptors are generated on the device;
    the oracle sees the device codes (host copy, 32 MB) and float64 copies of only the candidate rows."""
    from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
    from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
    from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex
    N, D, b, k, Q, n_check = 1_000_000, 4096, 256, 50, 64, 10
    g = torch.Generator(device="cuda").manual_seed(44)
    x = torch.empty((N, D), dtype=torch.float32, device="cuda")
    for s0 in range(0, N, 100_000):
        blk = torch.rand((100_000, D), generator=g, device="cuda")
        x[s0:s0 + 100_000] = blk / blk.sum(dim=1, keepdim=True)
    qs = torch.rand((Q, D), generator=g, device="cuda")
    qs[0] = x[17]
    qs[1] = x[29] * (1 + 1e-3 * torch.rand(D, generator=g, device="cuda"))
    qs = qs / qs.sum(dim=1, keepdim=True)
    f = ItqFunctor(bit_length=b, itq_iterations=5, random_seed=0)
    f.fit_matrix(x[:20_000])
    index = LSHNearestNeighborIndex(f, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(), "hik")
    index.build_index_matrix(x)
    rows, dists = index.nn_batch(qs, k)
    assert rows[0][0] == 17 and abs(dists[0][0]) < 1e-6 and rows[1][0] == 29
    codes = index._mirror.codes.cpu().numpy().view(np.uint32)
    table, _, off, crow = O.unique_code_table(codes)
    assert len(table) == index._mirror.num_codes
    qcodes = f.get_hash_packed(qs).cpu().numpy().view(np.uint32)
    qs64 = qs.cpu().numpy().astype(np.float64)
    ordered = 0
    for qi in range(n_check):
        _, near = O.hamming_topk(table, qcodes[qi:qi + 1], k)
        cand = np.concatenate([crow[off[c]:off[c + 1]] for c in near[0]])
        xc = x[torch.from_numpy(cand).cuda()].cpu().numpy().astype(np.float64)
        d = np.atleast_1d(O.DISTANCE_FUNCTIONS["hik"](qs64[qi], xc))
        o = np.argsort(d, kind="stable")[:k]
        np.testing.assert_allclose(dists[qi][:len(o)], d[o], rtol=1e-5, atol=1e-9)
        if np.diff(d[o]).min() > 1e-6 * d[o].max():
            assert list(rows[qi][:len(o)]) == list(cand[o])
            ordered += 1
    assert ordered >= n_check // 2


def test_c5_fit_and_build_10m_x_256(dev):
    """BASELINE configs[4] shape at 10M rows (the streaming tensor-core fit engages above 4 GiB of V):
    ItqFunctor.fit_matrix over 10M x 256-d -> 64 bits, then hash + unique table + CSR over all rows.
    The model must be orthonormal and as good as the FP64 fit of the same data (quantisation loss), the
    build must cover every row exactly once and agree with re-hashing."""
    from smqtk_indexing_b200 import engine, fit as fitops
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    n, D, b = 10_000_000, 256, 64
    g = torch.Generator(device="cuda").manual_seed(55)
    centres = torch.randn((64, D), generator=g, device="cuda") * torch.linspace(2.0, 0.2, D, device="cuda")
    X = torch.empty((n, D), dtype=torch.float32, device="cuda")
    for s0 in range(0, n, 1_000_000):
        c = torch.randint(0, 64, (1_000_000,), generator=g, device="cuda")
        X[s0:s0 + 1_000_000] = centres[c] + 0.15 * torch.randn((1_000_000, D), generator=g, device="cuda")
    assert n * b * 8 > fitops.STREAMING_V_BYTES                 # the streaming path is the one under test
    f = ItqFunctor(bit_length=b, itq_iterations=5, random_seed=0)
    f.fit_matrix(X, want_codes=False)
    r1 = np.asarray(f.rotation)
    np.testing.assert_allclose(r1.T @ r1, np.eye(b), atol=1e-9)
    # FP64 fit of the same rows (parity path), compared by quantisation loss on a sample
    _, m0, r0 = fitops.itq_fit(X, b, itq_iterations=5, random_seed=0, streaming=True, tensor_cores=False,
                               want_codes=False)
    np.testing.assert_allclose(np.asarray(f.mean_vec, dtype=np.float64), m0.astype(np.float64), rtol=0, atol=1e-6)
    xc = X[:50_000].double().cpu().numpy() - m0.astype(np.float64)
    loss = [np.square(np.sign(xc @ r) - xc @ r).sum() for r in (r0, r1)]
    assert abs(loss[1] - loss[0]) < 2e-3 * loss[0], loss
    m = engine.DeviceLshIndex()
    codes = f.get_hash_packed(X)
    m.set_rows(X, codes)
    assert m.num_rows == n and int((m.csr_off[1:] - m.csr_off[:-1]).sum()) == n
    assert torch.equal(torch.sort(m.csr_rows).values, torch.arange(n, device="cuda"))
    assert torch.equal(m.table[m.row_code], m.codes)
    assert bool((m.csr_off[1:] > m.csr_off[:-1]).all())
    t = m.table.to(torch.int64) & 0xFFFFFFFF
    key = (t[:, 0] << 32) | t[:, 1]                             # 64-bit codes: table strictly ascending as integers
    key = key ^ (-2 ** 63)
    assert bool((key[1:] > key[:-1]).all())


def test_c5_fit_and_build_throughput_shape():
    """fit + build on 256-d descriptors (reduced C5): the fitted rotation is orthonormal,
    the build's codes equal re-hashing, unique table + CSR cover every row once."""
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    from smqtk_indexing_b200 import engine
    n, D, b = 300_000, 256, 64
    X = torch.rand((n, D), generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
    f = ItqFunctor(bit_length=b, itq_iterations=5, random_seed=0)
    fit_codes = f.fit_matrix(X[:50_000])
    r = np.asarray(f.rotation)
    np.testing.assert_allclose(r.T @ r, np.eye(b), atol=1e-9)
    m = engine.DeviceLshIndex()
    m.set_rows(X, f.get_hash_packed(X))
    assert m.num_rows == n and int((m.csr_off[1:] - m.csr_off[:-1]).sum()) == n
    assert torch.equal(torch.sort(m.csr_rows).values, torch.arange(n, device="cuda"))
    assert torch.equal(m.table[m.row_code], m.codes)
    # training codes from fit (FP64 path) vs the hash kernel on the same rows: identical up to rounding-level flips
    from smqtk_indexing_b200.utils.bits import unpack_bits
    hk = unpack_bits(m.codes[:50_000].cpu().numpy().view(np.uint32), b)
    assert (hk != fit_codes).mean() < 1e-4
