"""
GPU parity tests at the C-ABI level: every kernel against the oracle on the same
seeded inputs.  Integer results must be bit exact; floating-point tolerances
are written next to each assertion.
"""
import numpy as np
import pytest

import golden_inputs as gi
import np_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from smqtk_indexing_b200 import device as D
    D.require_cuda()
    return D


def _rand_table(rng, U, b, W, dup=False):
    bits = rng.rand(U, b) > 0.5
    if dup:                       # heavy ties: few distinct bits set
        bits = rng.rand(U, b) > 0.9
    return O.pack_codes(bits, W)


def _run_topk(dev, table, q, k, idx_base=0):
    dt = dev.codes_to_device(table)
    dq = dev.codes_to_device(q)
    d, i = dev.hamming_topk(dt, dq, k, idx_base)
    torch.cuda.synchronize()
    return d.cpu().numpy(), i.cpu().numpy()


@pytest.mark.parametrize("b,W", [(1, 1), (5, 1), (32, 1), (33, 2), (64, 2), (100, 4), (128, 4),
                                 (256, 8), (500, 16), (1024, 32)])
def test_hamming_topk_bit_exact(dev, b, W):
    rng = np.random.RandomState(b)
    for U, Q, k in [(1, 1, 1), (7, 3, 10), (1000, 5, 10), (4097, 33, 10), (20000, 130, 7), (3000, 2, 100)]:
        table = _rand_table(rng, U, b, W)
        q = O.pack_codes(rng.rand(Q, b) > 0.5, W)
        q[0] = table[U // 2]
        od, oi = O.hamming_topk(table, q, k, idx_base=5)
        d, i = _run_topk(dev, table, q, k, idx_base=5)
        kk = od.shape[1]
        assert np.array_equal(d[:, :kk], od), (b, U, Q, k)
        assert np.array_equal(i[:, :kk], oi), (b, U, Q, k)
        assert (d[:, kk:] == -1).all() and (i[:, kk:] == -1).all()


def test_hamming_topk_heavy_ties(dev):
    rng = np.random.RandomState(7)
    for b, W in [(8, 1), (32, 1), (256, 8)]:
        table = _rand_table(rng, 50000, b, W, dup=True)
        q = O.pack_codes(rng.rand(16, b) > 0.9, W)
        od, oi = O.hamming_topk(table, q, 25)
        d, i = _run_topk(dev, table, q, 25)
        assert np.array_equal(d, od) and np.array_equal(i, oi)


def test_hamming_variants_agree(dev):
    rng = np.random.RandomState(11)
    table = _rand_table(rng, 100000, 256, 8)
    q = O.pack_codes(rng.rand(64, 256) > 0.5, 8)
    dt, dq = dev.codes_to_device(table), dev.codes_to_device(q)
    k0 = dev.hamming_scan_keys(dt, dq, 10, variant=0)
    k1 = dev.hamming_scan_keys(dt, dq, 10, variant=1)
    assert torch.equal(k0, k1)
    od, oi = O.hamming_topk(table, q, 10)
    d, i = dev.topk_merge(k0.unsqueeze(0).contiguous())
    assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(i.cpu().numpy(), oi)


def test_hamming_large_k(dev):
    rng = np.random.RandomState(13)
    table = _rand_table(rng, 5000, 64, 2)
    q = O.pack_codes(rng.rand(3, 64) > 0.5, 2)
    for k in (33, 500, 1000, 2048):
        od, oi = O.hamming_topk(table, q, k)
        d, i = _run_topk(dev, table, q, k)
        assert np.array_equal(d, od) and np.array_equal(i, oi), k


def test_hamming_any_k_beyond_the_scan_lists(dev):
    """the reference takes any n (heapq.nsmallest, linear.py:232-240); k beyond the scan
    kernels' 2048-entry lists goes through the exhaustive radix-sorted path, also in query chunks."""
    rng = np.random.RandomState(14)
    table = _rand_table(rng, 5000, 64, 2)
    table[100:140] = table[99]                                 # ties
    q = O.pack_codes(rng.rand(5, 64) > 0.5, 2)
    for k in (2049, 4999, 5000):
        od, oi = O.hamming_topk(table, q, k)
        d, i = _run_topk(dev, table, q, k)
        assert np.array_equal(d, od) and np.array_equal(i, oi), k
    d, i = _run_topk(dev, table, q, 6000)                      # more than the table holds: padded
    od, oi = O.hamming_topk(table, q, 5000)
    assert np.array_equal(d[:, :5000], od) and np.array_equal(i[:, :5000], oi)
    assert (d[:, 5000:] == -1).all() and (i[:, 5000:] == -1).all()
    old = dev.SORTED_TOPK_MAX_PAIRS
    dev.SORTED_TOPK_MAX_PAIRS = 2 * 5000                       # two queries per call
    try:
        d2, i2 = _run_topk(dev, table, q, 3000)
    finally:
        dev.SORTED_TOPK_MAX_PAIRS = old
    od, oi = O.hamming_topk(table, q, 3000)
    assert np.array_equal(d2, od) and np.array_equal(i2, oi)
    keys = dev.hamming_scan_keys(torch.from_numpy(table.view(np.int32)).cuda(), torch.from_numpy(q.view(np.int32)).cuda(),
                                 2500, idx_base=1 << 33)
    kd, ki = dev.decode_keys(keys)
    od, oi = O.hamming_topk(table, q, 2500)
    assert np.array_equal(kd.cpu().numpy(), od) and np.array_equal(ki.cpu().numpy(), oi + (1 << 33))


def test_sharded_scan_merge_equals_single(dev):
    """row-sharded table, per-shard top-k keys, sb_topk_merge == single scan."""
    rng = np.random.RandomState(17)
    table = _rand_table(rng, 30011, 256, 8)
    q = O.pack_codes(rng.rand(40, 256) > 0.5, 8)
    dq = dev.codes_to_device(q)
    cuts = np.linspace(0, len(table), 9).astype(int)
    keys = [dev.hamming_scan_keys(dev.codes_to_device(table[a:b]), dq, 10, idx_base=int(a))
            for a, b in zip(cuts[:-1], cuts[1:])]
    d, i = dev.topk_merge(torch.stack(keys).contiguous())
    od, oi = O.hamming_topk(table, q, 10)
    assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(i.cpu().numpy(), oi)


def _tc_keys(dev, table, q, k, idx_base=0):
    dt, dq = dev.codes_to_device(table), dev.codes_to_device(q)
    keys, flag = dev.hamming_scan_keys_tc(dt, dq, k, idx_base)
    d, i = dev.topk_merge(keys.unsqueeze(0).contiguous())
    torch.cuda.synchronize()
    return d.cpu().numpy(), i.cpu().numpy(), int(flag.item())


@pytest.fixture(params=["fp8", "fp4"])
def tc_fmt(request, dev):
    """Both operand formats of the tensor-core scan: packed FP4 (kind::mxf4, the default) and FP8 (kind::f8f6f4)."""
    old = dev.TC_SCAN_FORMAT
    dev.TC_SCAN_FORMAT = request.param
    yield request.param
    dev.TC_SCAN_FORMAT = old


@pytest.mark.parametrize("b,W", [(1, 1), (5, 1), (32, 1), (33, 2), (64, 2), (100, 4), (128, 4), (200, 8), (256, 8)])
def test_hamming_tensor_core_scan_bit_exact(dev, tc_fmt, b, W):
    """sb_hamming_scan_tc4 / sb_hamming_scan_tc (+-1 E2M1 / E4M3 dot products on tcgen05) return the oracle's
    keys exactly: ragged table sizes (partial granules / tiles / chunks), Q across query-block boundaries
    (240- and 256-column blocks), k up to 256, idx_base, a query equal to a table row."""
    rng = np.random.RandomState(100 + b)
    for U, Q, k in [(1, 1, 1), (7, 3, 10), (1000, 5, 10), (4097, 257, 10), (20000, 130, 7), (3000, 2, 100),
                    (70001, 300, 33), (9000, 20, 256), (5000, 600, 5), (3000, 1500, 3)]:
        table = _rand_table(rng, U, b, W)
        q = O.pack_codes(rng.rand(Q, b) > 0.5, W)
        q[0] = table[U // 2]
        od, oi = O.hamming_topk(table, q, k, idx_base=5)
        d, i, flag = _tc_keys(dev, table, q, k, idx_base=5)
        if flag:
            # only the tie-saturated toy widths may overflow (a 5-bit table is thousands of copies of 32
            # codes; real tables hold unique codes): the dispatcher then takes the XOR/POPC scan
            assert b <= 5
            dt, dq = dev.codes_to_device(table), dev.codes_to_device(q)
            keys = dev.hamming_scan_keys(dt, dq, k, idx_base=5, variant=dev.SCAN_VARIANT_TC)
            d, i = (t.cpu().numpy() for t in dev.topk_merge(keys.unsqueeze(0).contiguous()))
        kk = od.shape[1]
        assert np.array_equal(d[:, :kk], od), (b, U, Q, k)
        assert np.array_equal(i[:, :kk], oi), (b, U, Q, k)
        assert (d[:, kk:] == -1).all() and (i[:, kk:] == -1).all()


def test_hamming_tensor_core_scan_ties_and_clusters(dev, tc_fmt):
    """Tie-heavy tables, and a sorted table with large near-duplicate clusters (the order real ITQ
    codes arrive in): the golden-ratio visiting order keeps every chunk a uniform sample."""
    rng = np.random.RandomState(23)
    for b, W in [(8, 1), (32, 1), (256, 8)]:
        table = _rand_table(rng, 50000, b, W, dup=True)
        q = O.pack_codes(rng.rand(16, b) > 0.9, W)
        od, oi = O.hamming_topk(table, q, 25)
        # thousands of copies of a few codes: candidate buffers may overflow (-> XOR/POPC scan); exact either way
        keys = dev.hamming_scan_keys(dev.codes_to_device(table), dev.codes_to_device(q), 25, variant=dev.SCAN_VARIANT_TC)
        d, i = (t.cpu().numpy() for t in dev.topk_merge(keys.unsqueeze(0).contiguous()))
        assert np.array_equal(d, od) and np.array_equal(i, oi)
    centres = rng.rand(40, 256) > 0.5
    bits = centres[rng.randint(0, 40, 200000)] ^ (rng.rand(200000, 256) < 0.03)
    table = O.pack_codes(bits, 8)
    order = np.lexsort(tuple(table[:, w] for w in range(7, -1, -1)))
    table = np.ascontiguousarray(table[order])
    q = O.pack_codes(centres[rng.randint(0, 40, 300)] ^ (rng.rand(300, 256) < 0.03), 8)
    od, oi = O.hamming_topk(table, q, 10)
    d, i, flag = _tc_keys(dev, table, q, 10)
    assert flag == 0 and np.array_equal(d, od) and np.array_equal(i, oi)


def test_hamming_tensor_core_scan_far_codes_and_mixed_thresholds(dev, tc_fmt):
    """The FP4 scan packs three queries into 8-bit fields of one FP32 accumulator; a field leaves its window for
    pairs far above the threshold (d > 254 - (T - tq)).  Build exactly that: half of the queries have duplicates
    and neighbours at d <= 5 in the table (thresholds -> 0), the others only random codes (thresholds ~ 100),
    and the table holds complements / near-complements of the queries (d = 230 ... 256).  Borrows, carries and
    exponent drops may only ADD re-check work: the keys must equal the oracle's, without the fallback."""
    rng = np.random.RandomState(41)
    b, W, Q, k = 256, 8, 300, 10
    qbits = rng.rand(Q, b) > 0.5
    rows = [rng.rand(100000, b) > 0.5]
    for j in range(0, Q, 2):                                    # near neighbours of every second query
        near = np.repeat(qbits[j:j + 1], 14, axis=0)
        for r in range(4, 14):
            near[r, rng.choice(b, rng.randint(1, 6), replace=False)] ^= True
        rows.append(near)
    far = ~qbits[rng.randint(0, Q, 20000)]                      # complements, most with a few bits flipped back
    flips = rng.rand(20000, b) < (rng.rand(20000, 1) * 0.1)
    flips[:200] = False                                         # exact complements: d = 256
    rows.append(far ^ flips)
    bits = np.concatenate(rows)
    bits = bits[rng.permutation(len(bits))]
    table, q = O.pack_codes(bits, W), O.pack_codes(qbits, W)
    od, oi = O.hamming_topk(table, q, k)
    assert od[0, 3] == 0 and od[1, 0] > 60 and od[0, -1] <= 5   # the two kinds of query
    d, i, flag = _tc_keys(dev, table, q, k)
    assert flag == 0
    assert np.array_equal(d, od) and np.array_equal(i, oi)
    # every query close to one centre, 30000 rows close to its complement: millions of far pairs at once.  The
    # re-check list may overflow (-> flag, XOR/POPC scan through the dispatcher); the keys are exact either way
    centre = rng.rand(1, b) > 0.5
    qbits = centre ^ (rng.rand(Q, b) < 0.02)
    bits = np.concatenate([rng.rand(60000, b) > 0.5, (~centre) ^ (rng.rand(30000, b) < 0.03),
                           centre ^ (rng.rand(3000, b) < 0.05)])
    bits = bits[rng.permutation(len(bits))]
    table, q = O.pack_codes(bits, W), O.pack_codes(qbits, W)
    od, oi = O.hamming_topk(table, q, k)
    keys = dev.hamming_scan_keys(dev.codes_to_device(table), dev.codes_to_device(q), k, variant=dev.SCAN_VARIANT_TC)
    d, i = (t.cpu().numpy() for t in dev.topk_merge(keys.unsqueeze(0).contiguous()))
    assert np.array_equal(d, od) and np.array_equal(i, oi)
    # EVERY row far from every query (thresholds of 240 ... 256 bits: no room for the 8-bit windows above 252 --
    # the FP4 kernel must raise the flag rather than answer) and a batch whose thresholds straddle that limit
    for flip in (0.0, 0.03):
        bits = (~centre) ^ (rng.rand(70000, b) < flip)
        if flip:
            bits[:35000] = rng.rand(35000, b) > 0.5
        table = O.pack_codes(bits[rng.permutation(len(bits))], W)
        od, oi = O.hamming_topk(table, q, k)
        keys = dev.hamming_scan_keys(dev.codes_to_device(table), dev.codes_to_device(q), k, variant=dev.SCAN_VARIANT_TC)
        d, i = (t.cpu().numpy() for t in dev.topk_merge(keys.unsqueeze(0).contiguous()))
        assert np.array_equal(d, od) and np.array_equal(i, oi)


def test_hamming_scan_dispatch_and_overflow_fallback(dev):
    """hamming_scan_keys picks the tensor-core scan for large batches; identical keys either way.
    A table of identical rows overflows the candidate buffers (every row ties): the flag is raised
    and the XOR/POPC scan decides."""
    rng = np.random.RandomState(29)
    table = _rand_table(rng, 100000, 256, 8)
    q = O.pack_codes(rng.rand(200, 256) > 0.5, 8)
    dt, dq = dev.codes_to_device(table), dev.codes_to_device(q)
    assert not dev._tc_scan_pays(100000, 8, 200, 10) and dev._tc_scan_pays(10_000_000, 8, 64, 10)
    assert dev._tc_scan_pays(100000, 8, 4096, 10) and not dev._tc_scan_pays(10_000_000, 32, 4096, 10)
    k_auto = dev.hamming_scan_keys(dt, dq, 10)
    k_tc = dev.hamming_scan_keys(dt, dq, 10, variant=dev.SCAN_VARIANT_TC)
    k_popc = dev.hamming_scan_keys(dt, dq, 10, variant=1)
    assert torch.equal(k_auto, k_popc) and torch.equal(k_tc, k_popc)
    same = np.repeat(table[:1], 100000, axis=0)
    ds = dev.codes_to_device(same)
    before = dev.tc_scan_overflows()
    _lib = __import__("smqtk_indexing_b200._lib", fromlist=["x"])
    _lib.profile_fetch()
    _lib.profile_enable(True)
    keys = dev.hamming_scan_keys(ds, dq, 10, variant=dev.SCAN_VARIANT_TC)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    ran = [name for name, _ in _lib.profile_fetch()]
    # the overflow is handled ON THE DEVICE: the predicated XOR/POPC scan (sb_hamming_scan_if) ran after the
    # tensor-core scan and replaced the keys; the host never read the flag
    assert "ham_filter_tc_kernel" in ran and "hamming_scan_kernel" in ran
    assert dev.tc_scan_overflows() == before + 1
    assert torch.equal(keys, dev.hamming_scan_keys(ds, dq, 10, variant=1))
    with dev.force_popc():
        assert torch.equal(dev.hamming_scan_keys(ds, dq, 10), keys)
    # no overflow: the predicated kernels return at once and leave the tensor-core keys alone
    k_ok = dev.hamming_scan_keys(dt, dq, 10, variant=dev.SCAN_VARIANT_TC)
    assert dev.tc_scan_overflows() == before + 1 and torch.equal(k_ok, k_popc)


# ------------------------------------------------------------------ index build: sb_unique_codes
def _unique_case(dev, codes_np, rows_np=None):
    import torch
    codes = torch.from_numpy(np.ascontiguousarray(codes_np).view(np.int32)).cuda()
    rows = None if rows_np is None else torch.from_numpy(rows_np.astype(np.int64)).cuda()
    table, row_code, off, crow, mx = dev.unique_codes(codes, rows)
    torch.cuda.synchronize()
    sub = codes_np if rows_np is None else codes_np[rows_np]
    ot, oinv, ooff, orow = O.unique_code_table(sub)
    assert np.array_equal(table.cpu().numpy().view(np.uint32), ot)
    assert np.array_equal(off.cpu().numpy(), ooff)
    if rows_np is None:
        assert np.array_equal(row_code.cpu().numpy(), oinv)
        assert np.array_equal(crow.cpu().numpy(), orow)
    else:
        want = np.full(len(codes_np), -1, np.int64)
        want[rows_np] = oinv
        assert np.array_equal(row_code.cpu().numpy(), want)
        assert np.array_equal(crow.cpu().numpy(), rows_np[orow])
    assert mx == int(np.diff(ooff).max())


@pytest.mark.parametrize("n,W,distinct", [(1, 1, 1), (2, 8, 1), (100, 2, 7), (8192, 8, 8192), (8193, 4, 500),
                                          (100_003, 8, 100_003), (300_000, 1, 4000), (70_000, 32, 70_000),
                                          (50_000, 16, 333)])
def test_unique_codes_matches_oracle(dev, n, W, distinct):
    """sb_unique_codes (radix sort of a row permutation + boundaries + CSR, csrc/unique_codes.cu) vs the
    oracle's unique-code table (the array form of linear.py:163 / lsh.py:316-323): table, row -> code,
    CSR offsets and rows (ascending rows per code: the sort is stable), max rows per code."""
    rng = np.random.RandomState(n + W)
    pool = rng.randint(0, 2 ** 32, size=(distinct, W), dtype=np.uint64).astype(np.uint32)
    codes = pool[rng.randint(0, distinct, n)] if distinct < n else pool[rng.permutation(n)]
    _unique_case(dev, codes)


def test_unique_codes_zero_extended_clustered_and_subset(dev):
    """Digits every row agrees on are skipped (zero-extended codes), near-duplicate clusters that
    differ only in low words, all-equal codes, and an index over a subset of the rows (tombstones)."""
    rng = np.random.RandomState(5)
    n = 40_000
    low = rng.randint(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
    wide = np.zeros((n, 8), np.uint32)
    wide[:, 6:] = low                                          # a 64-bit code widened to 256 bits
    _unique_case(dev, wide)
    clustered = np.repeat(rng.randint(0, 2 ** 32, size=(40, 8), dtype=np.uint64).astype(np.uint32), n // 40, axis=0)
    clustered[:, 7] ^= rng.randint(0, 4, n).astype(np.uint32)  # clusters of 4 codes, differing in the last bits
    clustered = clustered[rng.permutation(n)]
    _unique_case(dev, clustered)
    _unique_case(dev, np.full((5000, 4), 0xDEADBEEF, np.uint32))
    live = np.sort(rng.choice(n, 17_001, replace=False))
    _unique_case(dev, clustered, live)
    _unique_case(dev, wide, np.array([n - 1]))


def test_hamming_empty_table(dev):
    q = dev.codes_to_device(np.zeros((2, 8), np.uint32))
    db = torch.empty((0, 8), dtype=torch.int32, device=q.device)
    d, i = dev.hamming_topk(db, q, 3)
    assert (d.cpu().numpy() == -1).all() and (i.cpu().numpy() == -1).all()


# --------------------------------------------------------------------- ITQ hash
#: |z_gpu - z_ref| <= HASH_EPS * |x - m|_2 * |r_j|_2 (fp32 inputs + fp32 FFMA
#: accumulation, D <= 4096); bits may only differ where |z_ref| is below that.
HASH_EPS = 1e-5


def _check_hash(dev, x, mean, rot, normalize, variant=1):
    b = rot.shape[1]
    W = (b + 31) // 32
    W = [w for w in (1, 2, 4, 8, 16, 32) if w >= W][0]
    X = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    m = torch.from_numpy(np.ascontiguousarray(mean, np.float32)).cuda()
    R = torch.from_numpy(np.ascontiguousarray(rot, np.float32)).cuda()
    codes, z = dev.itq_hash(X, m, R, normalize=normalize, want_z=True, variant=variant)
    torch.cuda.synchronize()
    z_ref = O.itq_project(np.asarray(x, np.float64), mean, rot, normalize)
    a = O.norm_vector(np.asarray(x, np.float64), normalize) - mean
    scale = np.linalg.norm(a, axis=1)[:, None] * np.linalg.norm(rot, axis=0)[None, :]
    err = np.abs(z.cpu().numpy() - z_ref)
    assert (err <= HASH_EPS * scale + 1e-30).all(), float((err / (scale + 1e-30)).max())
    bits_ref = z_ref >= 0
    got = dev.codes_to_host(codes)
    want = O.pack_codes(bits_ref, W)
    if not np.array_equal(got, want):
        from smqtk_indexing_b200.utils.bits import unpack_bits
        diff = unpack_bits(got, b) != bits_ref
        assert (np.abs(z_ref[diff]) <= HASH_EPS * scale[diff]).all()
    return np.mean(unpack_ok(got, want))


def unpack_ok(a, b):
    return a == b


def test_itq_hash_golden_models(dev, golden):
    g = golden("itq")
    for ci, (N, D, b, it, norm, seed, isz) in enumerate(g["cases"]):
        x, q = gi.itq_inputs(ci)
        normalize = None if norm < 0 else int(norm)
        for data in (x, q):
            _check_hash(dev, data, np.real(g["c%d_mean" % ci]), np.real(g["c%d_rot" % ci]), normalize)


def test_itq_hash_known_answer(dev):
    # reference tests/impls/lsh_functor/test_itq.py:304-336 incl. z == 0 -> True
    mean = np.zeros(2)
    rot = np.array([[1 / np.sqrt(2)], [1 / np.sqrt(2)]])
    pts = np.array([[1, 1], [-1, -1], [-1, 1], [-1.001, 1], [-1, 1.001], [1, -1], [1, -1.001], [1.001, -1]], float)
    X = torch.from_numpy(pts.astype(np.float32)).cuda()
    codes = dev.itq_hash(X, torch.zeros(2, device="cuda"), torch.from_numpy(rot.astype(np.float32)).cuda())
    assert list(dev.codes_to_host(codes)[:, 0]) == [1, 0, 1, 0, 1, 1, 0, 1]


@pytest.mark.parametrize("n,D,b", [(1, 2, 1), (3, 5, 5), (70, 33, 31), (129, 64, 33), (1000, 512, 256),
                                   (257, 100, 100), (65, 4096, 64), (100, 48, 1000)])
def test_itq_hash_shapes(dev, n, D, b):
    rng = np.random.RandomState(n + D + b)
    x = rng.rand(n, D).astype(np.float32)
    mean = x.mean(0).astype(np.float64)
    rot = np.linalg.qr(rng.randn(max(D, b), max(D, b)))[0][:D, :b]
    for normalize in (None, 2, 1, np.inf):
        _check_hash(dev, x, mean, rot, normalize)


@pytest.mark.parametrize("n,D,b", [(256, 16, 32), (1000, 512, 256), (777, 64, 64), (4096, 512, 256),
                                   (300, 128, 96), (5000, 256, 128), (40000, 48, 160), (513, 4096, 64)])
def test_itq_hash_tensor_core(dev, n, D, b):
    """tcgen05 3xTF32 kernel: same epsilon bar as the FFMA kernel, against the float64 oracle."""
    rng = np.random.RandomState(n + D + b)
    x = rng.rand(n, D).astype(np.float32)
    mean = x.mean(0).astype(np.float32).astype(np.float64)
    rot = np.linalg.qr(rng.randn(max(D, b), max(D, b)))[0][:D, :b]
    for normalize in (None, 2):
        _check_hash(dev, x, mean, rot, normalize, variant=2)


def test_itq_hash_tensor_core_matches_ffma(dev):
    """Both kernels give the same codes except where |z| is at rounding level; row
    pitch larger than D, zero rows and the z == 0 -> 1 rule included."""
    rng = np.random.RandomState(5)
    n, D, b = 3000, 128, 64
    big = torch.from_numpy(rng.rand(n, D + 32).astype(np.float32)).cuda()
    X = big[:, :D]
    X[7] = 0.0
    m = torch.zeros(D, device="cuda")
    R = torch.from_numpy(np.linalg.qr(rng.randn(D, D))[0][:, :b].astype(np.float32)).cuda()
    c1, z1 = dev.itq_hash(X, m, R, want_z=True, variant=1)
    c2, z2 = dev.itq_hash(X, m, R, want_z=True, variant=2)
    torch.cuda.synchronize()
    assert (dev.codes_to_host(c2)[7] == 0xFFFFFFFF).all()           # z == 0 -> bit set
    z1, z2 = z1.cpu().numpy(), z2.cpu().numpy()
    np.testing.assert_allclose(z2, z1, rtol=0, atol=1e-5)
    from smqtk_indexing_b200.utils.bits import unpack_bits
    d = unpack_bits(dev.codes_to_host(c1), b) != unpack_bits(dev.codes_to_host(c2), b)
    assert (np.abs(z1[d]) < 1e-5).all()


def test_itq_hash_tensor_core_rejects_unaligned(dev):
    X = torch.rand(600, 20, device="cuda")
    R = torch.rand(20, 32, device="cuda")
    assert dev.itq_rotation_image(R) is None
    with pytest.raises(ValueError):
        dev.itq_hash(X, None, R, variant=2)


# --------------------------------------------------------------------- re-rank
#: distances: |d_gpu - d_ref| <= 1e-5 * |d_ref| + atol (inputs are fp32-representable).
#: cosine's atol is the float64 reference's own conditioning floor: acos near 1
#: turns the last-ulp error of the similarity (1e-16) into sqrt(2e-16)*2/pi ~ 1e-8
#: (the reference itself returns 1.6e-8, not 0, for identical vectors).
RERANK_RTOL = 1e-5
RERANK_ATOL = {"euclidean": 1e-12, "hik": 1e-12, "cosine": 5e-8}


def test_rerank_metrics_golden(dev, golden):
    g = golden("metrics")
    for ci, D in enumerate(g["cases"]):
        q, c, qh, ch = gi.metrics_inputs(int(D))
        for metric, qq, cc in (("euclidean", q, c), ("cosine", q, c), ("hik", qh, ch)):
            q32 = qq.astype(np.float32)
            c32 = cc.astype(np.float32)
            db = torch.from_numpy(c32).cuda()
            tq = torch.from_numpy(q32[None, :]).cuda()
            idx = torch.arange(len(c32), device="cuda")
            off = torch.tensor([0, len(c32)], device="cuda")
            out = dev.rerank(db, tq, idx, off, metric).cpu().numpy()
            with np.errstate(all="ignore"):
                ref = np.atleast_1d(O.DISTANCE_FUNCTIONS[metric](q32.astype(np.float64), c32.astype(np.float64)))
            assert np.array_equal(np.isnan(out), np.isnan(ref))
            ok = ~np.isnan(ref)
            np.testing.assert_allclose(out[ok], ref[ok], rtol=RERANK_RTOL, atol=RERANK_ATOL[metric], err_msg=metric)


def test_rerank_known_answers(dev):
    # reference tests/impls/nn_index/test_lsh.py:102-136
    def one(metric, a, b):
        db = torch.tensor([b], dtype=torch.float32, device="cuda")
        q = torch.tensor([a], dtype=torch.float32, device="cuda")
        return float(dev.rerank(db, q, torch.zeros(1, dtype=torch.int64, device="cuda"),
                                torch.tensor([0, 1], device="cuda"), metric)[0])
    assert one("euclidean", [0, 0], [0, 1]) == 1.0
    assert abs(one("cosine", [1, 0], [0, 1]) - 1.0) < 1e-12
    assert abs(one("cosine", [1, 0], [1, 1]) - 0.5) < 1e-7
    assert one("hik", [0, 0], [0, 1]) == 1.0
    assert one("hik", [1, 0], [0, 1]) == 1.0
    assert one("hik", [1, 1], [0, 1]) == 0.0
    assert one("euclidean", [3, 4], [3, 4]) == 0.0
    assert one("cosine", [3, 4], [3, 4]) == 0.0


@pytest.mark.parametrize("tie_by_row", [False, True])
def test_rerank_select_large_lists_sorted_path(dev, tie_by_row):
    """candidate lists far beyond the rank-count kernel's shared-memory stage (m up to 1e5,
    fixed-pitch padding, ties, NaN) are ordered by the radix-sorted selection: (distance, position) or
    (distance, row) order exactly as numpy's stable lexsort."""
    rng = np.random.RandomState(23)
    sizes = [100_000, 0, 3, 40_000, 7000]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    M = int(off[-1])
    d = np.round(rng.rand(M), 3)                               # heavy ties
    d[off[3] + 5] = np.nan
    d[7] = -1.5                                                # negative values order before zero
    rows = rng.permutation(M).astype(np.int64)
    cnt = np.array([90_000, 0, 3, 40_000, 6500], np.int64)     # entries past cnt are padding
    n = 300
    t = lambda a: torch.from_numpy(a).cuda()
    if tie_by_row:
        got_r, got_d = dev.rerank_select_rows(t(d), t(off), t(cnt), t(rows), n, tie_by_row=True)
    else:
        got_r, got_d = dev.rerank_select(t(d), t(off), n)
        cnt = np.array(sizes, np.int64)
    got_r, got_d = got_r.cpu().numpy(), got_d.cpu().numpy()
    for qi in range(len(sizes)):
        m = int(cnt[qi])
        seg = d[off[qi]:off[qi] + m]
        key = np.where(np.isnan(seg), np.inf, seg)
        tie = rows[off[qi]:off[qi] + m] if tie_by_row else np.arange(m)
        o = np.lexsort((tie, key))[:n]
        want = rows[off[qi] + o] if tie_by_row else o + off[qi]
        assert np.array_equal(got_r[qi][:len(o)], want), qi
        assert (got_r[qi][len(o):] == -1).all()
        np.testing.assert_array_equal(got_d[qi][:len(o)], seg[o])


def test_rerank_select_order(dev):
    rng = np.random.RandomState(3)
    sizes = [0, 1, 5, 300, 5000, 17]
    off = np.concatenate([[0], np.cumsum(sizes)])
    d = rng.rand(off[-1])
    d[off[3]:off[3] + 50] = 0.25            # ties -> position order
    d[off[4] + 7] = np.nan                  # NaN sorts last
    pos, od = dev.rerank_select(torch.from_numpy(d).cuda(), torch.from_numpy(off).cuda(), 20)
    pos, od = pos.cpu().numpy(), od.cpu().numpy()
    for qi, m in enumerate(sizes):
        seg = d[off[qi]:off[qi + 1]]
        key = np.where(np.isnan(seg), np.inf, seg)
        o = np.lexsort((np.arange(m), key))[:20]
        assert list(pos[qi][:len(o)]) == list(o + off[qi])
        assert (pos[qi][len(o):] == -1).all()
        np.testing.assert_array_equal(od[qi][:len(o)], seg[o])


# --------------------------------------------------------------------- candidate expansion
@pytest.mark.parametrize("n", [1, 7, 130, 300])
def test_expand_candidates_fixed_pitch_equals_ragged(dev, n):
    """sb_expand_candidates (+ sb_rerank_select_rows) against the ragged torch expansion
    (+ position select): same candidates in the same (code rank, row) order."""
    from smqtk_indexing_b200 import engine
    rng = np.random.RandomState(n)
    U, Q = 500, 37
    counts = rng.randint(1, 6, size=U)
    counts[rng.randint(0, U, 5)] = 40                      # a few crowded codes
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rows = rng.permutation(int(off[-1])).astype(np.int64)
    code_rows = rng.randint(0, U, size=(Q, n)).astype(np.int64)
    code_rows[rng.rand(Q, n) < 0.1] = -1
    code_rows[3] = -1                                       # a query without candidates
    t = lambda a: torch.from_numpy(a).cuda()
    ci_r, co_r = engine.expand_candidates(t(code_rows), t(off), t(rows))
    pitch = n * int(counts.max())
    ci, co, cc = dev.expand_candidates(t(code_rows), t(off), t(rows), pitch)
    torch.cuda.synchronize()
    ci, co, cc, ci_r, co_r = (x.cpu().numpy() for x in (ci, co, cc, ci_r, co_r))
    assert np.array_equal(co, np.arange(Q + 1) * pitch)
    assert np.array_equal(cc, np.diff(co_r))
    for q in range(Q):
        seg = ci[q * pitch:(q + 1) * pitch]
        assert np.array_equal(seg[:cc[q]], ci_r[co_r[q]:co_r[q + 1]])
        assert (seg[cc[q]:] == -1).all()
    # selection on random distances (with ties and NaNs)
    d = rng.randint(0, 50, size=Q * pitch).astype(np.float64)
    d[rng.rand(Q * pitch) < 0.05] = np.nan
    k = min(n, 9)
    r_rows, r_d = dev.rerank_select_rows(t(d), t(co), t(cc), t(ci), k)
    pos, pd = dev.rerank_select(t(d), t(co), k)
    torch.cuda.synchronize()
    r_rows, r_d, pos, pd = (x.cpu().numpy() for x in (r_rows, r_d, pos, pd))
    for q in range(Q):
        m = min(k, cc[q])
        seg_d = d[q * pitch:q * pitch + cc[q]]
        o = np.lexsort((np.arange(cc[q]), np.where(np.isnan(seg_d), np.inf, seg_d)))[:m]
        assert np.array_equal(r_rows[q, :m], ci[q * pitch + o])
        np.testing.assert_array_equal(r_d[q, :m], seg_d[o])
        assert (r_rows[q, m:] == -1).all()


# --------------------------------------------------------------------- flat L2 index (row N1)
def _check_l2(dev, x, q, k):
    X = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    Qt = torch.from_numpy(np.ascontiguousarray(q, np.float32)).cuda()
    idx, dist = dev.l2_topk(X, Qt, k)
    torch.cuda.synchronize()
    orow, od = O.l2_topk(x.astype(np.float32), q.astype(np.float32), k)
    kk = od.shape[1]
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    np.testing.assert_allclose(dist[:, :kk], od, rtol=1e-5, atol=1e-12)
    assert (idx[:, kk:] == -1).all()
    for i in range(len(q)):
        # identical neighbours wherever the float64 distances are distinct at the 1e-9 level
        gaps = np.diff(od[i])
        if len(gaps) == 0 or gaps.min() > 1e-9 * max(od[i].max(), 1e-30):
            assert list(idx[i, :kk]) == list(orow[i]), i
        else:
            assert set(idx[i, :kk][od[i] < od[i][-1] - 1e-9]) <= set(orow[i])
    return idx, dist


@pytest.mark.parametrize("N,D,Q,k", [(5000, 128, 33, 10), (40000, 64, 300, 100), (3000, 16, 7, 256),
                                     (100, 32, 5, 10), (1, 16, 2, 3), (70000, 256, 64, 50)])
def test_l2_topk_matches_oracle(dev, N, D, Q, k):
    rng = np.random.RandomState(N + D)
    x = rng.rand(N, D).astype(np.float32)
    q = rng.rand(Q, D).astype(np.float32)
    q[0] = x[N // 2]                                           # an indexed vector: distance 0
    idx, dist = _check_l2(dev, x, q, k)
    assert idx[0, 0] == N // 2 and dist[0, 0] == 0.0


def test_l2_topk_duplicates_ties_and_brute_force_path(dev):
    rng = np.random.RandomState(9)
    x = rng.rand(6000, 48).astype(np.float32)
    x[100:140] = x[7]                                          # 41 identical rows: ties broken by row
    q = np.vstack([x[7], x[7] * 1.0001, rng.rand(3, 48)]).astype(np.float32)
    idx, dist = _check_l2(dev, x, q, 20)
    assert list(idx[0]) == [7] + list(range(100, 119)) and (dist[0] == 0).all()
    # D not a multiple of 16: the exact kernel over all rows
    xo = rng.rand(900, 37).astype(np.float32)
    _check_l2(dev, xo, rng.rand(6, 37).astype(np.float32), 12)


def test_l2_topk_overflow_falls_back(dev):
    """More near-ties than the survivor buffer holds: flagged and redone exactly."""
    rng = np.random.RandomState(2)
    x = np.tile(rng.rand(1, 32).astype(np.float32), (30000, 1))
    x[::7] += 1e-3 * rng.rand(len(x[::7]), 32).astype(np.float32)
    q = x[:3].copy()
    idx, dist = _check_l2(dev, x, q, 10)
