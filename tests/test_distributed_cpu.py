"""
world_size-2 (and 3) gloo tests of the multi-GPU plumbing on CPU: row-range
sharding, code all-gather, range-partitioned scan, key all-gather + merge,
candidate-distance exchange.  The per-rank compute steps are injected (oracle
backed, CPU) -- the point is that the sharded result is identical to the
single-index result, ties included.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import np_oracle as O


class OracleOps:
    """CPU stand-in for DeviceOps (tests only)."""

    def __init__(self, mean, rot, metric):
        self.mean, self.rot, self.metric = mean, rot, metric

    def hash(self, x):
        bits = O.itq_hash(x.numpy().astype(np.float64), self.mean, self.rot)
        return torch.from_numpy(O.pack_codes(bits, 1).view(np.int32).copy())

    def scan_keys(self, table, q_codes, n, idx_base):
        t = table.numpy().view(np.uint32)
        q = q_codes.numpy().view(np.uint32)
        keys = np.full((len(q), n), -1, np.int64)                  # 0xFFFF... = empty
        if len(t):
            d, i = O.hamming_topk(t, q, n, idx_base=idx_base)
            keys[:, :d.shape[1]] = (d.astype(np.int64) << 40) | i
        return torch.from_numpy(keys)

    def merge_keys(self, keys):
        k = keys.numpy().view(np.uint64)
        P, Q, n = k.shape
        od = np.full((Q, n), -1, np.int32)
        oi = np.full((Q, n), -1, np.int64)
        for q in range(Q):
            c = np.sort(k[:, q, :].ravel())
            c = c[c != np.uint64(0xFFFFFFFFFFFFFFFF)][:n]
            od[q, :len(c)] = (c >> np.uint64(40)).astype(np.int32)
            oi[q, :len(c)] = (c & np.uint64((1 << 40) - 1)).astype(np.int64)
        return torch.from_numpy(od), torch.from_numpy(oi)

    def rerank(self, x, q, cand_idx, cand_off):
        xn, qn = x.numpy().astype(np.float64), q.numpy().astype(np.float64)
        ci, co = cand_idx.numpy(), cand_off.numpy()
        out = np.full(len(ci), np.nan)
        for qi in range(len(qn)):
            for j in range(co[qi], co[qi + 1]):
                if 0 <= ci[j] < len(xn):
                    out[j] = O.DISTANCE_FUNCTIONS[self.metric](qn[qi], xn[ci[j]])
        return torch.from_numpy(out)

    def rerank_select(self, d, cand_off, n):
        dn, co = d.numpy(), cand_off.numpy()
        Q = len(co) - 1
        pos = np.full((Q, n), -1, np.int64)
        od = np.full((Q, n), np.nan)
        for qi in range(Q):
            seg = dn[co[qi]:co[qi + 1]]
            o = np.lexsort((np.arange(len(seg)), np.where(np.isnan(seg), np.inf, seg)))[:n]
            pos[qi, :len(o)] = o + co[qi]
            od[qi, :len(o)] = seg[o]
        return torch.from_numpy(pos), torch.from_numpy(od)


def _data():
    rng = np.random.RandomState(4)
    N, D, b = 1500, 16, 7                     # 7 bits: many descriptors share a code
    x = rng.rand(N, D).astype(np.float32)
    q = rng.rand(9, D).astype(np.float32)
    mean = x.mean(0).astype(np.float64)
    rot = np.linalg.qr(rng.randn(D, D))[0][:, :b]
    return x, q, mean, rot


def _worker(rank, world, port, cuts, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smqtk_indexing_b200.distributed import ShardedLshIndex
        x, q, mean, rot = _data()
        idx = ShardedLshIndex(None, "euclidean", ops=OracleOps(mean, rot, "euclidean"))
        idx.build(torch.from_numpy(x[cuts[rank]:cuts[rank + 1]]))
        res = {}
        for n in (1, 5, 40):
            rows, d = idx.query(torch.from_numpy(q), n)
            res[n] = (rows.numpy(), d.numpy())
            # the scan partitioned over the queries (9 queries over 2 / 3 ranks: ragged slices) gives the same answer
            idx.scan_partition = "queries"
            rows_q, d_q = idx.query(torch.from_numpy(q), n)
            idx.scan_partition = "auto"
            assert torch.equal(rows_q, rows) and torch.equal(d_q, d)
        # fewer queries than ranks: the trailing ranks scan a padding row and contribute nothing
        idx.scan_partition = "queries"
        rows_1, d_1 = idx.query(torch.from_numpy(q[:1]), 5)
        idx.scan_partition = "rows"
        rows_r, d_r = idx.query(torch.from_numpy(q[:1]), 5)
        idx.scan_partition = "auto"
        assert torch.equal(rows_1, rows_r) and torch.equal(d_1, d_r)
        assert torch.equal(rows_1[0], torch.from_numpy(res[5][0][0])) and torch.equal(d_1[0], torch.from_numpy(res[5][1][0]))
        out_q.put((rank, idx.num_rows, idx.num_codes, (idx.scan_lo, idx.scan_hi), res))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,cuts", [(2, [0, 700, 1500]), (3, [0, 10, 1490, 1500])])
def test_sharded_index_equals_single_index(world, cuts):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cuts, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-index truth from the oracle's array-form pipeline
    x, q, mean, rot = _data()
    codes = O.pack_codes(O.itq_hash(x.astype(np.float64), mean, rot), 1)
    table = O.unique_code_table(codes)[0]
    slices = sorted(r[3] for r in results)
    assert slices[0][0] == 0 and slices[-1][1] == len(table)
    assert all(a[1] == b[0] for a, b in zip(slices[:-1], slices[1:]))
    for rank, nrows, ncodes, _, res in results:
        assert nrows == len(x) and ncodes == len(table)
        for n, (rows, d) in res.items():
            for qi in range(len(q)):
                qw = O.pack_codes(O.itq_hash(q[qi].astype(np.float64), mean, rot), 1)
                orows, od = O.lsh_nn(x.astype(np.float64), codes, q[qi].astype(np.float64), qw, n, "euclidean")
                assert list(rows[qi][:len(orows)]) == list(orows), (rank, n, qi)
                np.testing.assert_array_equal(d[qi][:len(od)], od)
                assert (rows[qi][len(orows):] == -1).all()


def test_partition_bounds():
    from smqtk_indexing_b200.distributed import partition_bounds
    for total in (0, 1, 3, 4, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            c = partition_bounds(total, world)
            assert len(c) == world + 1 and c[0] == 0 and c[-1] == total
            assert all(a <= b for a, b in zip(c[:-1], c[1:]))
            assert all(x % 4 == 0 for x in c[1:-1])


# ----------------------------------------------------------------------------- sharded flat L2 index
class OracleFlatOps:
    """CPU stand-in for DeviceFlatOps (tests only)."""

    def prepare(self, x):
        return None

    def l2_topk(self, x, q, n, prepared):
        rows, d = O.l2_topk(x.numpy(), q.numpy(), n)
        idx = np.full((len(q), n), -1, np.int64)
        dd = np.full((len(q), n), np.nan)
        idx[:, :rows.shape[1]] = rows
        dd[:, :d.shape[1]] = d
        return torch.from_numpy(idx), torch.from_numpy(dd)

    def select(self, dist_all, idx_all, n):
        d, i = dist_all.numpy(), idx_all.numpy()
        Q = d.shape[0]
        rows = np.full((Q, n), -1, np.int64)
        od = np.full((Q, n), np.nan)
        for q in range(Q):
            ok = np.flatnonzero(i[q] >= 0)
            o = ok[np.lexsort((i[q][ok], d[q][ok]))][:n]
            rows[q, :len(o)] = i[q][o]
            od[q, :len(o)] = d[q][o]
        return torch.from_numpy(rows), torch.from_numpy(od)


def _flat_worker(rank, world, port, cuts, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smqtk_indexing_b200.distributed import ShardedFlatL2Index
        x, q, _, _ = _data()
        x[40:60] = x[3]                                  # duplicates spread over the shards
        idx = ShardedFlatL2Index(ops=OracleFlatOps())
        idx.build(torch.from_numpy(x[cuts[rank]:cuts[rank + 1]]))
        res = {n: tuple(t.numpy() for t in idx.query(torch.from_numpy(q), n)) for n in (1, 7, 30)}
        out_q.put((rank, idx.num_rows, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cuts", [(2, [0, 50, 1500]), (3, [0, 45, 45, 1500])])
def test_sharded_flat_l2_equals_single_index(world, cuts):
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, cuts, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, q, _, _ = _data()
    x[40:60] = x[3]
    for rank, nrows, res in results:
        assert nrows == len(x)
        for n, (rows, d) in res.items():
            orow, od = O.l2_topk(x, q, n)
            assert np.array_equal(rows, orow), (rank, n)
            np.testing.assert_array_equal(d, od)
