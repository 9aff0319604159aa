"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Seeded input generators shared by ``oracle/gen_golden.py`` (which feeds them to
the unmodified reference) and the tests (which feed them to the oracle and the
CUDA path), so fixtures only need to store the reference's OUTPUTS.
``numpy.random.RandomState`` streams are frozen across numpy versions.
"""
import numpy as np

LINEAR_NN_CASES = [  # (bits, rows, seed)
    (3, 8, 1), (5, 20, 2), (8, 200, 3), (32, 1500, 4), (64, 2000, 5),
    (256, 2000, 6), (100, 700, 7), (1000, 300, 8),
]


def linear_nn_inputs(b, U, seed):
    rng = np.random.RandomState(seed)
    db = rng.rand(U, b) > 0.5
    qs = rng.rand(6, b) > 0.5
    qs[1] = db[U // 2]                          # exact hit
    qs[2] = db[U // 3].copy()
    qs[2][0] ^= True                            # one bit away from a stored code
    return db, qs


def linear_nn_ns(unique_count):
    ns = {1, 4, 10, min(50, unique_count)}
    if unique_count <= 300:
        ns.add(unique_count)
    return sorted(ns)


ITQ_CASES = [  # (N, D, bits, iterations, normalize, seed, dtype)
    (5, 2, 1, 50, None, 0, "f8"),               # tests/impls/lsh_functor/test_itq.py:255-270
    (500, 16, 8, 50, None, 0, "f8"),
    (400, 32, 16, 20, 2, 3, "f8"),
    (600, 24, 5, 10, 1, 7, "f4"),
    (2000, 128, 64, 50, None, 0, "f8"),
    (1500, 64, 64, 5, None, 11, "f4"),
]


def itq_inputs(ci):
    N, D, b, it, norm, seed, dt = ITQ_CASES[ci]
    rng = np.random.RandomState(1000 + ci)
    if N == 5:
        x = np.array([[-2. + i, -2. + i] for i in range(5)])
        # tests/impls/lsh_functor/test_itq.py:304-336 query points
        q = np.array([[1, 1], [-1, -1], [-1, 1], [-1.001, 1], [-1, 1.001],
                      [1, -1], [1, -1.001], [1.001, -1]], float)
    else:
        x = rng.rand(N, D).astype(dt)
        q = rng.rand(64, D).astype(dt)
    return x, q


METRIC_DIMS = (2, 5, 128, 512, 4096)


def metrics_inputs(D):
    rng = np.random.RandomState(300 + D)
    q = rng.rand(D)
    c = rng.rand(40, D)
    c[0] = q                                    # identical
    c[1] = q * (1 + 1e-4 * rng.rand(D))         # near duplicate
    c[2] = 0.0                                  # zero vector
    c[3] = q * 3.0                              # parallel
    qh = q / q.sum()
    ch = c.copy()
    ch[2] = rng.rand(D)
    ch = ch / ch.sum(axis=1, keepdims=True)
    return q, c, qh, ch


LSH_CASES = [  # (N, D, bits, method, use_hash_index, seed)
    (1000, 256, 32, "euclidean", 1, 0),         # tests/impls/nn_index/test_lsh.py:754-832 shape
    (1000, 256, 32, "euclidean", 0, 0),
    (1000, 256, 32, "cosine", 1, 0),
    (800, 64, 16, "hik", 1, 5),
    (1200, 32, 6, "euclidean", 1, 9),           # few bits: heavy code collisions
]
LSH_NS = (1, 10, 50)


def lsh_inputs(ci):
    N, D, b, method, use_hi, seed = LSH_CASES[ci]
    rng = np.random.RandomState(seed)
    x = rng.rand(N, D)
    if method == "hik":
        x = x / x.sum(axis=1, keepdims=True)
    qs = rng.rand(8, D)
    qs[0] = x[255]
    qs[1] = x[0] * (1 + 0.01 * rng.rand(D))
    if method == "hik":
        qs = qs / qs.sum(axis=1, keepdims=True)
    return x, qs


# ---- incremental ingest: one build + update / remove steps through the reference's
# LSHNearestNeighborIndex.update_index / remove_from_index (lsh.py:331-450)
INGEST_SHAPE = (900, 32, 16, 11)                # rows, D, bits, seed
INGEST_N = 6                                    # neighbours asked per query
#: (op, first row, one-past-last row, stride): rows double as uuids; "add" = update_index (build_index
#: for the first step), "del" = remove_from_index.  Step 4 brings removed uuids back, step 6 re-adds
#: live uuids with their unchanged vectors (a no-op in the reference's set semantics).
INGEST_STEPS = [
    ("add", 0, 400, 1),
    ("add", 400, 650, 1),
    ("del", 3, 600, 4),
    ("add", 650, 900, 1),
    ("add", 3, 200, 8),
    ("del", 600, 900, 3),
    ("add", 100, 140, 1),
]


def ingest_inputs():
    """Clustered float32 descriptors (codes collide: several uuids per hash) and queries."""
    N, D, b, seed = INGEST_SHAPE
    rng = np.random.RandomState(seed)
    centres = rng.rand(45, D)
    x = (centres[rng.randint(0, 45, N)] + 0.04 * rng.randn(N, D)).astype(np.float32)
    qs = (centres[rng.randint(0, 45, 12)] + 0.04 * rng.randn(12, D)).astype(np.float32)
    qs[0] = x[1]
    return x, qs


def ingest_step_rows(step):
    op, lo, hi, stride = step
    return op, list(range(lo, hi, stride))
