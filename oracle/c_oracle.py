"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Builds (gcc) and binds ``oracle/lsh_oracle.c``.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lsh_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liblsh_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    base = ["gcc", "-O3", "-mpopcnt", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    try:
        subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True)
    except (subprocess.CalledProcessError, FileNotFoundError):
        subprocess.run(base, check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = ctypes.CDLL(LIB)
        lib.oracle_max_threads.restype = c_int
        lib.oracle_hamming_topk.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int64,
                                            c_void_p, c_void_p, c_int]
        lib.oracle_itq_hash.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p, c_int]
        lib.oracle_distances.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def hamming_topk(db, q, k, idx_base=0, nthreads=0):
    db = np.ascontiguousarray(db, dtype=np.uint32)
    q = np.ascontiguousarray(np.atleast_2d(q), dtype=np.uint32)
    Q = q.shape[0]
    d = np.empty((Q, k), np.int32)
    i = np.empty((Q, k), np.int64)
    load().oracle_hamming_topk(db.ctypes.data, db.shape[0], db.shape[1], q.ctypes.data, Q, k, idx_base,
                               d.ctypes.data, i.ctypes.data, nthreads)
    return d, i


def itq_hash(x, mean, rot, norm_kind=0, want_z=False, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    rot = np.ascontiguousarray(rot, dtype=np.float64)
    n, D = x.shape
    b = rot.shape[1]
    bits = np.empty((n, b), np.uint8)
    z = np.empty((n, b), np.float64) if want_z else None
    load().oracle_itq_hash(x.ctypes.data, n, D, mean.ctypes.data, rot.ctypes.data, b, norm_kind, bits.ctypes.data,
                           z.ctypes.data if want_z else None, nthreads)
    return (bits.astype(bool), z) if want_z else bits.astype(bool)


def distances(q, rows, metric):
    q = np.ascontiguousarray(q, dtype=np.float32)
    rows = np.ascontiguousarray(np.atleast_2d(rows), dtype=np.float32)
    out = np.empty(rows.shape[0], np.float64)
    load().oracle_distances(q.ctypes.data, rows.ctypes.data, rows.shape[0], rows.shape[1],
                            {"euclidean": 0, "cosine": 1, "hik": 2}[metric], out.ctypes.data)
    return out
