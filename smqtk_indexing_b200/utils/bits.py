"""
Host-side marshalling between the three spellings of a hash code:

* ``bool[b]`` bit vector, index 0 = most significant bit (the plugin API type;
  reference: smqtk_indexing/impls/lsh_functor/itq.py:46-49),
* arbitrary precision Python ``int`` (the reference's storage type;
  smqtk_indexing/utils/bits.py:4-56),
* ``uint32[W]`` packed words -- the device layout.  A row is the *integer
  value* written big-endian in ``W`` 32-bit words (word 0 most significant),
  i.e. bit ``j`` of a ``b``-bit vector is integer bit ``p = b-1-j`` and lives in
  word ``W-1-p//32`` at position ``p%32``.  Rows therefore sort like the
  reference's ints, and widening ``W`` only prepends zero words.

These are API-boundary conversions (O(bytes) numpy ops), not the compute path:
bulk hashing packs its bits on the GPU in the hash kernel's epilogue.
"""
from typing import Iterable, List, Sequence

import numpy as np

#: Word counts the scan kernels are specialised for (bits = 32 * W).
SUPPORTED_WORDS = (1, 2, 4, 8, 16, 32)
MAX_BITS = 32 * SUPPORTED_WORDS[-1]


def words_for_bits(bits: int) -> int:
    """Smallest supported word count holding ``bits`` bits."""
    if bits < 1:
        raise ValueError("bit length must be positive, got %d" % bits)
    for w in SUPPORTED_WORDS:
        if bits <= 32 * w:
            return w
    raise ValueError("bit length %d exceeds the supported maximum of %d"
                     % (bits, MAX_BITS))


def bit_vector_to_int_large(v: Sequence) -> int:
    """``bool[b]`` (any truthy/falsy sequence) -> Python int, index 0 = MSB."""
    a = np.asarray(v)
    if a.ndim != 1:
        raise ValueError("expected a 1D bit vector")
    if a.size == 0:
        return 0
    by = np.packbits(np.concatenate(
        [np.zeros((-a.size) % 8, np.bool_), a.astype(np.bool_)]))
    return int.from_bytes(by.tobytes(), "big")


def int_to_bit_vector_large(integer: int, bits: int = 0) -> np.ndarray:
    """Python int -> ``bool`` vector of ``bits`` entries (or the minimum needed,
    at least one).

    :raises ValueError: ``bits`` is too small for ``integer``.
    """
    integer = int(integer)
    size = max(integer.bit_length(), 1)
    if bits and bits < size:
        raise ValueError("%d bits too small to represent integer value %d."
                         % (bits, integer))
    n = bits or size
    nbytes = (n + 7) // 8
    by = np.frombuffer(integer.to_bytes(nbytes, "big"), dtype=np.uint8)
    return np.unpackbits(by)[8 * nbytes - n:].astype(np.bool_)


def pack_bits(bitmat: np.ndarray, words: int) -> np.ndarray:
    """``bool[n, b]`` -> ``uint32[n, words]`` (layout: module docstring)."""
    bm = np.asarray(bitmat)
    if bm.ndim == 1:
        bm = bm[None, :]
    bm = bm.astype(np.bool_, copy=False)
    n, b = bm.shape
    if b > 32 * words:
        raise ValueError("%d bits do not fit in %d words" % (b, words))
    full = np.zeros((n, 32 * words), np.bool_)
    if b:
        full[:, 32 * words - b:] = bm
    by = np.packbits(full, axis=1)                      # [n, 4W] big-endian bytes
    return np.ascontiguousarray(by.view(">u4").astype(np.uint32))


def unpack_bits(wordmat: np.ndarray, bits: int) -> np.ndarray:
    """``uint32[n, W]`` -> ``bool[n, bits]``.

    :raises ValueError: a row has a set bit above position ``bits``.
    """
    wm = np.ascontiguousarray(np.asarray(wordmat, dtype=np.uint32))
    if wm.ndim == 1:
        wm = wm[None, :]
    n, w = wm.shape
    full = np.unpackbits(wm.astype(">u4").view(np.uint8), axis=1)   # [n, 32W]
    if bits < 32 * w and full[:, :32 * w - bits].any():
        raise ValueError("%d bits too small to represent a stored code" % bits)
    if bits > 32 * w:
        out = np.zeros((n, bits), np.bool_)
        out[:, bits - 32 * w:] = full
        return out
    return full[:, 32 * w - bits:].astype(np.bool_)


def ints_to_words(ints: Iterable[int], words: int) -> np.ndarray:
    """Python ints -> ``uint32[n, words]``."""
    ints = list(ints)
    buf = bytearray(4 * words * len(ints))
    step = 4 * words
    for i, v in enumerate(ints):
        buf[i * step:(i + 1) * step] = int(v).to_bytes(step, "big")
    a = np.frombuffer(bytes(buf), dtype=">u4").reshape(len(ints), words)
    return np.ascontiguousarray(a.astype(np.uint32))


def words_to_ints(wordmat: np.ndarray) -> List[int]:
    """``uint32[n, W]`` -> list of Python ints."""
    wm = np.ascontiguousarray(np.asarray(wordmat, dtype=np.uint32))
    if wm.ndim == 1:
        wm = wm[None, :]
    raw = wm.astype(">u4").tobytes()
    step = 4 * wm.shape[1]
    return [int.from_bytes(raw[i:i + step], "big") for i in range(0, len(raw), step)]


def widen_words(wordmat: np.ndarray, words: int) -> np.ndarray:
    """Re-express ``uint32[n, W0]`` rows in ``words >= W0`` words (zero-extend)."""
    wm = np.asarray(wordmat, dtype=np.uint32)
    n, w0 = wm.shape
    if words < w0:
        raise ValueError("cannot narrow %d words to %d" % (w0, words))
    if words == w0:
        return np.ascontiguousarray(wm)
    out = np.zeros((n, words), np.uint32)
    out[:, words - w0:] = wm
    return out
