"""
ctypes binding of ``libsmqtk_b200.so`` (declared in ``include/smqtk_b200.h``).

The library is the ONLY implementation of the compute path: there is no numpy
or torch fallback.  If it is missing (not built) every entry point raises.
"""
import ctypes
import os
import re
import threading
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from typing import Dict, List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libsmqtk_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "smqtk_b200.h")

SB_OK = 0
KEY_EMPTY = 0xFFFFFFFFFFFFFFFF
KEY_ROW_BITS = 40
NORM_NONE, NORM_LP, NORM_INF, NORM_L0 = 0, 1, 2, 3
METRIC_EUCLIDEAN, METRIC_COSINE, METRIC_HIK = 0, 1, 2
METRICS = {"euclidean": METRIC_EUCLIDEAN, "cosine": METRIC_COSINE, "hik": METRIC_HIK}

_P = c_void_p
# name -> (restype, argtypes); mirrors include/smqtk_b200.h one to one
SIGNATURES: Dict[str, tuple] = {
    "sb_version": (c_int32, []),
    "sb_last_error": (c_char_p, []),
    "sb_launch_count": (c_uint64, []),
    "sb_profile_enable": (c_int32, [c_int32]),
    "sb_profile_fetch": (c_int64, [_P, _P, c_int64]),
    "sb_itq_hash": (c_int32, [_P, c_int64, c_int32, c_int64, _P, _P, c_int32, c_int32, c_float,
                              _P, c_int32, _P, c_int32, _P]),
    "sb_itq_row_div": (c_int32, [_P, c_int64, c_int32, c_int64, c_int32, c_float, _P, _P]),
    "sb_itq_rotation_image_bytes": (c_size_t, [c_int32, c_int32]),
    "sb_itq_rotation_image": (c_int32, [_P, c_int32, c_int32, _P, _P]),
    "sb_itq_hash_tc": (c_int32, [_P, c_int64, c_int32, c_int64, _P, _P, c_int32, _P, _P, c_int32, _P, _P]),
    "sb_hamming_scan_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "sb_hamming_scan_tc_supported": (c_int32, [c_int64, c_int32, c_int32, c_int32]),
    "sb_hamming_scan_tc_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "sb_hamming_scan_tc": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sb_hamming_scan_tc4_supported": (c_int32, [c_int64, c_int32, c_int32, c_int32]),
    "sb_hamming_scan_tc4_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "sb_hamming_scan_tc4": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sb_hamming_scan": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, c_size_t, _P]),
    "sb_hamming_scan_variant": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P,
                                          c_size_t, c_int32, _P]),
    "sb_hamming_scan_if": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sb_topk_merge": (c_int32, [_P, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "sb_hamming_topk": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sb_rerank": (c_int32, [_P, c_int64, c_int32, c_int64, _P, c_int32, c_int64, _P, _P, c_int64, c_int32, _P, _P]),
    "sb_rerank_shard": (c_int32, [_P, c_int64, c_int64, c_int32, c_int64, _P, c_int32, c_int64, _P, _P, c_int64,
                                  c_int32, _P, _P]),
    "sb_rerank_peer": (c_int32, [_P, _P, c_int32, c_int32, c_int64, _P, c_int32, c_int64, _P, _P, c_int64, c_int64, c_int32,
                                 c_int32, _P, _P]),
    "sb_rerank_pitched": (c_int32, [_P, c_int64, c_int32, c_int64, _P, c_int32, c_int64, _P, _P, c_int64, c_int32, _P, _P]),
    "sb_enable_peer_access": (c_int32, [c_int32]),
    "sb_ipc_export": (c_int32, [_P, _P, _P]),
    "sb_ipc_import": (c_int32, [_P, _P]),
    "sb_ipc_release": (c_int32, [_P]),
    "sb_rerank_select": (c_int32, [_P, _P, c_int32, c_int32, _P, _P, _P]),
    "sb_rerank_select_rows": (c_int32, [_P, _P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P]),
    "sb_rerank_base": (c_int32, [_P, c_int64, c_int64, c_int32, c_int64, _P, c_int32, c_int64, _P, _P, c_int64,
                                 c_int32, _P, _P]),
    "sb_expand_candidates": (c_int32, [_P, c_int32, c_int32, _P, _P, c_int64, _P, _P, _P, _P]),
    "sb_unique_codes_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "sb_unique_codes": (c_int32, [_P, c_int64, c_int32, _P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "sb_rerank_select_sorted_workspace_bytes": (c_size_t, [c_int64]),
    "sb_rerank_select_sorted": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P, _P, c_size_t, _P]),
    "sb_hamming_topk_sorted_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "sb_hamming_topk_sorted": (c_int32, [_P, c_int64, c_int32, _P, c_int32, c_int32, c_int64, _P, _P, c_size_t, _P]),
    "sb_l2_prepare": (c_int32, [_P, c_int64, c_int32, c_int64, _P, _P, _P]),
    "sb_l2_topk_supported": (c_int32, [c_int64, c_int32, c_int64, c_int32]),
    "sb_l2_topk_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "sb_l2_topk": (c_int32, [_P, c_int64, c_int32, c_int64, _P, _P, _P, c_int32, c_int64, c_int32,
                             _P, _P, _P, _P, c_size_t, _P]),
    "sb_fit_row_div": (c_int32, [_P, c_int32, c_int64, c_int32, c_int64, c_int32, c_double, _P, _P]),
    "sb_fit_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "sb_fit_col_mean": (c_int32, [_P, c_int32, c_int64, c_int32, c_int64, _P, _P, _P, c_size_t, _P]),
    "sb_fit_gram": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, c_int32,
                              _P, c_int32, c_int64, c_int32, _P, _P, c_int32,
                              c_int64, c_double, _P, _P, c_size_t, _P]),
    "sb_fit_gram_bits_tc_supported": (c_int32, [c_int64, c_int32, c_int64, c_int32]),
    "sb_fit_gram_bits_tc_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "sb_fit_gram_bits_tc": (c_int32, [_P, c_int32, c_int32, _P, c_int64, c_int32, c_int64, _P, _P, c_float, _P, _P, c_size_t, _P]),
    "sb_fit_project": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, _P, c_int32, c_int64, _P, _P, c_int32, _P]),
}

_lock = threading.Lock()
_lib = None


class LibraryMissingError(RuntimeError):
    pass


def declared_symbols() -> List[str]:
    """Function names declared (``SB_API``) in include/smqtk_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return re.findall(r"^SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", text, flags=re.M)


def load() -> ctypes.CDLL:
    """Load the C-ABI library (once).  Raises ``LibraryMissingError`` when it has
    not been built -- by design there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise LibraryMissingError(
                    "%s not found: build it with `python -m smqtk_indexing_b200.build` "
                    "(or __graft_entry__.build()).  The CUDA library is the only "
                    "implementation of the LSH hot path; there is no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


ENV_ASSUME_USABLE = "SMQTK_B200_ASSUME_USABLE"
_usable = None


def usable() -> bool:
    """``is_usable()`` of every plugin class: the C-ABI library is present AND a CUDA device is visible
    (the reference's optional impls answer the same question about their imports, e.g.
    smqtk_indexing/impls/hash_index/sklearn_balltree.py:43-45).  ``SMQTK_B200_ASSUME_USABLE=1`` skips the
    device half only -- for host-logic tests and config generation in a container without a GPU; compute
    entry points still raise there."""
    global _usable
    if not os.path.exists(LIB_PATH):
        return False
    if os.environ.get(ENV_ASSUME_USABLE, "") not in ("", "0"):
        return True
    if _usable is None:
        try:
            import torch
            _usable = bool(torch.cuda.is_available())
        except Exception:
            _usable = False
    return _usable


def check(rc: int) -> None:
    if rc != SB_OK:
        msg = load().sb_last_error()
        raise RuntimeError("libsmqtk_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().sb_launch_count())


def profile_enable(on: bool) -> None:
    """Bracket every kernel launch of the library with CUDA events (bench only)."""
    load().sb_profile_enable(1 if on else 0)


def profile_fetch(cap: int = 1 << 16):
    """[(kernel name, milliseconds)] of the launches recorded since the last fetch."""
    names = (c_char_p * cap)()
    ms = (c_float * cap)()
    n = int(load().sb_profile_fetch(ctypes.cast(names, c_void_p), ctypes.cast(ms, c_void_p), cap))
    return [(names[i].decode(), float(ms[i])) for i in range(min(n, cap))]
