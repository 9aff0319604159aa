"""
Bookkeeping for tables of packed hash codes: sort / unique / union / difference on
``int32[rows, W]`` tensors that carry uint32 bit patterns (word 0 most significant).

CUDA tensors go through ``sb_unique_codes`` (csrc/unique_codes.cu: one radix-sort + flag + scan +
scatter pipeline, no library sort) -- the index-build path of LinearHashIndex / LSHNearestNeighborIndex
(reference linear.py:148-204, lsh.py:316-329).  HOST tensors (the world-size-2 gloo tests of the
sharded index, which run the collective logic without a GPU) take a small torch restatement of the
same contract; it is never used for CUDA tensors.
"""
from typing import Optional, Tuple

import torch

_SIGN = -2 ** 31


def _flip(t: torch.Tensor) -> torch.Tensor:
    return torch.bitwise_xor(t, torch.tensor(_SIGN, dtype=torch.int32, device=t.device))


def lex_order(codes: torch.Tensor) -> torch.Tensor:
    """Permutation that sorts the rows ascending by (unsigned) integer value; STABLE, so
    rows with equal codes keep their relative order.  LSD radix over 64-bit digits:
    W/2 stable device sorts of int64 keys instead of a comparison sort over rows."""
    n, w = codes.shape
    order = None
    mask = 0xFFFFFFFF
    for j in range(w, 0, -2):                       # least significant word pair first
        lo = codes[:, j - 1].to(torch.int64) & mask
        if j >= 2:
            key = ((codes[:, j - 2].to(torch.int64) & mask) << 32) | lo
        else:
            key = lo
        key = key ^ (-2 ** 63)                       # unsigned order through a signed sort
        if order is not None:
            key = key[order]
        idx = torch.sort(key, stable=True).indices
        order = idx if order is None else order[idx]
    return order


def build_table(codes: torch.Tensor, rows: Optional[torch.Tensor] = None, with_max: bool = False):
    """Everything the index derives from per-row codes, in one pass:

    :param rows: int64 ascending row numbers to index (live rows); None = every row.
    :return: (table int32[U, W] sorted unique, row_code int64[all rows] -> table row (-1 for rows
              not in ``rows``), csr_off int64[U + 1], csr_rows int64[len(rows)]) with the rows of each
              code in ascending row order (+ the largest number of rows on one code when ``with_max``).
    """
    if codes.dim() != 2 or codes.dtype != torch.int32:
        raise ValueError("codes must be int32[rows, W]")
    n = codes.shape[0] if rows is None else int(rows.numel())
    dev = codes.device
    if n == 0:
        e = torch.empty(0, dtype=torch.int64, device=dev)
        rc = torch.full((codes.shape[0],), -1, dtype=torch.int64, device=dev)
        out = (codes[:0].clone(), rc, torch.zeros(1, dtype=torch.int64, device=dev), e)
        return out + (0,) if with_max else out
    if codes.is_cuda:
        from . import device
        table, row_code, csr_off, csr_rows, mx = device.unique_codes(codes.contiguous(), rows)
        return (table, row_code, csr_off, csr_rows, mx) if with_max else (table, row_code, csr_off, csr_rows)
    # ---- host tensors only (gloo tests of the collective logic) ----
    if rows is not None:
        t, rc, off, cr = build_table(codes[rows].contiguous())
        row_code = torch.full((codes.shape[0],), -1, dtype=torch.int64, device=dev)
        row_code[rows] = rc
        out = (t, row_code, off, rows[cr])
        return out + (int((off[1:] - off[:-1]).max()),) if with_max else out
    order = lex_order(codes)
    s = codes[order]
    new = torch.ones(n, dtype=torch.bool, device=dev)
    if n > 1:
        new[1:] = (s[1:] != s[:-1]).any(dim=1)
    first = torch.nonzero(new).reshape(-1)           # position of each code's first row
    table = s[first].contiguous()
    code_of_sorted = torch.cumsum(new.to(torch.int64), 0) - 1
    row_code = torch.empty(n, dtype=torch.int64, device=dev)
    row_code[order] = code_of_sorted
    csr_off = torch.empty(first.numel() + 1, dtype=torch.int64, device=dev)
    csr_off[:-1] = first
    csr_off[-1] = n
    if with_max:
        return table, row_code, csr_off, order, int((csr_off[1:] - csr_off[:-1]).max())
    return table, row_code, csr_off, order


def sort_unique(codes: torch.Tensor, return_inverse: bool = False, return_counts: bool = False):
    """Rows of ``codes`` sorted ascending by integer value, duplicates removed.

    :return: table int32[U, W] (+ inverse int64[rows] -> table row, + counts int64[U])
    """
    table, row_code, csr_off, _ = build_table(codes)
    out = [table]
    if return_inverse:
        out.append(row_code)
    if return_counts:
        out.append(csr_off[1:] - csr_off[:-1])
    return out[0] if len(out) == 1 else tuple(out)


def union(table: torch.Tensor, new_codes: torch.Tensor) -> torch.Tensor:
    """Sorted-unique union of two code sets (same W)."""
    return sort_unique(torch.cat([table, new_codes], dim=0))


def difference(table: torch.Tensor, remove: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``table`` minus ``remove``.

    :return: (remaining table, missing int32[m, W]) where ``missing`` lists the
        rows of ``remove`` that are not in ``table`` (callers raise KeyError on
        a non-empty ``missing`` *before* replacing their table).
    """
    rem = sort_unique(remove)
    both, inv, counts = sort_unique(torch.cat([table, rem], dim=0), return_inverse=True, return_counts=True)
    n_t = table.shape[0]
    from_rem = inv[n_t:]
    missing = rem[counts[from_rem] == 1]
    keep = torch.ones(both.shape[0], dtype=torch.bool, device=table.device)
    keep[from_rem] = False
    return both[keep].contiguous(), missing


def widen(codes: torch.Tensor, words: int) -> torch.Tensor:
    """Zero-extend rows to ``words`` words (prepends most-significant zeros)."""
    n, w0 = codes.shape
    if words == w0:
        return codes
    if words < w0:
        raise ValueError("cannot narrow %d words to %d" % (w0, words))
    out = torch.zeros((n, words), dtype=torch.int32, device=codes.device)
    out[:, words - w0:] = codes
    return out


def group_rows(inverse: torch.Tensor, num_codes: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR ``code -> descriptor rows``: (offsets int64[U+1], rows int64[N]) with the
    rows of each code in ascending order (stable sort of the inverse map)."""
    order = torch.sort(inverse, stable=True).indices
    counts = torch.bincount(inverse, minlength=num_codes)
    off = torch.zeros(num_codes + 1, dtype=torch.int64, device=inverse.device)
    torch.cumsum(counts, 0, out=off[1:])
    return off, order
