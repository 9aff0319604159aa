"""
Device-side bookkeeping for tables of packed hash codes: lexicographic
sort / unique / membership on ``int32[rows, W]`` tensors that carry uint32 bit
patterns (word 0 most significant).

These are index-maintenance steps (build / update / remove), not the query hot
path; they are expressed with torch's device sort (plumbing) so that the table
never leaves HBM.  Unsigned order is obtained by flipping the sign bit of every
word before the signed lexicographic sort.
"""
from typing import Tuple

import torch

_SIGN = -2 ** 31


def _flip(t: torch.Tensor) -> torch.Tensor:
    return torch.bitwise_xor(t, torch.tensor(_SIGN, dtype=torch.int32, device=t.device))


def sort_unique(codes: torch.Tensor, return_inverse: bool = False, return_counts: bool = False):
    """Rows of ``codes`` sorted ascending by integer value, duplicates removed.

    :return: table int32[U, W] (+ inverse int64[rows] -> table row, + counts int64[U])
    """
    if codes.dim() != 2 or codes.dtype != torch.int32:
        raise ValueError("codes must be int32[rows, W]")
    if codes.shape[0] == 0:
        out = [codes.clone()]
        if return_inverse:
            out.append(torch.empty(0, dtype=torch.int64, device=codes.device))
        if return_counts:
            out.append(torch.empty(0, dtype=torch.int64, device=codes.device))
        return out[0] if len(out) == 1 else tuple(out)
    res = torch.unique(_flip(codes), dim=0, sorted=True, return_inverse=return_inverse,
                       return_counts=return_counts)
    if isinstance(res, tuple):
        return (_flip(res[0]).contiguous(),) + tuple(res[1:])
    return _flip(res).contiguous()


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
