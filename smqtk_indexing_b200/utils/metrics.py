"""
Distance functions of the LSH path with the reference's signatures (reference:
smqtk_indexing/utils/metrics.py), evaluated on the GPU by the same kernels the
index uses (``sb_rerank`` / ``sb_hamming_topk``) -- there is no numpy
implementation in this package; the numpy restatement lives in ``oracle/`` and
is test infrastructure only.

Inputs are rounded to float32 on upload; results are float64 (see DESIGN.md for
the accuracy contract: 1e-5 relative + the cosine conditioning floor).
"""
from typing import Union

import numpy as np


def _pairwise(i: np.ndarray, j: np.ndarray, metric: str) -> Union[float, np.ndarray]:
    import torch
    from smqtk_indexing_b200 import device
    i = np.asarray(i)
    j = np.asarray(j)
    if i.ndim == 2 and j.ndim == 1:
        i, j = j, i                      # symmetric metrics: keep the 1-D one as the query
    if i.ndim == 2 and j.ndim == 2:
        if i.shape != j.shape:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (i.shape, j.shape))
        return np.array([_pairwise(a, b, metric) for a, b in zip(i, j)])
    if i.shape[-1] != j.shape[-1]:
        raise ValueError("operands could not be broadcast together with shapes %s %s" % (i.shape, j.shape))
    dev = device.device()
    q = torch.from_numpy(np.ascontiguousarray(i, dtype=np.float32)[None, :]).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(j), dtype=np.float32)).to(dev)
    m = c.shape[0]
    out = device.rerank(c, q, torch.arange(m, dtype=torch.int64, device=dev),
                        torch.tensor([0, m], dtype=torch.int64, device=dev), metric).cpu().numpy()
    return float(out[0]) if j.ndim == 1 else out


def histogram_intersection_distance(a: np.ndarray, b: np.ndarray) -> Union[float, np.ndarray]:
    """1 - sum(min(a, b)); 1-D/2-D broadcasting as in the reference (metrics.py:7-46)."""
    return _pairwise(a, b, "hik")


def histogram_intersection_distance_fast(i: np.ndarray, j: np.ndarray) -> float:
    """1-D form (metrics.py:49-70)."""
    return _pairwise(i, j, "hik")


def euclidean_distance(i: np.ndarray, j: np.ndarray) -> Union[float, np.ndarray]:
    """sqrt(sum((i - j)^2)) (metrics.py:73-86)."""
    return _pairwise(i, j, "euclidean")


def cosine_distance(i: np.ndarray, j: np.ndarray, pos_vectors: bool = True) -> Union[float, np.ndarray]:
    """Angular distance (1 + pos_vectors) * acos(cos_sim) / pi (metrics.py:120-137)."""
    d = _pairwise(i, j, "cosine")
    return d if pos_vectors else d * 0.5


def cosine_similarity(i: np.ndarray, j: np.ndarray) -> Union[float, np.ndarray]:
    """cos of the angular distance (metrics.py:89-117)."""
    return np.cos(np.asarray(cosine_distance(i, j)) * (np.pi / 2.0))


def hamming_distance(i: int, j: int) -> int:
    """Number of differing bits of two Python ints (metrics.py:140-155)."""
    import torch  # noqa: F401
    from smqtk_indexing_b200 import device
    from smqtk_indexing_b200.utils import bits as bitutil
    w = bitutil.words_for_bits(max(int(i).bit_length(), int(j).bit_length(), 1))
    db = device.codes_to_device(bitutil.ints_to_words([i], w))
    q = device.codes_to_device(bitutil.ints_to_words([j], w))
    d, _ = device.hamming_topk(db, q, 1)
    return int(d[0, 0].item())
