"""
ITQ model fitting (reference: smqtk_indexing/impls/lsh_functor/itq.py:291-387 and
``_find_itq_rotation`` :239-289).  The row-count-scaled contractions -- row
norms, column means, covariance, PCA projection and the two products of every
ITQ iteration -- run on the GPU in FP64 (``sb_fit_*`` kernels).  The D x D
eigen-decomposition and the b x b SVDs are O(D^3) / O(b^3) dense factorizations
with no data parallelism over rows; they use host LAPACK exactly like the
reference, which also keeps eigenvector / singular-vector sign conventions
identical to it.

Algorithmic fidelity: same steps, same seeded ``randn`` + SVD initial rotation,
and the reference's update ``R = Vh . U^T`` (numpy's ``svd`` returns V^H;
itq.py:276-277) -- not the textbook Procrustes ``V . U^T``.
"""
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .device import _ptr, _stream, device as _device, norm_spec, require_cuda
from .utils.bits import unpack_bits, words_for_bits

_KIND = {torch.float32: 0, torch.float64: 1}


def _workspace(n: int, ma: int, mb: int, dev) -> torch.Tensor:
    nbytes = _lib.load().sb_fit_workspace_bytes(n, ma, mb)
    return torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)


def row_div(x: torch.Tensor, normalize) -> Optional[torch.Tensor]:
    kind, p = norm_spec(normalize)
    if kind == _lib.NORM_NONE:
        return None
    n, d = x.shape
    out = torch.empty((n,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().sb_fit_row_div(_ptr(x), _KIND[x.dtype], n, d, x.stride(0), kind, p, _ptr(out), _stream()))
    return out


def col_mean(x: torch.Tensor, div: Optional[torch.Tensor]) -> torch.Tensor:
    n, d = x.shape
    ws = _workspace(n, d, d, x.device)
    out = torch.empty((d,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().sb_fit_col_mean(_ptr(x), _KIND[x.dtype], n, d, x.stride(0), _ptr(div), _ptr(out),
                                           _ptr(ws), ws.numel(), _stream()))
    return out


def gram(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0, a_div=None, a_mean=None, b_div=None, b_mean=None,
         a_bits: int = 0) -> torch.Tensor:
    """scale * opA^T opB.  ``a`` may be packed sign bits (int32[n, W], ``a_bits`` > 0)."""
    n = a.shape[0]
    a_kind = 2 if a_bits else _KIND[a.dtype]
    ma = a_bits if a_bits else a.shape[1]
    mb = b.shape[1]
    ws = _workspace(n, ma, mb, b.device)
    out = torch.empty((ma, mb), dtype=torch.float64, device=b.device)
    _lib.check(_lib.load().sb_fit_gram(
        _ptr(a), a_kind, a.stride(0), ma, _ptr(a_div), _ptr(a_mean), a_bits,
        _ptr(b), _KIND[b.dtype], b.stride(0), mb, _ptr(b_div), _ptr(b_mean), 0,
        n, scale, _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


def centred_bound(x: torch.Tensor, mean: Optional[torch.Tensor], div: Optional[torch.Tensor]) -> float:
    """An upper bound on ``|x / div - mean|`` (column min / max reductions, no N x D temporary);
    tight when there is no row normalisation."""
    lo, hi = torch.aminmax(x, dim=0)
    lo, hi = lo.double(), hi.double()
    if div is None:
        m = mean.double() if mean is not None else torch.zeros_like(lo)
        b = float(torch.maximum((hi - m).abs(), (lo - m).abs()).max())
    else:
        b = float(torch.maximum(lo.abs(), hi.abs()).max()) / max(float(div.min()), 1e-30)
        if mean is not None:
            b += float(mean.abs().max())
    return min(max(b * (1.0 + 1e-6), 1e-30), 1e29)


def gram_bits_tc(codes: torch.Tensor, bits: int, x: torch.Tensor, mean32: Optional[torch.Tensor] = None,
                 div32: Optional[torch.Tensor] = None, bound: Optional[float] = None) -> torch.Tensor:
    """Tensor-core ``UX^T . (x / div - mean)`` -> float64[bits, D] (``sb_fit_gram_bits_tc``).
    ``bound``: an upper bound on ``|x / div - mean|`` (computed when not given)."""
    n, d = x.shape
    if bound is None:
        bound = centred_bound(x, mean32, div32)
    lib = _lib.load()
    nbytes = lib.sb_fit_gram_bits_tc_workspace_bytes(n, d, bits)
    ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=x.device)
    out = torch.empty((bits, d), dtype=torch.float64, device=x.device)
    _lib.check(lib.sb_fit_gram_bits_tc(_ptr(codes), codes.shape[1], bits, _ptr(x), n, d, x.stride(0), _ptr(mean32),
                                       _ptr(div32), float(bound), _ptr(out), _ptr(ws), nbytes, _stream()))
    return out


def gram_bits_tc_supported(x: torch.Tensor, bits: int) -> bool:
    return (x.dtype == torch.float32 and x.data_ptr() % 16 == 0
            and bool(_lib.load().sb_fit_gram_bits_tc_supported(x.shape[0], x.shape[1], x.stride(0), bits)))


def project(a: torch.Tensor, bm: torch.Tensor, a_div=None, a_mean=None, want_values: bool = True,
            want_codes: bool = False):
    """opA . bm (f64[K, M]) -> values f64[n, M] and/or sign-bit codes int32[n, W]."""
    n, k = a.shape
    m = bm.shape[1]
    vals = torch.empty((n, m), dtype=torch.float64, device=a.device) if want_values else None
    w = words_for_bits(m) if want_codes else 0
    codes = torch.empty((n, w), dtype=torch.int32, device=a.device) if want_codes else None
    _lib.check(_lib.load().sb_fit_project(_ptr(a), _KIND[a.dtype], a.stride(0), k, _ptr(a_div), _ptr(a_mean),
                                          _ptr(bm), m, n, _ptr(vals), _ptr(codes), w, _stream()))
    return vals, codes


def _eig_descending(c: np.ndarray, bit_length: int) -> np.ndarray:
    """Top ``bit_length`` eigenvectors as columns, ordered like itq.py:356-376
    (general ``eig`` + stable descending sort).  For a rank-deficient covariance
    ``eig`` can return complex pairs with ~1e-17 imaginary parts (the reference
    then silently carries a complex rotation); the symmetric solver is used in
    that case so the model stays real."""
    l, pc = np.linalg.eig(c)
    if np.iscomplexobj(l) or np.iscomplexobj(pc):
        l, pc = np.linalg.eigh(c)
    order = sorted(range(len(l)), key=lambda i: l[i], reverse=True)
    return np.array([pc[:, i] for i in order[:bit_length]]).transpose()


def _allreduce_sum(t: torch.Tensor, group) -> torch.Tensor:
    """SUM over the ranks that hold row shards (the fit is data-parallel over rows; only the
    [D], [D, D] and [b, D] / [b, b] partial sums cross NVLink).  No-op for a single process."""
    if group is not None:
        import torch.distributed as dist
        if dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def itq_rotation(v: torch.Tensor, n_iter: int, random_seed: Optional[int], group=None) -> Tuple[torch.Tensor, np.ndarray]:
    """``_find_itq_rotation`` (itq.py:239-289) on a device-resident ``v`` f64[n, b] (this rank's rows).

    :return: (codes int32[n, W] of the final rotation, rotation f64[b, b] on host)
    """
    bit = v.shape[1]
    if random_seed is not None:
        np.random.seed(random_seed)
    r = np.random.randn(bit, bit)
    u11, _, _ = np.linalg.svd(r)
    r = u11[:, :bit]
    for _ in range(n_iter):
        r_dev = torch.from_numpy(np.ascontiguousarray(r)).to(v.device)
        _, ux = project(v, r_dev, want_values=False, want_codes=True)      # sign(v . r)
        c = _allreduce_sum(gram(ux, v, a_bits=bit), group).cpu().numpy()    # ux^T . v
        ub, _, ua = np.linalg.svd(c)
        r = np.dot(ua, ub.transpose())
    r_dev = torch.from_numpy(np.ascontiguousarray(r)).to(v.device)
    _, codes = project(v, r_dev, want_values=False, want_codes=True)
    return codes, r


#: Above this size the PCA-projected training matrix V (float64[N, b]) is not materialised:
#: every ITQ iteration works from X directly (see ``itq_rotation_streaming``).
STREAMING_V_BYTES = 4 << 30
#: Below this many rows the FP64 Gram kernel is used even when the tensor-core one applies
TC_GRAM_MIN_ROWS = 1 << 16


def itq_rotation_streaming(xt: torch.Tensor, div, mean, pc_top: np.ndarray, n_iter: int,
                           random_seed: Optional[int], normalize=None,
                           tensor_cores: Optional[bool] = None, group=None) -> Tuple[torch.Tensor, np.ndarray]:
    """``_find_itq_rotation`` (itq.py:239-289) without the N x b matrix ``v = x . pc_top``:
    with P = pc_top, ``z = v.r = x.(P r)`` and ``c = ux^T v = (ux^T x) P``, so an iteration is one
    projection of X by the D x b matrix ``P r`` (sign bits only) and one b x D Gram
    ``ux^T x`` -- same arithmetic up to FP64 re-association, 4 N D b flop per iteration instead
    of 4 N b^2, and no 8 N b bytes of HBM (102 GB for 50M x 256 bits)."""
    from . import device
    bit = pc_top.shape[1]
    if random_seed is not None:
        np.random.seed(random_seed)
    r = np.random.randn(bit, bit)
    u11, _, _ = np.linalg.svd(r)
    r = u11[:, :bit]
    # float32 training data of an aligned shape: the sign step IS ItqFunctor.get_hash with the
    # rotation P r, so it runs on the production tensor-core hash kernel (3xTF32; bits can differ
    # from the FP64 projection only for |z| at rounding level -- the same bits queries will get)
    mean32 = mean.to(torch.float32) if (xt.dtype == torch.float32 and tensor_cores is not False) else None
    use_tc = mean32 is not None and device.itq_tc_supported(xt, bit)
    # ... and the +-1 Gram ux^T . x runs on the tensor cores too (fixed-point hi/lo split: exact accumulation)
    use_tc_gram = mean32 is not None and gram_bits_tc_supported(xt, bit) and xt.shape[0] >= TC_GRAM_MIN_ROWS
    div32 = div.to(torch.float32) if (use_tc_gram and div is not None) else None
    bound = centred_bound(xt, mean32, div32) if use_tc_gram else None

    def sign_codes(rot: np.ndarray) -> torch.Tensor:
        pr = np.ascontiguousarray(pc_top @ rot)
        if use_tc:
            return device.itq_hash(xt, mean32, torch.from_numpy(pr.astype(np.float32)).to(xt.device),
                                   normalize=normalize, variant=2)
        prd = torch.from_numpy(pr).to(xt.device)
        return project(xt, prd, a_div=div, a_mean=mean, want_values=False, want_codes=True)[1]

    for _ in range(n_iter):
        ux = sign_codes(r)                                                                  # sign(x . P r)
        if use_tc_gram:
            g = gram_bits_tc(ux, bit, xt, mean32, div32, bound)                                     # ux^T . x   [b, D]
        else:
            g = gram(ux, xt, a_bits=bit, b_div=div, b_mean=mean)
        g = _allreduce_sum(g, group).cpu().numpy()
        ub, _, ua = np.linalg.svd(g @ pc_top)
        r = np.dot(ua, ub.transpose())
    return sign_codes(r), r


def itq_fit(x, bit_length: int, itq_iterations: int = 50, normalize=None,
            random_seed: Optional[int] = None, dev=None, streaming: Optional[bool] = None,
            tensor_cores: Optional[bool] = None, want_codes: bool = True, group=None):
    """Fit on ``x`` ([N, D] numpy array or CUDA tensor, float32 or float64).

    :param group: a ``torch.distributed`` process group whose ranks each pass THEIR row shard of the
        training matrix (collective call): column sums, the covariance and every iteration's
        ``ux^T v`` are summed over the ranks with all-reduces of [D], [D, D] and [b, D] values
        (itq.py:338-383 is a sum over rows throughout); every rank ends with the same model and the
        training codes of its own rows.

    :param streaming: never materialise ``v`` (default: automatic, when it would exceed
        ``STREAMING_V_BYTES``).
    :param tensor_cores: streaming fit only. ``False`` keeps every N-scaled product in FP64 (the
        parity path: the rotation agrees with the numpy reference to 1e-8); ``None`` / ``True`` take
        the sign bits from the 3xTF32 hash kernel and, from 65536 rows, ``ux^T x`` from the
        fixed-point tensor-core Gram (the Gram is good to 1e-8, but a few rounding-level sign flips
        per iteration make the rotation differ at the 1e-3 level -- an equally good ITQ model).
    :param want_codes: ``False`` skips the device-to-host copy and the N x b bool expansion of the
        training codes (the reference returns them; 12.8 GB of host memory for 50M x 256 bits) and
        returns ``None`` in their place.
    :return: (codes bool[N, b], mean_vec [D] in x's dtype, rotation float64[D, b])
    """
    require_cuda()
    if isinstance(x, torch.Tensor):
        xt = x if x.is_cuda else x.to(_device(dev))
        out_dtype = np.float32 if xt.dtype == torch.float32 else np.float64
    else:
        xa = np.asarray(x)
        if xa.dtype not in (np.float32, np.float64):
            xa = xa.astype(np.float64)
        out_dtype = xa.dtype
        xt = torch.from_numpy(np.ascontiguousarray(xa)).to(_device(dev))
    if xt.dtype not in _KIND:
        xt = xt.to(torch.float64)
    if xt.stride(1) != 1:
        xt = xt.contiguous()
    n, d = xt.shape
    if d < bit_length:
        raise ValueError("Input descriptors have fewer features than requested bit encoding.")
    with torch.cuda.device(xt.device):
        div = row_div(xt, normalize)
        n_total = n
        if group is not None:
            cnt = _allreduce_sum(torch.tensor([n], dtype=torch.float64, device=xt.device), group)
            n_total = int(cnt.item())
            mean = _allreduce_sum(col_mean(xt, div) * float(n), group) / float(n_total)
        else:
            mean = col_mean(xt, div)
        cov = _allreduce_sum(gram(xt, xt, scale=1.0, a_div=div, a_mean=mean, b_div=div, b_mean=mean), group) \
            if group is not None else gram(xt, xt, scale=1.0 / max(n - 1, 1), a_div=div, a_mean=mean, b_div=div, b_mean=mean)
        if group is not None:
            cov = cov / float(max(n_total - 1, 1))
        pc_top = _eig_descending(np.atleast_2d(cov.cpu().numpy()), bit_length)
        pc_dev = torch.from_numpy(np.ascontiguousarray(pc_top)).to(xt.device)
        if streaming is None:
            streaming = n * bit_length * 8 > STREAMING_V_BYTES
        if group is not None:                                       # every rank must take the same branch
            flag = _allreduce_sum(torch.tensor([1.0 if streaming else 0.0], dtype=torch.float64, device=xt.device), group)
            streaming = bool(flag.item() > 0)
        if streaming:
            codes, r = itq_rotation_streaming(xt, div, mean, pc_top, itq_iterations, random_seed, normalize,
                                              tensor_cores, group)
        else:
            v, _ = project(xt, pc_dev, a_div=div, a_mean=mean)
            codes, r = itq_rotation(v, itq_iterations, random_seed, group)
        codes_host = codes.cpu().numpy().view(np.uint32) if want_codes else None
    mean_vec = mean.cpu().numpy().astype(out_dtype)
    return (unpack_bits(codes_host, bit_length) if want_codes else None), mean_vec, np.dot(pc_top, r)
