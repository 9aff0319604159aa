"""
B200 ``SimpleRPFunctor``: random-projection LSH baseline through the same hash kernels as
``ItqFunctor``.

Mirror of smqtk_indexing/impls/lsh_functor/simple_rp.py:15-127: ``fit`` draws
``rps = randn(dim, bit_length)`` from ``numpy.random.seed(random_seed)``; ``get_hash`` is
``norm(v - mean_vec) . rps >= 0``.  Two facts about the reference shape this class:

* the reference never assigns ``mean_vec`` (simple_rp.py:40, :52), so its ``fit`` /
  ``get_hash`` raise ``TypeError`` and ``has_model()`` is never true.  Here ``fit`` sets
  ``mean_vec`` to the column mean of the training descriptors -- the evident intent
  (``has_model`` tests it, ``_norm_vector`` subtracts it);
* normalisation is applied AFTER centring (simple_rp.py:52-59) and divides each row by a
  positive scalar, so it cannot change the sign of a projection (a zero row projects to 0 ->
  bit set, with or without it): the bits are ``(v - mean_vec) . rps >= 0`` for every
  ``normalize`` setting, which is exactly ``sb_itq_hash`` without row normalisation.
"""
import logging
from typing import Any, Dict, Iterable, Optional, Sequence, Union

import numpy as np

from smqtk_descriptors import DescriptorElement
from smqtk_descriptors.utils import parallel_map

from smqtk_indexing_b200.interfaces import LshFunctor
from smqtk_indexing_b200.utils import bits as bitutil

LOG = logging.getLogger(__name__)


class SimpleRPFunctor(LshFunctor):
    """Sign of random projections of mean-centred descriptors (baseline functor)."""

    @classmethod
    def is_usable(cls) -> bool:
        # library built and a CUDA device visible (no CPU implementation exists to fall back to)
        from smqtk_indexing_b200 import _lib
        return _lib.usable()

    def __init__(self, bit_length: int = 8, normalize: Optional[Union[int, float, str]] = None,
                 random_seed: Optional[int] = None):
        super(SimpleRPFunctor, self).__init__()
        self.bit_length = bit_length
        self.normalize = normalize
        self.random_seed = random_seed
        self.rps: Optional[np.ndarray] = None
        self.mean_vec: Optional[np.ndarray] = None
        self._dev_model = None

    def get_config(self) -> Dict[str, Any]:
        return {
            "bit_length": self.bit_length,
            "normalize": self.normalize,
            "random_seed": self.random_seed,
        }

    def has_model(self) -> bool:
        return self.mean_vec is not None and self.rps is not None

    # ``rotation`` alias: lets LSHNearestNeighborIndex treat this functor like ItqFunctor
    @property
    def rotation(self) -> Optional[np.ndarray]:
        return self.rps

    def _device_model(self, dev):
        import torch
        from smqtk_indexing_b200 import device
        d = device.device(dev)
        if self._dev_model is None or self._dev_model[0] != d:
            mean = torch.from_numpy(np.ascontiguousarray(self.mean_vec, dtype=np.float32)).to(d)
            rps = torch.from_numpy(np.ascontiguousarray(self.rps, dtype=np.float32)).to(d)
            self._dev_model = (d, mean, rps, device.itq_rotation_image(rps))
        return self._dev_model[1:]

    def get_hash_packed(self, descriptors, variant: int = 0):
        """Batch hashing on the device: float32 ``[n, D]`` (CUDA tensor or array-like) ->
        packed codes int32[n, W]."""
        import torch
        from smqtk_indexing_b200 import device
        if self.rps is None or self.mean_vec is None:
            raise RuntimeError("Random projection model not constructed. Call `fit` first!")
        x = descriptors
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if not x.is_cuda:
            x = x.to(device.device())
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.stride(-1) != 1:
            x = x.contiguous()
        mean, rps, image = self._device_model(x.device)
        return device.itq_hash(x, mean, rps, normalize=None, variant=variant, r_image=image)

    def get_hash(self, descriptor: np.ndarray) -> np.ndarray:
        if self.rps is None or self.mean_vec is None:
            raise RuntimeError("Random projection model not constructed. Call `fit` first!")
        from smqtk_indexing_b200 import device
        d = np.asarray(descriptor)
        codes = device.codes_to_host(self.get_hash_packed(np.atleast_2d(d)))
        out = bitutil.unpack_bits(codes, self.bit_length)
        return out[0] if d.ndim == 1 else out

    def fit(self, descriptors: Iterable[DescriptorElement], use_multiprocessing: bool = True) -> np.ndarray:
        """Draw the projections, set the centring vector, return the training codes
        (``bool[N, bit_length]``).  :raises RuntimeError: a model is already loaded."""
        if self.has_model():
            raise RuntimeError("Model components have already been loaded.")
        if not isinstance(descriptors, Sequence):
            descriptors = list(descriptors)
        x = np.asarray(list(parallel_map(lambda d_: d_.vector(), descriptors,
                                         use_multiprocessing=use_multiprocessing)))
        return self.fit_matrix(x)

    def fit_matrix(self, x: np.ndarray) -> np.ndarray:
        if self.has_model():
            raise RuntimeError("Model components have already been loaded.")
        x = np.asarray(x)
        n, dim = x.shape
        np.random.seed(self.random_seed)
        self.rps = np.random.randn(dim, self.bit_length)
        self.mean_vec = x.mean(axis=0)
        self._dev_model = None
        return self.get_hash(x)
