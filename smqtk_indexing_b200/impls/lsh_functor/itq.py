"""
B200 ``ItqFunctor``: Iterative-Quantization hash functor whose hashing (and the
N-scaled parts of fitting) run on the GPU through ``libsmqtk_b200``.

Drop-in for ``smqtk_indexing.impls.lsh_functor.itq.ItqFunctor`` (reference:
smqtk_indexing/impls/lsh_functor/itq.py:32-408): same constructor, config keys,
model attributes (``mean_vec``, ``rotation``), ``.npy`` model caches, methods and
exceptions.  ``get_hash`` = ``((x / norm(x)) - mean_vec) . rotation >= 0`` is
computed by ``sb_itq_hash`` in FP32; bits can differ from the float64 reference
only where ``|z| <= 1e-5 * |x - mean| * |r_j|``.
"""
import logging
from collections.abc import Sequence
from copy import deepcopy
from io import BytesIO
from typing import Any, Dict, Iterable, Optional, Type, TypeVar, Union

import numpy as np

from smqtk_core.configuration import from_config_dict, make_default_config, to_config_dict
from smqtk_core.dict import merge_dict
from smqtk_dataprovider import DataElement
from smqtk_descriptors import DescriptorElement
from smqtk_descriptors.utils import parallel_map

from smqtk_indexing_b200.interfaces import LshFunctor
from smqtk_indexing_b200.utils import bits as bitutil

LOG = logging.getLogger(__name__)
T_IF = TypeVar("T_IF", bound="ItqFunctor")


class ItqFunctor(LshFunctor):
    """ITQ hash functor (Gong & Lazebnik, CVPR 2011); bit 0 is the MSB."""

    @classmethod
    def is_usable(cls) -> bool:
        # library built and a CUDA device visible (no CPU implementation exists to fall back to)
        from smqtk_indexing_b200 import _lib
        return _lib.usable()

    @classmethod
    def get_default_config(cls) -> Dict[str, Any]:
        default = super(ItqFunctor, cls).get_default_config()
        data_element_default_config = make_default_config(DataElement.get_impls())
        default['mean_vec_cache'] = data_element_default_config
        default['rotation_cache'] = deepcopy(data_element_default_config)
        return default

    @classmethod
    def from_config(cls: Type[T_IF], config_dict: Dict, merge_default: bool = True) -> T_IF:
        if merge_default:
            config_dict = merge_dict(cls.get_default_config(), config_dict)
        data_element_impls = DataElement.get_impls()
        for key in ('mean_vec_cache', 'rotation_cache'):
            elem = None
            if config_dict[key] and config_dict[key]['type']:
                elem = from_config_dict(config_dict[key], data_element_impls)
            config_dict[key] = elem
        return super(ItqFunctor, cls).from_config(config_dict, False)

    def __init__(
        self,
        mean_vec_cache: Optional[DataElement] = None,
        rotation_cache: Optional[DataElement] = None,
        bit_length: int = 8,
        itq_iterations: int = 50,
        normalize: Optional[Union[int, float, str]] = None,
        random_seed: Optional[int] = None
    ):
        super(ItqFunctor, self).__init__()
        self.mean_vec_cache_elem = mean_vec_cache
        self.rotation_cache_elem = rotation_cache
        self.bit_length = bit_length
        self.itq_iterations = itq_iterations
        self.normalize = normalize
        self.random_seed = random_seed

        # Validate the normalisation order the way the reference does (itq.py:162-164),
        # and additionally that the device kernel supports it.
        if normalize is not None:
            self._norm_vector(np.random.rand(8))
            from smqtk_indexing_b200.device import norm_spec
            norm_spec(normalize)

        self._mean_vec: Optional[np.ndarray] = None
        self._rotation: Optional[np.ndarray] = None
        self._dev_model = None  # (device, mean f32[D], rotation f32[D, b])
        self.load_model()

    # -- model attributes; assigning either drops the device copy -----------------
    @property
    def mean_vec(self) -> Optional[np.ndarray]:
        return self._mean_vec

    @mean_vec.setter
    def mean_vec(self, v: Optional[np.ndarray]) -> None:
        self._mean_vec = v
        self._dev_model = None

    @property
    def rotation(self) -> Optional[np.ndarray]:
        return self._rotation

    @rotation.setter
    def rotation(self, v: Optional[np.ndarray]) -> None:
        self._rotation = v
        self._dev_model = None

    def _norm_vector(self, v: np.ndarray) -> np.ndarray:
        """Host mirror of the kernel's row normalisation (reference itq.py:172-191);
        used for argument validation and by ``fit``'s host-side bookkeeping."""
        if self.normalize is not None:
            n = np.linalg.norm(v, self.normalize, v.ndim - 1, keepdims=True)
            n[n == 0.] = 1.
            return v / n
        return v

    def get_config(self) -> Dict[str, Any]:
        c = merge_dict(self.get_default_config(), {
            "bit_length": self.bit_length,
            "itq_iterations": self.itq_iterations,
            "normalize": self.normalize,
            "random_seed": self.random_seed,
        })
        if self.mean_vec_cache_elem:
            c['mean_vec_cache'] = to_config_dict(self.mean_vec_cache_elem)
        if self.rotation_cache_elem:
            c['rotation_cache'] = to_config_dict(self.rotation_cache_elem)
        return c

    def has_model(self) -> bool:
        return (self.mean_vec is not None) and (self.rotation is not None)

    def load_model(self) -> None:
        if (self.mean_vec_cache_elem
                and not self.mean_vec_cache_elem.is_empty()
                and self.rotation_cache_elem
                and not self.rotation_cache_elem.is_empty()):
            self.mean_vec = np.load(BytesIO(self.mean_vec_cache_elem.get_bytes()))
            self.rotation = np.load(BytesIO(self.rotation_cache_elem.get_bytes()))

    def save_model(self) -> None:
        if (self.mean_vec_cache_elem and self.rotation_cache_elem and
                self.mean_vec_cache_elem.writable() and
                self.rotation_cache_elem.writable() and
                self.mean_vec is not None and self.rotation is not None):
            for elem, arr in ((self.mean_vec_cache_elem, self.mean_vec),
                              (self.rotation_cache_elem, self.rotation)):
                b = BytesIO()
                np.save(b, arr)
                elem.set_bytes(b.getvalue())

    # ------------------------------------------------------------------ hashing
    def device_model(self, dev=None):
        """(mean f32[D], rotation f32[D, b]) tensors on ``dev`` (cached)."""
        if self.mean_vec is None:
            raise Exception("Can't compute hash code: mean vector is none.")
        elif self.rotation is None:
            raise Exception("Can't compute hash code: rotation matrix is none.")
        import torch
        from smqtk_indexing_b200 import device
        d = device.device(dev)
        if self._dev_model is None or self._dev_model[0] != d:
            mean = torch.from_numpy(np.ascontiguousarray(np.real(self.mean_vec), dtype=np.float32).ravel()).to(d)
            rot = np.real(np.asarray(self.rotation))
            rot = rot.reshape(rot.shape[0], -1)
            rotation = torch.from_numpy(np.ascontiguousarray(rot, dtype=np.float32)).to(d)
            # rotation pre-split for the tensor-core kernel (None: shape not supported)
            self._dev_model = (d, mean, rotation, device.itq_rotation_image(rotation))
        return self._dev_model[1], self._dev_model[2]

    def get_hash_packed(self, descriptors, variant: int = 0):
        """Batch hashing on the device: ``descriptors`` is a float32 CUDA tensor
        (or array-like, uploaded) of shape [n, D]; returns packed codes
        int32[n, W] (uint32 bit patterns, layout in utils/bits.py)."""
        import torch
        from smqtk_indexing_b200 import device
        if not isinstance(descriptors, torch.Tensor):
            descriptors = torch.from_numpy(np.ascontiguousarray(descriptors, dtype=np.float32))
        x = descriptors
        if not x.is_cuda:
            x = x.to(device.device())
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.stride(-1) != 1:
            x = x.contiguous()
        mean, rotation = self.device_model(x.device)
        return device.itq_hash(x, mean, rotation, normalize=self.normalize, variant=variant,
                               r_image=self._dev_model[3])

    def get_hash(self, descriptor: np.ndarray) -> np.ndarray:
        """Hash one descriptor ``[D]`` (or a matrix ``[n, D]``) -> ``bool[b]``
        (``bool[n, b]``), index 0 = most significant bit."""
        if self.mean_vec is None:
            raise Exception("Can't compute hash code: mean vector is none.")
        elif self.rotation is None:
            raise Exception("Can't compute hash code: rotation matrix is none.")
        from smqtk_indexing_b200 import device
        d = np.asarray(descriptor)
        codes = device.codes_to_host(self.get_hash_packed(np.atleast_2d(d)))
        b = np.asarray(self.rotation).reshape(np.asarray(self.rotation).shape[0], -1).shape[1]
        out = bitutil.unpack_bits(codes, b)
        return out[0] if d.ndim == 1 else out

    # ------------------------------------------------------------------ fitting
    def fit(self, descriptors: Iterable[DescriptorElement], use_multiprocessing: bool = True) -> np.ndarray:
        """Fit the ITQ model (reference itq.py:291-387): normalise, centre,
        PCA to ``bit_length`` dims, ``itq_iterations`` rotation refinements.
        The N-scaled contractions run on the GPU (``fit.py``); the D x D
        eigen-decomposition and b x b SVDs use host LAPACK like the reference.

        :raises RuntimeError: a model is already loaded.
        :raises ValueError: descriptors have fewer features than ``bit_length``.
        :return: ``bool[N, bit_length]`` codes of the training descriptors.
        """
        if self.has_model():
            raise RuntimeError("Model components have already been loaded.")
        if not isinstance(descriptors, Sequence):
            descriptors = list(descriptors)
        if len(descriptors[0].vector()) < self.bit_length:
            raise ValueError("Input descriptors have fewer features than "
                             "requested bit encoding. Hash codes will be "
                             "smaller than requested due to PCA decomposition "
                             "result being bound by number of features.")
        x = np.asarray(list(
            parallel_map(lambda d_: d_.vector(), descriptors,
                         use_multiprocessing=use_multiprocessing)
        ))
        return self.fit_matrix(x)

    def fit_matrix(self, x, want_codes: bool = True, group=None):
        """``fit`` on an assembled ``[N, D]`` matrix (numpy or CUDA tensor).  ``want_codes=False``
        skips the N x b bool matrix the reference's ``fit`` returns (returns ``None``).
        ``group``: a ``torch.distributed`` group whose ranks each pass their ROW SHARD of the training
        matrix (collective): the partial sums are all-reduced, every rank gets the same model."""
        if self.has_model():
            raise RuntimeError("Model components have already been loaded.")
        from smqtk_indexing_b200 import fit as fitops
        if x.shape[1] < self.bit_length:
            raise ValueError("Input descriptors have fewer features than "
                             "requested bit encoding.")
        codes, mean_vec, rotation = fitops.itq_fit(
            x, self.bit_length, self.itq_iterations, self.normalize, self.random_seed, want_codes=want_codes,
            group=group)
        self.mean_vec = mean_vec
        self.rotation = rotation
        self.save_model()
        return codes
