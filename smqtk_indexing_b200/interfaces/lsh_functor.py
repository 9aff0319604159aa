"""LshFunctor interface mirror (reference: smqtk_indexing/interfaces/lsh_functor.py:11-41)."""
import abc

import numpy as np

from smqtk_core import Configurable, Pluggable


class LshFunctor(Configurable, Pluggable):
    """Maps a descriptor vector to a locality-sensitive bit vector."""

    def __call__(self, descriptor: np.ndarray) -> np.ndarray:
        return self.get_hash(descriptor)

    @abc.abstractmethod
    def get_hash(self, descriptor: np.ndarray) -> np.ndarray:
        """
        :param descriptor: float vector ``[D]``.
        :return: ``bool[b]`` hash code, index 0 = most significant bit.
        """
