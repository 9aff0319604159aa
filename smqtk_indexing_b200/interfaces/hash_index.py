"""HashIndex interface mirror (reference: smqtk_indexing/interfaces/hash_index.py:10-182)."""
import abc
from typing import Iterable, Sequence, Tuple

import numpy as np

from smqtk_core import Configurable, Pluggable

from ._validation import call_if_not_empty


class HashIndex(Configurable, Pluggable):
    """
    Nearest-neighbour index over *unique* hash codes (bit vectors) under the
    Hamming metric.  ``nn`` never returns the same code twice.
    """

    def __len__(self) -> int:
        return self.count()

    @staticmethod
    def _empty_iterable_exception() -> BaseException:
        return ValueError("No hash vectors in provided iterable.")

    def build_index(self, hashes: Iterable[np.ndarray]) -> None:
        """Replace the index content. :raises ValueError: empty iterable."""
        call_if_not_empty(hashes, self._build_index, self._empty_iterable_exception())

    def update_index(self, hashes: Iterable[np.ndarray]) -> None:
        """Add codes to the index. :raises ValueError: empty iterable."""
        call_if_not_empty(hashes, self._update_index, self._empty_iterable_exception())

    def remove_from_index(self, hashes: Iterable[np.ndarray]) -> None:
        """Remove codes. :raises ValueError: empty iterable.
        :raises KeyError: a code is not indexed (index left unmodified)."""
        call_if_not_empty(hashes, self._remove_from_index, self._empty_iterable_exception())

    def nn(self, h: np.ndarray, n: int = 1) -> Tuple[np.ndarray, Sequence[float]]:
        """``n`` nearest codes to ``h`` and their normalised Hamming distances
        (``popcount / len(h)``). :raises ValueError: the index is empty."""
        if not self.count():
            raise ValueError("No index currently set to query from!")
        return self._nn(h, n)

    @abc.abstractmethod
    def count(self) -> int: ...

    @abc.abstractmethod
    def _build_index(self, hashes: Iterable[np.ndarray]) -> None: ...

    @abc.abstractmethod
    def _update_index(self, hashes: Iterable[np.ndarray]) -> None: ...

    @abc.abstractmethod
    def _remove_from_index(self, hashes: Iterable[np.ndarray]) -> None: ...

    @abc.abstractmethod
    def _nn(self, h: np.ndarray, n: int = 1) -> Tuple[np.ndarray, Tuple[float, ...]]: ...
