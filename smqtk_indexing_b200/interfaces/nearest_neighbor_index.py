"""NearestNeighborsIndex interface mirror
(reference: smqtk_indexing/interfaces/nearest_neighbor_index.py:13-184)."""
import abc
from typing import Hashable, Iterable, Tuple

from smqtk_core import Configurable, Pluggable
from smqtk_descriptors import DescriptorElement

from ._validation import call_if_not_empty


class NearestNeighborsIndex(Configurable, Pluggable):
    """Descriptor nearest-neighbour index; implementations are thread safe."""

    def __len__(self) -> int:
        return self.count()

    @staticmethod
    def _empty_iterable_exception() -> BaseException:
        return ValueError("No DescriptorElement instances in provided iterable.")

    def build_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        """Replace the index content. :raises ValueError: empty iterable."""
        call_if_not_empty(descriptors, self._build_index, self._empty_iterable_exception())

    def update_index(self, descriptors: Iterable[DescriptorElement]) -> None:
        """Add descriptors. :raises ValueError: empty iterable."""
        call_if_not_empty(descriptors, self._update_index, self._empty_iterable_exception())

    def remove_from_index(self, uids: Iterable[Hashable]) -> None:
        """Remove by uid. :raises ValueError: empty iterable.
        :raises KeyError: unknown uid (index left unmodified)."""
        call_if_not_empty(uids, self._remove_from_index, self._empty_iterable_exception())

    def nn(self, d: DescriptorElement, n: int = 1) -> Tuple[Tuple[DescriptorElement, ...], Tuple[float, ...]]:
        """``n`` nearest descriptors and their distances.
        :raises ValueError: ``d`` has no vector, or the index is empty."""
        if not d.has_vector():
            raise ValueError("Query descriptor did not have a vector set!")
        elif not self.count():
            raise ValueError("No index currently set to query from!")
        return self._nn(d, n)

    @abc.abstractmethod
    def count(self) -> int: ...

    @abc.abstractmethod
    def _build_index(self, descriptors: Iterable[DescriptorElement]) -> None: ...

    @abc.abstractmethod
    def _update_index(self, descriptors: Iterable[DescriptorElement]) -> None: ...

    @abc.abstractmethod
    def _remove_from_index(self, uids: Iterable[Hashable]) -> None: ...

    @abc.abstractmethod
    def _nn(self, d: DescriptorElement, n: int = 1) -> Tuple[Tuple[DescriptorElement, ...], Tuple[float, ...]]: ...
