"""
Minimal stand-in for ``smqtk-descriptors`` 0.18 (DescriptorElement /
DescriptorSet surface used by SMQTK-Indexing: lsh.py:305-306,316-321,412,450,
472,501,507-509; itq.py:333-336).  No hot-path arithmetic lives here.
"""
import abc
from typing import Any, Hashable, Iterable, Iterator, Optional

import numpy

from smqtk_core import Configurable, Pluggable


class DescriptorElement(Configurable, Pluggable):
    """A uniquely identified container of one descriptor vector."""

    def __init__(self, uuid: Hashable):
        super().__init__()
        self._uuid = uuid

    def __hash__(self) -> int:
        return hash(self.uuid())

    def __eq__(self, other: Any) -> bool:
        if isinstance(other, DescriptorElement):
            a, b = self.vector(), other.vector()
            if a is None or b is None:
                return a is None and b is None
            return bool(numpy.array_equal(a, b))
        return False

    def __ne__(self, other: Any) -> bool:
        return not (self == other)

    def __repr__(self) -> str:
        return "%s{uuid: %s}" % (type(self).__name__, self.uuid())

    def uuid(self) -> Hashable:
        return self._uuid

    @abc.abstractmethod
    def has_vector(self) -> bool: ...

    @abc.abstractmethod
    def vector(self) -> Optional[numpy.ndarray]: ...

    @abc.abstractmethod
    def set_vector(self, new_vec: Any) -> "DescriptorElement": ...


class DescriptorSet(Configurable, Pluggable):
    """uuid -> DescriptorElement collection."""

    def __len__(self) -> int:
        return self.count()

    def __iter__(self) -> Iterator[DescriptorElement]:
        return self.descriptors()

    def __contains__(self, item: Any) -> bool:
        if isinstance(item, DescriptorElement):
            return self.has_descriptor(item.uuid())
        return False

    def get_many_vectors(self, uuids: Iterable[Hashable]):
        return [d.vector() for d in self.get_many_descriptors(uuids)]

    @abc.abstractmethod
    def count(self) -> int: ...

    @abc.abstractmethod
    def clear(self) -> None: ...

    @abc.abstractmethod
    def has_descriptor(self, uuid: Hashable) -> bool: ...

    @abc.abstractmethod
    def add_descriptor(self, descriptor: DescriptorElement) -> None: ...

    @abc.abstractmethod
    def add_many_descriptors(self, descriptors: Iterable[DescriptorElement]) -> None: ...

    @abc.abstractmethod
    def get_descriptor(self, uuid: Hashable) -> DescriptorElement: ...

    @abc.abstractmethod
    def get_many_descriptors(self, uuids: Iterable[Hashable]) -> Iterator[DescriptorElement]: ...

    @abc.abstractmethod
    def remove_descriptor(self, uuid: Hashable) -> None: ...

    @abc.abstractmethod
    def remove_many_descriptors(self, uuids: Iterable[Hashable]) -> None: ...

    @abc.abstractmethod
    def keys(self) -> Iterator[Hashable]: ...

    @abc.abstractmethod
    def descriptors(self) -> Iterator[DescriptorElement]: ...

    @abc.abstractmethod
    def items(self) -> Iterator: ...
