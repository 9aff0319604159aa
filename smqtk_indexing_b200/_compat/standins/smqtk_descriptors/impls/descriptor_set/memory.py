"""Stand-in for ``smqtk_descriptors.impls.descriptor_set.memory``."""
import threading
from typing import Any, Dict, Hashable, Iterable, Iterator, Optional

from smqtk_descriptors import DescriptorElement, DescriptorSet


class MemoryDescriptorSet(DescriptorSet):
    """Dictionary-backed descriptor set (no persistence in the stand-in)."""

    @classmethod
    def is_usable(cls) -> bool:
        return True

    def __init__(self, cache_element: Optional[Any] = None, pickle_protocol: int = -1):
        super().__init__()
        self.cache_element = cache_element
        self.pickle_protocol = pickle_protocol
        self._table: Dict[Hashable, DescriptorElement] = {}
        self._lock = threading.RLock()

    def get_config(self) -> Dict[str, Any]:
        return {"cache_element": None, "pickle_protocol": self.pickle_protocol}

    def count(self) -> int:
        return len(self._table)

    def clear(self) -> None:
        with self._lock:
            self._table = {}

    def has_descriptor(self, uuid: Hashable) -> bool:
        return uuid in self._table

    def add_descriptor(self, descriptor: DescriptorElement) -> None:
        with self._lock:
            self._table[descriptor.uuid()] = descriptor

    def add_many_descriptors(self, descriptors: Iterable[DescriptorElement]) -> None:
        with self._lock:
            for d in descriptors:
                self._table[d.uuid()] = d

    def get_descriptor(self, uuid: Hashable) -> DescriptorElement:
        return self._table[uuid]

    def get_many_descriptors(self, uuids: Iterable[Hashable]) -> Iterator[DescriptorElement]:
        for uid in uuids:
            yield self._table[uid]

    def remove_descriptor(self, uuid: Hashable) -> None:
        with self._lock:
            del self._table[uuid]

    def remove_many_descriptors(self, uuids: Iterable[Hashable]) -> None:
        with self._lock:
            uuids = list(uuids)
            for uid in uuids:
                if uid not in self._table:
                    raise KeyError(uid)
            for uid in uuids:
                del self._table[uid]

    def keys(self) -> Iterator[Hashable]:
        return iter(list(self._table.keys()))

    def descriptors(self) -> Iterator[DescriptorElement]:
        return iter(list(self._table.values()))

    def items(self) -> Iterator:
        return iter(list(self._table.items()))
