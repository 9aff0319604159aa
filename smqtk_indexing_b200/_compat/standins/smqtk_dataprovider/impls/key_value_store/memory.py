"""Stand-in for ``smqtk_dataprovider.impls.key_value_store.memory``."""
import threading
from typing import Any, Dict, Hashable, Iterable, Iterator, Mapping, Optional

from smqtk_dataprovider import KeyValueStore, NO_DEFAULT_VALUE


class MemoryKeyValueStore(KeyValueStore):
    """Dictionary-backed key-value store (no persistence in the stand-in)."""

    @classmethod
    def is_usable(cls) -> bool:
        return True

    def __init__(self, cache_element: Optional[Any] = None):
        super().__init__()
        self._cache_element = cache_element
        self._table: Dict[Hashable, Any] = {}
        self._table_lock = threading.RLock()

    def __repr__(self) -> str:
        return super().__repr__() + "[cache_element: %s]>" % self._cache_element

    def get_config(self) -> Dict[str, Any]:
        return {"cache_element": None}

    def count(self) -> int:
        with self._table_lock:
            return len(self._table)

    def keys(self) -> Iterator[Hashable]:
        with self._table_lock:
            return iter(list(self._table.keys()))

    def values(self) -> Iterator[Any]:
        with self._table_lock:
            return iter(list(self._table.values()))

    def is_read_only(self) -> bool:
        ce = self._cache_element
        return bool(ce is not None and ce.is_read_only())

    def has(self, key: Hashable) -> bool:
        with self._table_lock:
            return key in self._table

    def add(self, key: Hashable, value: Any) -> "MemoryKeyValueStore":
        super().add(key, value)
        with self._table_lock:
            self._table[key] = value
        return self

    def add_many(self, d: Mapping[Hashable, Any]) -> "MemoryKeyValueStore":
        super().add_many(d)
        with self._table_lock:
            self._table.update(d)
        return self

    def remove(self, key: Hashable) -> "MemoryKeyValueStore":
        super().remove(key)
        with self._table_lock:
            del self._table[key]
        return self

    def remove_many(self, keys: Iterable[Hashable]) -> "MemoryKeyValueStore":
        super().remove_many(keys)
        keys = set(keys)
        with self._table_lock:
            missing = keys.difference(self._table)
            if missing:
                raise KeyError(missing)
            for k in keys:
                del self._table[k]
        return self

    def get(self, key: Hashable, default: Any = NO_DEFAULT_VALUE) -> Any:
        with self._table_lock:
            if default is NO_DEFAULT_VALUE:
                return self._table[key]
            return self._table.get(key, default)

    def clear(self) -> "MemoryKeyValueStore":
        super().clear()
        with self._table_lock:
            self._table.clear()
        return self
