"""
Minimal stand-in for ``smqtk-dataprovider`` 0.16 (DataElement / KeyValueStore
surface used by SMQTK-Indexing: itq.py:213-237, linear.py:126-142,
lsh.py:279,323,372,439-441,486,494).  No hot-path arithmetic lives here.
"""
import abc
from typing import Any, Dict, Hashable, Iterable, Iterator, Mapping, Optional

from smqtk_core import Configurable, Pluggable
from smqtk_dataprovider.exceptions import ReadOnlyError, InvalidUriError  # noqa: F401


class DataElement(Configurable, Pluggable):
    """Abstract blob of bytes with a content type."""

    def __bool__(self) -> bool:
        return True

    @abc.abstractmethod
    def content_type(self) -> Optional[str]: ...

    @abc.abstractmethod
    def is_empty(self) -> bool: ...

    @abc.abstractmethod
    def get_bytes(self) -> bytes: ...

    @abc.abstractmethod
    def writable(self) -> bool: ...

    @abc.abstractmethod
    def set_bytes(self, b: bytes) -> None: ...

    def is_read_only(self) -> bool:
        return not self.writable()


def from_uri(uri: str, impl_generator: Any = None) -> DataElement:
    from smqtk_dataprovider.impls.data_element.file import DataFileElement
    if uri.startswith("file://"):
        return DataFileElement(uri[len("file://"):])
    if "://" not in uri:
        return DataFileElement(uri)
    raise InvalidUriError(uri, "no stand-in implementation resolves this URI")


class _NoDefault:
    pass


NO_DEFAULT_VALUE = _NoDefault()


class KeyValueStore(Configurable, Pluggable):
    """Abstract key -> value mapping (keys hashable)."""

    def __len__(self) -> int:
        return self.count()

    def __contains__(self, item: Hashable) -> bool:
        return self.has(item)

    @abc.abstractmethod
    def __repr__(self) -> str:
        return "<%s" % type(self).__name__

    @abc.abstractmethod
    def count(self) -> int: ...

    @abc.abstractmethod
    def keys(self) -> Iterator[Hashable]: ...

    def values(self) -> Iterator[Any]:
        for k in self.keys():
            yield self.get(k)

    @abc.abstractmethod
    def is_read_only(self) -> bool: ...

    @abc.abstractmethod
    def has(self, key: Hashable) -> bool: ...

    @abc.abstractmethod
    def add(self, key: Hashable, value: Any) -> "KeyValueStore":
        if self.is_read_only():
            raise ReadOnlyError("Cannot add to read-only instance %s." % self)
        return self

    @abc.abstractmethod
    def add_many(self, d: Mapping[Hashable, Any]) -> "KeyValueStore":
        if self.is_read_only():
            raise ReadOnlyError("Cannot add to read-only instance %s." % self)
        return self

    @abc.abstractmethod
    def remove(self, key: Hashable) -> "KeyValueStore":
        if self.is_read_only():
            raise ReadOnlyError("Cannot remove from read-only instance %s." % self)
        return self

    @abc.abstractmethod
    def remove_many(self, keys: Iterable[Hashable]) -> "KeyValueStore":
        if self.is_read_only():
            raise ReadOnlyError("Cannot remove from read-only instance %s." % self)
        return self

    @abc.abstractmethod
    def get(self, key: Hashable, default: Any = NO_DEFAULT_VALUE) -> Any: ...

    def get_many(self, keys: Iterable[Hashable], default: Any = NO_DEFAULT_VALUE) -> Iterable[Any]:
        for k in keys:
            yield self.get(k, default)

    @abc.abstractmethod
    def clear(self) -> "KeyValueStore":
        if self.is_read_only():
            raise ReadOnlyError("Cannot clear a read-only %s instance."
                                % type(self).__name__)
        return self
