"""Stand-in for ``smqtk_dataprovider.impls.data_element.memory``."""
import threading
from typing import Any, Dict, Optional

from smqtk_dataprovider import DataElement
from smqtk_dataprovider.exceptions import ReadOnlyError

BYTES_CONFIG_ENCODING = "latin-1"


class DataMemoryElement(DataElement):
    """In-memory byte blob."""

    @classmethod
    def is_usable(cls) -> bool:
        return True

    @classmethod
    def from_config(cls, config_dict: Dict, merge_default: bool = True) -> "DataMemoryElement":
        config_dict = dict(config_dict)
        b = config_dict.get("bytes")
        if isinstance(b, str):
            config_dict["bytes"] = b.encode(BYTES_CONFIG_ENCODING)
        return super().from_config(config_dict, merge_default)

    def __init__(self, bytes: Optional[bytes] = None,
                 content_type: Optional[str] = None, readonly: bool = False):
        super().__init__()
        self._bytes = bytes
        self._content_type = content_type
        self._readonly = bool(readonly)
        self._lock = threading.RLock()

    def __repr__(self) -> str:
        return "DataMemoryElement{len(bytes): %d, content_type: %s, readonly: %s}" \
            % (len(self._bytes or b""), self._content_type, self._readonly)

    def get_config(self) -> Dict[str, Any]:
        b = self._bytes
        return {
            "bytes": b.decode(BYTES_CONFIG_ENCODING) if b is not None else None,
            "content_type": self._content_type,
            "readonly": self._readonly,
        }

    def content_type(self) -> Optional[str]:
        return self._content_type

    def is_empty(self) -> bool:
        return not bool(self._bytes)

    def get_bytes(self) -> bytes:
        return self._bytes or b""

    def writable(self) -> bool:
        return not self._readonly

    def set_bytes(self, b: bytes) -> None:
        if not self.writable():
            raise ReadOnlyError("This memory element cannot be written to.")
        with self._lock:
            self._bytes = b
