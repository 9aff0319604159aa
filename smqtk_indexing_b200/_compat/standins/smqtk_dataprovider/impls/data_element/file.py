"""Stand-in for ``smqtk_dataprovider.impls.data_element.file``."""
import os
from typing import Any, Dict, Optional

from smqtk_dataprovider import DataElement
from smqtk_dataprovider.exceptions import ReadOnlyError
from smqtk_dataprovider.utils.file import safe_create_dir


class DataFileElement(DataElement):
    """Bytes stored in a file on the local filesystem."""

    @classmethod
    def is_usable(cls) -> bool:
        return True

    def __init__(self, filepath: str, readonly: bool = False,
                 explicit_mimetype: Optional[str] = None):
        super().__init__()
        self._filepath = os.path.abspath(os.path.expanduser(filepath))
        self._readonly = bool(readonly)
        self._explicit_mimetype = explicit_mimetype

    def __repr__(self) -> str:
        return "DataFileElement{filepath: %s, readonly: %s}" % (self._filepath, self._readonly)

    def get_config(self) -> Dict[str, Any]:
        return {"filepath": self._filepath, "readonly": self._readonly,
                "explicit_mimetype": self._explicit_mimetype}

    def content_type(self) -> Optional[str]:
        return self._explicit_mimetype

    def is_empty(self) -> bool:
        return not (os.path.exists(self._filepath) and os.path.getsize(self._filepath) > 0)

    def get_bytes(self) -> bytes:
        if self.is_empty():
            return b""
        with open(self._filepath, "rb") as f:
            return f.read()

    def writable(self) -> bool:
        return not self._readonly

    def set_bytes(self, b: bytes) -> None:
        if not self.writable():
            raise ReadOnlyError("This file element is read only.")
        safe_create_dir(os.path.dirname(self._filepath))
        tmp = self._filepath + ".tmp"
        with open(tmp, "wb") as f:
            f.write(b)
        os.replace(tmp, self._filepath)
