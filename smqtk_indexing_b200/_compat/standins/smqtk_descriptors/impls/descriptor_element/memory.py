"""Stand-in for ``smqtk_descriptors.impls.descriptor_element.memory``."""
from typing import Any, Dict, Hashable, Optional

import numpy

from smqtk_descriptors import DescriptorElement


class DescriptorMemoryElement(DescriptorElement):
    """Descriptor vector held in process memory."""

    @classmethod
    def is_usable(cls) -> bool:
        return True

    def __init__(self, uuid: Hashable):
        super().__init__(uuid)
        self._v: Optional[numpy.ndarray] = None

    def get_config(self) -> Dict[str, Any]:
        return {}

    def has_vector(self) -> bool:
        return self._v is not None

    def vector(self) -> Optional[numpy.ndarray]:
        return None if self._v is None else numpy.copy(self._v)

    def set_vector(self, new_vec: Any) -> "DescriptorMemoryElement":
        self._v = None if new_vec is None else numpy.copy(numpy.asarray(new_vec))
        return self
