"""
smqtk_indexing_b200 -- B200-native implementation of SMQTK-Indexing's LSH
nearest-neighbour hot path (ITQ hashing -> Hamming scan / top-k -> re-rank)
behind the smqtk plugin API.  See DESIGN.md.
"""
from . import _compat  # noqa: F401  (makes smqtk_core / _dataprovider / _descriptors importable)
from .interfaces import HashIndex, LshFunctor, NearestNeighborsIndex  # noqa: F401

__version__ = "0.1.0"
