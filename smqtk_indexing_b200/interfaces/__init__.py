"""
Plugin interfaces of the LSH path.

If the upstream ``smqtk_indexing`` distribution is importable its own abstract
bases are re-exported, so the B200 implementations register as implementations
of the *same* interfaces (``HashIndex.get_impls()`` then finds them and they can
be named in configs consumed by upstream's ``LSHNearestNeighborIndex.from_config``,
reference: smqtk_indexing/impls/nn_index/lsh.py:88-97,135-156).  Otherwise the
mirrors defined in this package are used; they keep upstream's method names,
argument meaning and exception contract.
"""
from smqtk_indexing_b200 import _compat  # noqa: F401  (makes smqtk_* importable)

try:  # pragma: no cover - depends on the environment
    from smqtk_indexing.interfaces.hash_index import HashIndex
    from smqtk_indexing.interfaces.lsh_functor import LshFunctor
    from smqtk_indexing.interfaces.nearest_neighbor_index import NearestNeighborsIndex
    UPSTREAM_INTERFACES = True
except ImportError:
    from .hash_index import HashIndex
    from .lsh_functor import LshFunctor
    from .nearest_neighbor_index import NearestNeighborsIndex
    UPSTREAM_INTERFACES = False

__all__ = ["HashIndex", "LshFunctor", "NearestNeighborsIndex", "UPSTREAM_INTERFACES"]
