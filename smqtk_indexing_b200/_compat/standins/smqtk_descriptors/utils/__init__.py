"""Stand-in for ``smqtk_descriptors.utils``."""
from typing import Any, Callable, Iterable, Iterator


def parallel_map(work_func: Callable, *sequences: Iterable, **kwargs: Any) -> Iterator:
    """Order-preserving map.  The real implementation fans work out to a
    thread/process pool; the reference only relies on the results arriving in
    input order (lsh.py:507-513 zips them with their inputs)."""
    return map(work_func, *sequences)
