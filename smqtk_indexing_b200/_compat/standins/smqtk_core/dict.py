"""Stand-in for ``smqtk_core.dict`` (see package docstring)."""
import copy
from typing import Dict


def merge_dict(a: Dict, b: Dict, deep_copy: bool = False) -> Dict:
    """Merge ``b`` into ``a`` *in place* (nested dicts merged recursively) and
    return ``a``; with ``deep_copy`` the result is an independent copy."""
    if deep_copy:
        a = copy.deepcopy(a)
        b = copy.deepcopy(b)
    for k, v in b.items():
        if k in a and isinstance(a[k], dict) and isinstance(v, dict):
            merge_dict(a[k], v)
        else:
            a[k] = v
    return a
