"""
Compatibility shim for the three un-vendored SMQTK dependencies
(``smqtk-core``, ``smqtk-dataprovider``, ``smqtk-descriptors``; pinned by the
reference at pyproject.toml:39-41).

If the real distributions are importable they are used untouched.  Otherwise the
minimal stand-ins under ``_compat/standins`` are put on ``sys.path`` under the
same top-level names, so that both this package's plugin classes and (in the
build container only) the unmodified reference sources import the same
``Configurable`` / ``Pluggable`` / container types.
"""
import importlib
import os
import sys

_STANDIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")
USING_STANDINS = False


def ensure() -> bool:
    """Make ``smqtk_core``, ``smqtk_dataprovider`` and ``smqtk_descriptors``
    importable.  Returns True when the stand-ins are in use."""
    global USING_STANDINS
    try:
        importlib.import_module("smqtk_core")
        importlib.import_module("smqtk_dataprovider")
        importlib.import_module("smqtk_descriptors")
    except ImportError:
        if _STANDIN_DIR not in sys.path:
            sys.path.insert(0, _STANDIN_DIR)
        for m in ("smqtk_core", "smqtk_dataprovider", "smqtk_descriptors"):
            importlib.import_module(m)
        USING_STANDINS = True
    else:
        mod = sys.modules["smqtk_core"]
        USING_STANDINS = os.path.abspath(getattr(mod, "__file__", "")).startswith(_STANDIN_DIR)
    return USING_STANDINS


ensure()
