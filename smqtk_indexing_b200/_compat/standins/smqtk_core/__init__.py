"""
Minimal stand-in for ``smqtk-core`` (0.18.x API surface used by SMQTK-Indexing).

Only used when the real ``smqtk_core`` distribution is not importable (it is not
vendored by the reference and there is no package index in the build image).
It carries NO arithmetic of the LSH hot path -- only the plugin / configuration
plumbing the reference's classes are written against
(reference call sites: smqtk_indexing/impls/nn_index/lsh.py:65-158,257-269).
"""
import abc
import importlib
import inspect
import os
from typing import Any, Dict, Set, Type, TypeVar

C = TypeVar("C", bound="Configurable")
P = TypeVar("P", bound="Pluggable")

ENTRYPOINT_GROUP = "smqtk_plugins"
ENV_PLUGIN_MODULES = "SMQTK_PLUGIN_PATH"


class Configurable(metaclass=abc.ABCMeta):
    """Objects constructible from / serialisable to JSON-compliant dicts."""

    __slots__ = ()

    @classmethod
    def get_default_config(cls) -> Dict[str, Any]:
        sig = inspect.signature(cls.__init__)
        cfg: Dict[str, Any] = {}
        for name, p in list(sig.parameters.items())[1:]:
            if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                continue
            cfg[name] = None if p.default is p.empty else p.default
        return cfg

    @classmethod
    def from_config(cls: Type[C], config_dict: Dict, merge_default: bool = True) -> C:
        if merge_default:
            from smqtk_core.dict import merge_dict
            config_dict = merge_dict(cls.get_default_config(), config_dict)
        return cls(**config_dict)

    @abc.abstractmethod
    def get_config(self) -> Dict[str, Any]:
        """JSON-compliant constructor configuration of this instance."""


def _all_subclasses(cls: type) -> Set[type]:
    out: Set[type] = set()
    stack = list(cls.__subclasses__())
    while stack:
        c = stack.pop()
        if c not in out:
            out.add(c)
            stack.extend(c.__subclasses__())
    return out


_EP_LOADED = False


def _load_plugin_modules() -> None:
    """Import modules advertised under the ``smqtk_plugins`` entry-point group
    and in ``$SMQTK_PLUGIN_PATH`` (colon separated module paths)."""
    global _EP_LOADED
    if _EP_LOADED:
        return
    _EP_LOADED = True
    try:
        from importlib import metadata
        eps = metadata.entry_points()
        group = eps.select(group=ENTRYPOINT_GROUP) if hasattr(eps, "select") \
            else eps.get(ENTRYPOINT_GROUP, [])
        for ep in group:
            try:
                ep.load()
            except Exception:  # pragma: no cover - broken third-party plugin
                pass
    except Exception:  # pragma: no cover
        pass
    for mod in filter(None, os.environ.get(ENV_PLUGIN_MODULES, "").split(":")):
        try:
            importlib.import_module(mod)
        except Exception:  # pragma: no cover
            pass


class Pluggable(metaclass=abc.ABCMeta):
    """Interface marker: concrete, usable subclasses are discoverable."""

    __slots__ = ()

    @classmethod
    def get_impls(cls: Type[P]) -> Set[Type[P]]:
        _load_plugin_modules()
        return {
            c for c in _all_subclasses(cls)
            if not inspect.isabstract(c) and c.is_usable()
        }

    @classmethod
    @abc.abstractmethod
    def is_usable(cls) -> bool:
        """Whether this implementation can run in the current environment."""

    def __init__(self) -> None:
        if not self.is_usable():
            raise RuntimeError("Implementation class '%s' is not currently "
                               "usable." % type(self).__name__)


class Plugfigurable(Pluggable, Configurable):
    __slots__ = ()
