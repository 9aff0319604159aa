"""Stand-in for ``smqtk_core.configuration`` (see package docstring).

Config dict shape (reference: smqtk_indexing/impls/nn_index/lsh.py:88-156):
``{"type": "<module>.<Class>" | None, "<module>.<Class>": {...ctor kwargs...}}``
"""
import inspect
import json
from typing import Any, Dict, Iterable, Optional, Sequence, Tuple, Type, TypeVar

from smqtk_core import Configurable

T = TypeVar("T", bound=Configurable)


def _type_key(cls: type) -> str:
    return "%s.%s" % (cls.__module__, cls.__name__)


def make_default_config(configurable_iter: Iterable[Type[Configurable]]) -> Dict[str, Any]:
    d: Dict[str, Any] = {"type": None}
    for cls in configurable_iter:
        d[_type_key(cls)] = cls.get_default_config()
    return d


def cls_conf_to_config_dict(cls: type, conf: Dict) -> Dict[str, Any]:
    k = _type_key(cls)
    return {"type": k, k: conf}


def to_config_dict(c_inst: Configurable) -> Dict[str, Any]:
    if isinstance(c_inst, type) or not isinstance(c_inst, Configurable):
        raise ValueError("c_inst must be an instance and its type must "
                         "subclass from Configurable.")
    return cls_conf_to_config_dict(type(c_inst), c_inst.get_config())


def cls_conf_from_config_dict(config: Dict, type_iter: Iterable[type]) -> Tuple[type, Dict]:
    if "type" not in config:
        raise ValueError("Configuration dictionary given does not have an "
                         "implementation type specification.")
    t = config["type"]
    type_map = {_type_key(c): c for c in type_iter}
    # Also accept bare class names, as smqtk-core does.
    by_name = {c.__name__: c for c in type_map.values()}
    if t is None:
        raise ValueError("No implementation type specified. Options: %s"
                         % sorted(type_map))
    if t not in config:
        raise ValueError("Implementation type specified as '%s', but no "
                         "configuration block was present for that type." % t)
    cls = type_map.get(t, by_name.get(t))
    if cls is None:
        raise ValueError("Implementation type specified as '%s', but no "
                         "plugin implementations are available for that type. "
                         "Available: %s" % (t, sorted(type_map)))
    return cls, config[t]


def from_config_dict(config: Dict, type_iter: Iterable[Type[T]], *args: Any) -> T:
    cls, conf = cls_conf_from_config_dict(config, type_iter)
    if not (isinstance(cls, type) and issubclass(cls, Configurable)):
        raise ValueError("Resolved type is not Configurable: %r" % (cls,))
    return cls.from_config(conf, *args)


def configuration_test_helper(
    inst: T,
    different_ctor_params: Optional[Sequence[str]] = None,
    different_default_config_params: Optional[Sequence[str]] = None,
) -> Tuple[T, T, T]:
    """Round-trip ``inst`` through ``get_config`` / ``from_config`` twice and
    check stability; returns (inst, copy1, copy2) for attribute checks."""
    cls = type(inst)
    ctor_params = set(list(inspect.signature(cls.__init__).parameters)[1:])
    if different_ctor_params is None:
        assert set(cls.get_default_config()) == ctor_params, \
            "default config keys differ from constructor parameters"
    cfg0 = inst.get_config()
    assert json.loads(json.dumps(cfg0)) == cfg0, "config not JSON compliant"
    i1 = cls.from_config(cfg0)
    cfg1 = i1.get_config()
    assert cfg1 == cfg0, "config changed after first round trip"
    i2 = cls.from_config(cfg1)
    assert i2.get_config() == cfg0, "config changed after second round trip"
    return inst, i1, i2
