"""Registry API of the drop-in boundary.

The reference builds its operators through ``mmengine.Registry`` objects (reference
fish_diffusion/archs/diffsinger/diffusions/builder.py:1-15, fish_diffusion/modules/vocoders/builder.py:1-3):
``DIFFUSIONS.build(cfg)``, ``DENOISERS.build(cfg)``, ``VOCODERS.build(cfg)`` pop ``type`` from the config dict and
instantiate the registered class with the remaining keys.  mmengine is used when it is importable; otherwise this
file provides the small compatible subset the path needs (same method names, same error types).
"""
from __future__ import annotations

import copy

try:  # pragma: no cover - mmengine is absent from the build image
    from mmengine import Registry as _MMRegistry  # type: ignore
except Exception:  # noqa: BLE001
    _MMRegistry = None


class _MiniRegistry:
    """Subset of mmengine.Registry: register_module (call or decorator form), build, get, __contains__."""

    def __init__(self, name: str):
        self._name = name
        self._module_dict: dict[str, type] = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f"Registry(name={self._name}, items={sorted(self._module_dict)})"

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, module, name=None, force=False):
        if not callable(module):
            raise TypeError(f"module must be Callable, but got {type(module)}")
        names = [module.__name__] if name is None else ([name] if isinstance(name, str) else list(name))
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f"{n} is already registered in {self._name} at {self._module_dict[n].__module__}")
            self._module_dict[n] = module

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if module is not None:
            self._register(module=module, name=name, force=force)
            return module

        def _decorator(cls):
            self._register(module=cls, name=name, force=force)
            return cls

        return _decorator

    def build(self, cfg, *args, **kwargs):
        if not isinstance(cfg, dict):
            raise TypeError(f"cfg should be a dict, ConfigDict or Config, but got {type(cfg)}")
        if "type" not in cfg:
            raise KeyError(f'`cfg` must contain the key "type", but got {cfg}')
        args_ = copy.copy(dict(cfg))
        obj_type = args_.pop("type")
        if isinstance(obj_type, str):
            obj_cls = self.get(obj_type)
            if obj_cls is None:
                raise KeyError(f"{obj_type} is not in the {self._name} registry. "
                               f"Please check whether the value of `{obj_type}` is correct.")
        elif callable(obj_type):
            obj_cls = obj_type
        else:
            raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
        return obj_cls(*args, **args_, **kwargs)


Registry = _MMRegistry if _MMRegistry is not None else _MiniRegistry

DIFFUSIONS = Registry("diffusions")
DENOISERS = Registry("denoisers")
VOCODERS = Registry("vocoders")
