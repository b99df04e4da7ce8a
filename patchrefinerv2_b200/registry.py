"""Minimal model registry with the reference's call shape (estimator/registry/registry.py:7,
estimator/models/builder.py:6-8): ``MODELS.register_module()`` decorator and
``build_model(cfg)`` -> ``cls(**cfg_without_type)``."""
from __future__ import annotations


class Registry:
    def __init__(self, name: str):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force: bool = False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cls = self._modules.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry (have: {sorted(self._modules)})")
        return cls(**cfg)


MODELS = Registry("model")


def build_model(cfg):
    return MODELS.build(cfg)
