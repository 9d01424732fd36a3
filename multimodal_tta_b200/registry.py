"""Plugin surface: the reference's name->class registries (/root/reference/src/registry.py).

When this package is imported inside the reference repo (``src.registry`` importable) the
reference's own registry objects are used, so ``@register_model("unet_b200")`` lands in the very
``MODELS`` map that ``ExperimentManager.setup_model`` reads
(/root/reference/src/core/experiment_manager.py:88-96).  Stand-alone (tests, bench, GPU box) a
behaviour-identical mirror is used: duplicate name -> printed warning + overwrite
(registry.py:30-32), unknown name -> KeyError (registry.py:42-43); pinned by
tests/golden/registry_golden.json.
"""
from __future__ import annotations

from typing import Callable, Dict, Type

try:  # pragma: no cover - only inside the reference checkout
    from src.registry import (EVALUATION_STRATEGIES, MODELS, PLUGINS, Registry,  # type: ignore
                              get_evaluation_strategy, get_model, get_plugin,
                              register_evaluation_strategy, register_model, register_plugin)

    USING_REFERENCE_REGISTRY = True
except Exception:  # ImportError or a partial reference tree
    USING_REFERENCE_REGISTRY = False

    class Registry:  # type: ignore[no-redef]
        def __init__(self, name: str):
            self.name = name
            self._registry: Dict[str, Type] = {}

        def register(self, name: str, cls: Type = None) -> Callable:
            def _register(c: Type) -> Type:
                if name in self._registry:
                    print(f"Warning: {name} is already registered in {self.name}")
                self._registry[name] = c
                return c

            return _register(cls) if cls is not None else _register

        def get(self, name: str) -> Type:
            if name not in self._registry:
                raise KeyError(f"{name} is not registered in {self.name}")
            return self._registry[name]

        def has(self, name: str) -> bool:
            return name in self._registry

        def list_all(self) -> list:
            return list(self._registry.keys())

        def clear(self) -> None:
            self._registry.clear()

    MODELS = Registry("models")
    EVALUATION_STRATEGIES = Registry("evaluation_strategies")
    PLUGINS = Registry("plugins")

    def register_model(name: str):
        return MODELS.register(name)

    def register_evaluation_strategy(name: str):
        return EVALUATION_STRATEGIES.register(name)

    def register_plugin(name: str):
        return PLUGINS.register(name)

    def get_model(name: str) -> Type:
        return MODELS.get(name)

    def get_evaluation_strategy(name: str) -> Type:
        return EVALUATION_STRATEGIES.get(name)

    def get_plugin(name: str) -> Type:
        return PLUGINS.get(name)
