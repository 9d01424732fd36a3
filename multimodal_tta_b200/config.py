"""Config access with the reference's semantics (/root/reference/src/utils/config.py:7-32).

Under real Hydra/OmegaConf the reference's own ``DictConfig``/``OmegaConf.select`` are used
unchanged.  When omegaconf is not installed (this image) a small duck-typed ``DictConfig``
shim with dotted-path ``select`` stands in, so that the plugin classes and the tests read the
same way in both environments.
"""
from __future__ import annotations

import os
from typing import Any, Mapping, Optional, Type

try:  # pragma: no cover - exercised only where omegaconf exists
    from omegaconf import DictConfig, ListConfig, OmegaConf  # type: ignore

    HAVE_OMEGACONF = True

    def create(obj: Any = None) -> "DictConfig":
        return OmegaConf.create({} if obj is None else obj)

    def select(cfg: "DictConfig", path: str) -> Any:
        return OmegaConf.select(cfg, path)

except ImportError:
    HAVE_OMEGACONF = False

    class DictConfig(dict):  # type: ignore[no-redef]
        """Attribute-access dict; nested mappings are wrapped recursively."""

        def __init__(self, content: Optional[Mapping] = None):
            super().__init__()
            for k, v in (content or {}).items():
                self[k] = v

        def __setitem__(self, k, v):
            super().__setitem__(k, _wrap(v))

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    ListConfig = list  # type: ignore[misc,assignment]

    def _wrap(v: Any) -> Any:
        if isinstance(v, DictConfig):
            return v
        if isinstance(v, Mapping):
            return DictConfig(v)
        if isinstance(v, (list, tuple)):
            return [_wrap(x) for x in v]
        return v

    def create(obj: Any = None) -> "DictConfig":
        return DictConfig(obj or {})

    def select(cfg: "DictConfig", path: str) -> Any:
        cur: Any = cfg
        for part in path.split("."):
            if isinstance(cur, Mapping) and part in cur:
                cur = cur[part]
            else:
                return None
        return cur


def require_config(cfg: "DictConfig", path: str, type_: Optional[Type] = None) -> Any:
    """Required value; ValueError when missing/None, TypeError on a non-DictConfig or wrong type."""
    if not isinstance(cfg, DictConfig):
        raise TypeError(f"`cfg` must be DictConfig, got {type(cfg).__name__}")
    value = select(cfg, path)
    if value is None:
        raise ValueError(f"Required configuration missing: {path}")
    if type_ is not None and not isinstance(value, type_):
        raise TypeError(f"Config '{path}' must be {type_.__name__}, got {type(value).__name__}")
    return value


def get_config(cfg: "DictConfig", path: str, default: Any = None, type_: Optional[Type] = None) -> Any:
    """Optional value with default; same type rules as the reference."""
    if not isinstance(cfg, DictConfig):
        raise TypeError(f"`cfg` must be DictConfig, got {type(cfg).__name__}")
    value = select(cfg, path)
    value = default if value is None else value
    if type_ is not None and value is not None and not isinstance(value, type_):
        raise TypeError(f"Config '{path}' must be {type_.__name__}, got {type(value).__name__}")
    return value


def _deep_merge(dst: dict, src: Mapping) -> dict:
    for k, v in src.items():
        if isinstance(v, Mapping) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = dict(v) if isinstance(v, Mapping) else v
    return dst


def compose_yaml(*paths: str, overrides: Optional[Mapping] = None) -> "DictConfig":
    """Tiny stand-in for Hydra composition used by tests and the CLI shim: deep-merge YAML files
    in order (later wins; a leading ``# @package _global_`` file merges at the root, any other
    file merges under its group name = parent directory name), then apply ``overrides``."""
    import yaml

    root: dict = {}
    for p in paths:
        with open(p, "r") as f:
            text = f.read()
        data = yaml.safe_load(text) or {}
        data.pop("defaults", None)
        first = text.lstrip().splitlines()[0] if text.strip() else ""
        if first.replace(" ", "").startswith("#@package_global_"):
            _deep_merge(root, data)
        else:
            group = os.path.basename(os.path.dirname(os.path.abspath(p)))
            _deep_merge(root.setdefault(group, {}), data)
    if overrides:
        _deep_merge(root, overrides)
    return create(root)
