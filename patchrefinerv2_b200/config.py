"""Minimal loader for the reference's mmengine-style python configs (mmengine is not a dependency here).

Covers what ``tools/test.py`` needs (README.md:57-77 of the reference): ``Config.fromfile`` with recursive
``_base_`` inheritance (configs/patchrefiner_dav2/pr_u4k.py:1-5), dict-merge semantics with the ``_delete_`` key,
attribute access (``cfg.model.config``), and ``--cfg-option a.b.c=value`` dotted overrides
(docs/user_infer.md:113-130).  Config files are plain Python executed in an empty namespace; every top-level
name that does not start with an underscore and is not a module becomes a config key.
"""
from __future__ import annotations

import ast
import copy
import os
import types
from typing import Any, Dict, Iterable, List

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    """dict with attribute access, recursively (mmengine.config.ConfigDict behaviour used by the reference)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = _wrap(value)

    def to_dict(self) -> dict:
        return _unwrap(self)

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, ConfigDict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_unwrap(x) for x in v)
    return v


def merge_dict(child: dict, base: dict) -> dict:
    """``child`` overrides ``base`` key by key; nested dicts merge unless the child carries ``_delete_=True``."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get(DELETE_KEY, False):
            out[k] = merge_dict(v, out[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            out[k] = copy.deepcopy(v)
    return out


def _load_file(path: str, _stack: tuple = ()) -> dict:
    path = os.path.abspath(path)
    if path in _stack:
        raise ValueError(f"circular _base_ chain: {' -> '.join(_stack + (path,))}")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    ns: Dict[str, Any] = {"__file__": path}
    with open(path, "r") as fh:
        exec(compile(fh.read(), path, "exec"), ns)          # the reference's configs are trusted python files
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop(BASE_KEY, [])
    if isinstance(bases, str):
        bases = [bases]
    merged: dict = {}
    for b in bases:
        b_cfg = _load_file(os.path.join(os.path.dirname(path), b), _stack + (path,))
        dup = set(merged) & set(b_cfg)
        if dup:
            raise KeyError(f"duplicate keys {sorted(dup)} in the _base_ files of {path}")
        merged.update(b_cfg)
    return merge_dict(cfg, merged)


def parse_option_value(text: str):
    """``--cfg-option`` values: python literals when they parse (numbers, lists, tuples, None, True, quoted strings), else the raw string."""
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def parse_cfg_options(items: Iterable[str]) -> Dict[str, Any]:
    out = {}
    for it in items or []:
        if "=" not in it:
            raise ValueError(f"--cfg-option expects key=value, got {it!r}")
        k, v = it.split("=", 1)
        out[k.strip()] = parse_option_value(v.strip())
    return out


class Config(ConfigDict):
    filename: str = ""

    @classmethod
    def fromfile(cls, path: str) -> "Config":
        cfg = cls(_wrap(_load_file(path)))
        object.__setattr__(cfg, "filename", os.path.abspath(path))
        return cfg

    def merge_from_dict(self, options: Dict[str, Any]) -> None:
        """Dotted-key overrides (``general_dataloader.dataset.rgb_image_dir='./examples/'``); list indices allowed."""
        for key, value in options.items():
            node: Any = self
            parts: List[str] = key.split(".")
            for p in parts[:-1]:
                if isinstance(node, list):
                    node = node[int(p)]
                else:
                    if p not in node or not isinstance(node[p], (dict, list)):
                        node[p] = ConfigDict()
                    node = node[p]
            last = parts[-1]
            if isinstance(node, list):
                node[int(last)] = _wrap(value)
            else:
                node[last] = _wrap(value)
