"""Import shim that lets the UNMODIFIED reference (read-only at /root/reference) run on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing under ``patchrefinerv2_b200/`` may import this file.
It is used (a) by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/`` and (b) by ``tests/test_oracle_vs_reference.py`` (skipped when
``/root/reference`` is absent, e.g. on the GPU box) to pin the restated oracle in
``oracle/pr_oracle.py`` against the reference's own code.

The reference imports ~10 packages that are not installed here and do no arithmetic on the
inference path (mmengine, timm, matplotlib, kornia, skimage, imageio, prettytable,
torchmetrics, h5py; xformers stays absent so the eager attention path runs).  They are replaced by permissive stub modules; the three
mmengine names the path really uses (``Registry``, ``print_log``, ``ConfigDict``) get
minimal working implementations.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

def _find_root() -> str:
    """The reference tree: $PRV2_REFERENCE_ROOT, else /root/reference (build container), else the verbatim copy that
    oracle/build_ref.py leaves under oracle/_ref (the only one that exists on the GPU box)."""
    env = os.environ.get("PRV2_REFERENCE_ROOT")
    if env:
        return env
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for cand in ("/root/reference", here):
        if os.path.isdir(os.path.join(cand, "estimator")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUB_ROOTS = (
    "mmengine", "timm", "matplotlib", "kornia", "skimage", "imageio", "prettytable",
    "torchmetrics", "h5py", "wandb",
)


class _Anything:
    """Callable / subclassable placeholder returned for every unknown stub attribute."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # behaves as a transparent decorator
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        _populate(module)


class Registry:
    """Tiny stand-in for ``mmengine.Registry`` (dict + build(type -> cls(**cfg)))."""

    def __init__(self, name, parent=None, locations=None, **_):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self._modules[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self._modules[key]

    def build(self, cfg):
        cfg = dict(cfg)
        cls = self._modules[cfg.pop("type")]
        return cls(**cfg)


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __deepcopy__(self, memo):
        import copy
        return _AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ConfigDict(dict):
    """Attribute-access dict with recursive ``to_dict`` (what PatchRefiner.__init__ needs)."""

    def __init__(self, *a, **k):
        super().__init__()
        for key, val in dict(*a, **k).items():
            self[key] = self._wrap(val)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)

    def to_dict(self):
        # The reference reads nested entries by attribute *after* PretrainedConfig.from_dict
        # (patchrefiner.py:62-78: ``config.coarse_branch.type``), so nested levels must stay
        # attribute-accessible dict subclasses.
        def un(v):
            if isinstance(v, dict):
                return _AttrDict({kk: un(vv) for kk, vv in v.items()})
            if isinstance(v, (list, tuple)):
                return type(v)(un(x) for x in v)
            return v
        return dict(un(self))


def _print_log(msg, logger=None, level=None):
    if os.environ.get("PRV2_SHIM_VERBOSE"):
        print(msg)


def _populate(module):
    name = module.__name__
    if name == "mmengine":
        module.Registry = Registry
        module.print_log = _print_log
    elif name == "mmengine.config":
        module.ConfigDict = ConfigDict
    elif name == "mmengine.registry":
        module.Registry = Registry
        module.MODELS = Registry("mm_model")
        module.DATASETS = Registry("mm_dataset")
    elif name == "mmengine.logging":
        module.print_log = _print_log
    elif name == "mmengine.dist":
        module.get_dist_info = lambda: (0, 1)
    elif name in ("timm.models.layers", "timm.layers"):
        import torch.nn as nn

        class DropPath(nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        module.DropPath = DropPath
        module.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        module.trunc_normal_ = nn.init.trunc_normal_
        module.Conv2dSame = nn.Conv2d


_installed = False


def install():
    """Make ``import estimator`` work.  Idempotent.  Leaves cwd at the reference root because
    external/depth_anything/dpt.py:140 uses a relative torch-hub path."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True            # the tree is read-only
    import transformers  # noqa: F401  (must be imported before the timm stub exists)
    from transformers import PretrainedConfig  # noqa: F401
    import huggingface_hub  # noqa: F401
    sys.meta_path.insert(0, _StubFinder())
    for p in (os.path.join(REFERENCE_ROOT, "external"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.chdir(REFERENCE_ROOT)
    _installed = True


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "estimator"))


def build_reference_patchrefiner(cfg: dict, coarse_sd_path: str, fine_sd_path: str):
    """Build the reference ``PatchRefiner`` (estimator/models/patchrefiner.py:54) from a plain
    dict shaped like ``configs/patchrefiner_dav2/pr_u4k.py:10-53``.  ``*_sd_path`` are the
    DepthAnythingV2 state-dict files the constructor unconditionally torch.load()s
    (patchrefiner.py:94,118)."""
    install()
    import copy
    from estimator.models.patchrefiner import PatchRefiner
    c = copy.deepcopy(cfg)
    c["coarse_branch"]["pretrained"] = coarse_sd_path
    c["refiner"]["fine_branch"]["pretrained"] = fine_sd_path
    c.setdefault("pretrain_coarse_model", None)
    c.setdefault("pretrain_fine_model", None)
    c.setdefault("pretrained", None)
    c.setdefault("pre_norm_bbox", True)
    c.setdefault("sigloss", dict(type="SILogLoss"))
    model = PatchRefiner(ConfigDict(c))
    model.eval()
    return model


def build_reference_patchrefinerplus(cfg: dict, coarse_sd_path: str, encoder_factory):
    """Build the reference ``PatchRefinerPlus`` (estimator/models/patchrefinerplus.py:59) from a plain dict shaped like
    ``configs/patchrefinerv2_dav2/plus_mobile_u4k_*.py``.  ``timm.create_model`` (lightweight_refiner.py:259-262) is the one
    call into the absent timm package: it is answered with ``encoder_factory()`` (a 3-channel-stem module exposing
    ``conv_stem`` and ``default_cfg``; the reference then performs its own 4-channel stem surgery on it, :158-164)."""
    install()
    import copy
    import timm
    timm.create_model = lambda name, pretrained=False, features_only=True, **kw: encoder_factory()
    import estimator.models.blocks.lightweight_refiner as lwr
    lwr.timm.create_model = timm.create_model
    from estimator.models.patchrefinerplus import PatchRefinerPlus
    c = copy.deepcopy(cfg)
    c["coarse_branch"]["pretrained"] = coarse_sd_path
    model = PatchRefinerPlus(ConfigDict(c))
    model.eval()
    return model
