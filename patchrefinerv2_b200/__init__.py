"""patchrefinerv2_b200: B200-native (sm_100a) implementation of PatchRefinerV2's tiled
high-resolution inference hot path behind the reference's estimator-model API."""
from .registry import MODELS, build_model  # noqa: F401
from .model import PatchRefiner, PatchRefinerPlus  # noqa: F401
from .bifusion import BiDirectionalFusion, BiDirectionalFusionHeavy  # noqa: F401
from .zoe import ZoeDepthBinsHead  # noqa: F401

__all__ = ["MODELS", "build_model", "PatchRefiner", "PatchRefinerPlus", "BiDirectionalFusion"]
