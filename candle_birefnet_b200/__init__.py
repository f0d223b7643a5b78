"""candle_birefnet_b200: B200-native (sm_100a) BiRefNet Swin-L forward behind the reference crate's API.

The hot path lives in libbirefnet_b200.so (CUDA, hand-written tcgen05/TMA kernels); this package is the host-side
mirror of `BiRefNet::new(BiRefNetConfig::swin_l(), vb)` / `forward_logits` plus the image-sharding driver.
"""
from .model import BiRefNet, BiRefNetConfig, SwinConfig  # noqa: F401
from ._lib import BrnError, lib  # noqa: F401
from . import ops  # noqa: F401
from .ops import DeformableConv2d  # noqa: F401  (src/lib.rs:13 re-export)

__all__ = ["BiRefNet", "BiRefNetConfig", "SwinConfig", "BrnError", "DeformableConv2d", "ops", "lib"]
