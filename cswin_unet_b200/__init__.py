"""cswin_unet_b200 — B200-native (sm_100a) implementation of the CSWin-UNet hot path
(LePEAttention / CSWinBlock and the Merge_Block / CARAFE stages around them) behind the
reference's own nn.Module API.  Host code is Python/PyTorch (device memory, streams,
torch.distributed); all arithmetic runs in hand-written CUDA through the C ABI in
include/cswin_b200.h.  There is no CPU path and no library fallback.
"""
from . import synth  # noqa: F401  (pure numpy; safe without the extension)
from ._lib import CswinError, build, launch_count, lib, simt_fallback_count, tc_launch_count  # noqa: F401
from .engine import SliceEngine, predict_volume, shard_slices  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .model import CSWinTransformer, CSwinUnet, cswin_tiny_224  # noqa: F401
from .train import TrainStep, seg_loss  # noqa: F401
from .modules import (CARAFE, CARAFE4, CSWinBlock, DropPath, LePEAttention, Merge_Block, Mlp,  # noqa: F401
                      bump_param_epoch as invalidate_weight_caches, img2windows, windows2img)

__all__ = ["LePEAttention", "CSWinBlock", "Mlp", "Merge_Block", "CARAFE", "CARAFE4", "DropPath", "img2windows",
           "windows2img", "CSWinTransformer", "CSwinUnet", "cswin_tiny_224", "SliceEngine", "shard_slices", "predict_volume", "TrainStep", "seg_loss", "install", "uninstall", "build", "lib",
           "launch_count", "tc_launch_count", "simt_fallback_count", "CswinError", "synth", "invalidate_weight_caches"]
