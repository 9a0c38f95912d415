"""Import the UNMODIFIED reference (BoloniniD/CSWin-UNet, networks/cswin_unet.py) as a comparator.

The reference is pure Python, so its two model files travel to the GPU box as `baseline/_ref/networks/*.py` (git-ignored, like the
built .so; `fetch()` — called by `__graft_entry__.build()` in the build container — copies them from /root/reference, nothing is
edited).  Used by: `bench.py --impl reference` / `--impl reference-cuda` / the `reference_cuda` key of the main line (the eager
PyTorch path the north star wants beaten, SURVEY 8d "Timing method"), and the live GPU parity tests (tests/test_gpu_reference.py:
`install()` into the reference's own assembly; trained-model bf16 acceptance).  Nothing in `cswin_unet_b200/` imports this.

The only thing injected is the 3-symbol `timm.models.layers` shim of SURVEY Appendix B (timm is not installed)."""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_SOURCE = "/root/reference"
FILES = ("networks/cswin_unet.py", "networks/vision_transformer.py")


def fetch(source: str = REF_SOURCE) -> Optional[str]:
    """Copy the reference's model files verbatim into baseline/_ref/ (build container only). Returns the directory or None."""
    if not os.path.isfile(os.path.join(source, FILES[0])):
        return REF_COPY if available() else None
    for rel in FILES:
        dst = os.path.join(REF_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source, rel), dst)
    return REF_COPY


def ref_dir() -> Optional[str]:
    for d in (REF_COPY, REF_SOURCE):
        if os.path.isfile(os.path.join(d, FILES[0])):
            return d
    return None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_COPY, FILES[0])) or os.path.isfile(os.path.join(REF_SOURCE, FILES[0]))


def install_timm_shim() -> None:
    if "timm.models.layers" in sys.modules:
        return
    import torch

    class DropPath(torch.nn.Module):                       # timm.models.layers.DropPath semantics and RNG consumption
        def __init__(self, drop_prob=0., scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0. or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                m.div_(keep)
            return x * m

    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: (x, x)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def import_reference():
    """-> the reference's `networks.cswin_unet` module (unmodified source), or raises RuntimeError when it is not available."""
    d = ref_dir()
    if d is None:
        raise RuntimeError("the reference is not available: neither baseline/_ref/networks/cswin_unet.py (run __graft_entry__.build() "
                           "in the build container) nor /root/reference exists")
    install_timm_shim()
    if d not in sys.path:
        sys.path.insert(0, d)
    mod = sys.modules.get("networks.cswin_unet")
    if mod is None or not getattr(mod, "__file__", "").startswith(d):
        for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
            del sys.modules[k]
        mod = importlib.import_module("networks.cswin_unet")
    return mod


T224 = dict(img_size=224, patch_size=4, in_chans=3, num_classes=9, embed_dim=64, depth=[1, 2, 9, 1], split_size=[1, 2, 7, 7],
            num_heads=[2, 4, 8, 16], mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., drop_path_rate=0.2)


def build_reference_model(**overrides):
    """CSWinTransformer of the reference with the cswin_tiny_224_lite hyper-parameters (configs/cswin_tiny_224_lite.yaml:4-10,
    config.py:49-66; the yacs config system itself is not installed) — constructed directly, as SURVEY 8d config 1 does."""
    import contextlib
    import io
    ref = import_reference()
    hp = dict(T224)
    hp.update(overrides)
    with contextlib.redirect_stdout(io.StringIO()):        # the constructor prints `depth [...]` (cswin_unet.py:349)
        return ref.CSWinTransformer(**hp)
