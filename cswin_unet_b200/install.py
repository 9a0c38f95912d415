"""`install()` — swap the native hot-path classes into a checkout of the reference.

    import sys; sys.path.insert(0, "/path/to/CSWin-UNet")
    import cswin_unet_b200; cswin_unet_b200.install()
    from networks.vision_transformer import CSwinUnet      # train.py / test.py unchanged from here on

`CSWinTransformer.__init__` (networks/cswin_unet.py:351-434) resolves `CSWinBlock`, `Merge_Block`, `CARAFE`,
`CARAFE4` as module globals at call time, and `CSWinBlock.__init__` resolves `LePEAttention` the same way, so
rebinding the names before the model is constructed is enough; parameter names, creation order (hence the
random init under a fixed seed) and state_dict keys are unchanged.
"""
from __future__ import annotations

import importlib
from typing import Iterable

from . import modules

HOT_PATH_CLASSES = ("LePEAttention", "CSWinBlock", "Merge_Block", "CARAFE", "CARAFE4", "img2windows", "windows2img")


def install(target: str = "networks.cswin_unet", names: Iterable[str] = HOT_PATH_CLASSES):
    """Rebind `names` inside the reference module `target`; returns {name: original object} for `uninstall`."""
    mod = importlib.import_module(target)
    saved = {}
    for n in names:
        saved[n] = getattr(mod, n)
        setattr(mod, n, getattr(modules, n))
    return saved


def uninstall(saved, target: str = "networks.cswin_unet") -> None:
    mod = importlib.import_module(target)
    for n, obj in saved.items():
        setattr(mod, n, obj)
