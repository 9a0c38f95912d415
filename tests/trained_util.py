"""Weights / inputs of the briefly-trained reference fixtures (tests/golden/trained_*.npz, written by
tests/golden/make_trained_golden.py): the trained sub-network is stored, every other tensor is the `synth` reference-init value
regenerated from its name."""
import numpy as np
import torch

from cswin_unet_b200 import synth
from tests import golden_util as G

CONFIGS = {
    "t224": dict(img_size=224, num_classes=9, split_size=[1, 2, 7, 7], n_test=4),
    "512": dict(img_size=512, num_classes=3, split_size=[1, 2, 8, 8], n_test=2),
}


def load(config: str):
    return G.load(f"trained_{config}")


def state_dict(z, shapes):
    """Full state dict: trained tensors from the fixture, the rest `synth.synth_state_dict(mode='refinit', seed=1234)`."""
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234, mode="refinit").items()}
    n = 0
    for k in z.files:
        if k.startswith("w."):
            assert k[2:] in sd and tuple(z[k].shape) == tuple(shapes[k[2:]]), k
            sd[k[2:]] = torch.from_numpy(z[k].astype(np.float32))
            n += 1
    assert n >= 40, "fixture holds no trained tensors"
    return sd


def test_inputs(config: str):
    c = CONFIGS[config]
    xs, ys = synth.synth_seg_batch(c["n_test"], c["img_size"], c["num_classes"], seed=10_000)
    return torch.from_numpy(xs).repeat(1, 3, 1, 1), ys


def logits_rows(logits: torch.Tensor) -> np.ndarray:
    """NCHW logits -> (pixels, classes) rows, the layout the fixture's `logits` group was packed in."""
    return logits.permute(0, 2, 3, 1).reshape(-1, logits.shape[1]).detach().double().cpu().numpy()
