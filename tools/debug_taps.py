"""Per-stage error localisation: native model (GPU) vs oracle (CPU fp64) on the same synthetic weights/input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
from oracle import cswin_oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "refinit"
kind = sys.argv[2] if len(sys.argv) > 2 else "ct"
dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
m = cw.cswin_tiny_224(num_classes=9).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
sdn = synth.synth_state_dict(shapes, seed=1234, mode=mode)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sdn.items()}, strict=True)
m = m.cuda(); m.compute_dtype = dtype
sd = {k: torch.from_numpy(v).double() for k, v in sdn.items()}
x = torch.from_numpy(synth.synth_image_batch(2, 3, 224, seed=0, kind=kind))
taps = {}
with torch.no_grad():
    ref = O.cswin_unet_forward(sd, x.double(), taps=taps)
got = {}
def hook(name):
    def f(mod, inp, out): got[name] = out.detach().double().cpu()
    return f
for n in ("stage1", "stage2", "stage3", "stage4", "stage_up4", "stage_up3", "stage_up2", "stage_up1"):
    getattr(m, n)[-1].register_forward_hook(hook(n))
for n in ("merge1", "merge2", "merge3", "upsample4", "upsample3", "upsample2"):
    getattr(m, n).register_forward_hook(hook(n))
with torch.no_grad():
    y = m(x.cuda()).double().cpu()
for n in ("stage1", "stage2", "stage3", "stage4", "stage_up4", "stage_up3", "stage_up2", "stage_up1"):
    e = (got[n] - taps[n]).abs()
    print(f"{n:10s} max-abs {e.max():.3e}  ref absmax {taps[n].abs().max():.3e}  argmax idx {np.unravel_index(int(e.argmax()), e.shape)}")
e = (y - ref).abs()
print(f"logits     max-abs {e.max():.3e} at {np.unravel_index(int(e.argmax()), e.shape)} ref absmax {ref.abs().max():.3e}")
print("pixels with err>1e-4:", int((e.amax(1) > 1e-4).sum()), "of", e.shape[0]*e.shape[2]*e.shape[3])
bad = (e.amax(1) > 1e-4).nonzero()
print(bad[:20].tolist())
