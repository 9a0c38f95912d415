import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cswin_unet_b200 as cw
from cswin_unet_b200 import synth
from cswin_unet_b200.train import seg_loss
B = 4
m = cw.cswin_tiny_224(num_classes=9).train()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
m = m.cuda(); m.compute_dtype = torch.bfloat16
x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).cuda()
y = torch.from_numpy(synth.synth_labels(B, 224, 9, seed=0)).cuda()
opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
def fwd(): return m(x)
def fwd_loss(): return seg_loss(m(x), y, 9)
def fwd_bwd():
    l = seg_loss(m(x), y, 9); l.backward(); return l
def full():
    l = fwd_bwd(); opt.step(); return l
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        opt.zero_grad(set_to_none=True); full()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for mode in ("global", "thread_local"):
    for name, fn in (("fwd", fwd), ("fwd_loss", fwd_loss), ("fwd_bwd", fwd_bwd), ("full", full)):
        opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, capture_error_mode=mode):
                fn()
            g.replay(); torch.cuda.synchronize()
            print(mode, name, "OK")
        except Exception as e:
            print(mode, name, "FAILED:", str(e).splitlines()[0][:150])
            torch.cuda.synchronize()
