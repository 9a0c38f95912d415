"""GPU tests of the persistent stage kernel (csrc/stage_tc.cu, cswin_stage_fwd through the C ABI): all CSWinBlocks of a stage
(networks/cswin_unet.py:160-181 inside the stage loops :462-478 / :505-533) in one dataflow launch, against the composed
native path (5 launches per block, same tiles) and against the CPU oracle in fp64; repeated launches (the kernel must leave
its counters zeroed), CUDA-graph replay, and the whole model with the stage kernel on every stage."""
import pytest
import torch

import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, ops, synth
from oracle import cswin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (C, reso, heads, split, last_stage, B, n_blocks): the four stage shapes of cswin_tiny_224 (ragged last row tile: M % 128 != 0
# for B = 3 at stage 3, B = 5 at stage 4), a 10-block stage (two launches: > max_blocks), one-block stages, odd problem counts
STAGE_CASES = [(256, 14, 8, 7, False, 3, 3), (256, 14, 8, 7, False, 24, 9), (64, 56, 2, 1, False, 2, 1), (128, 28, 4, 2, False, 3, 2),
               (512, 7, 16, 7, True, 5, 1), (128, 8, 4, 8, True, 3, 2), (64, 8, 2, 2, False, 1, 10), (192, 14, 6, 7, False, 2, 2)]


def _blocks(C, reso, heads, split, last, n, tag="stage"):
    out, sds = [], []
    for i in range(n):
        blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).eval()
        sd = {k: torch.from_numpy(synth.synth_tensor(f"{tag}/{C}/{reso}/{i}/" + k, tuple(v.shape), 21)) for k, v in blk.state_dict().items()}
        for k in sd:                                      # LayerNorm affine away from (1, 0) so that the fold is exercised
            if k.endswith("norm1.weight") or k.endswith("norm2.weight"):
                sd[k] = 1.0 + 0.2 * sd[k] / sd[k].abs().max().clamp_min(1e-6)
        blk.load_state_dict(sd, strict=True)
        out.append(blk.to(DEV))
        sds.append(sd)
    return out, sds


def _composed(blocks, x):
    with torch.no_grad():
        for b in blocks:
            x = b(x)
    return x


@pytest.mark.parametrize("C,reso,heads,split,last,B,n", STAGE_CASES)
def test_stage_kernel_equals_composed_and_oracle(C, reso, heads, split, last, B, n, monkeypatch):
    blocks, sds = _blocks(C, reso, heads, split, last, n)
    L = reso * reso
    x = torch.from_numpy(synth.synth_tensor(f"stage_in/{C}/{reso}", (B, L, C), 22)).bfloat16()
    monkeypatch.setattr(modules, "STAGE_EXEC_DIMS", ())
    yc = _composed(blocks, x.to(DEV)).float().cpu()
    monkeypatch.setattr(modules, "STAGE_EXEC_DIMS", (C,))
    n0, t0 = cw.launch_count(), cw.tc_launch_count()
    with torch.no_grad():
        ys = modules.run_stage(blocks, x.to(DEV).clone())
    launches = cw.launch_count() - n0
    assert launches == 1 + (n + 8) // 9, launches            # row_stats + one persistent launch per <= 9 blocks
    ys = ys.float().cpu()
    assert torch.isfinite(ys).all()
    d = (ys - yc).abs().max().item()
    # fp64 oracle on the same bf16 input (fp32 weights as loaded)
    ref = x.double()
    for sd in sds:
        ref = O.cswin_block({k: v.double() for k, v in sd.items()}, "", ref, reso, heads, split, last)
    es, ec = (ys.double() - ref).abs().max().item(), (yc.double() - ref).abs().max().item()
    print(f"[stage C={C} reso={reso} B={B} blocks={n}] stage vs composed {d:.3e}; vs fp64 oracle: stage {es:.3e}, composed {ec:.3e}")
    assert es <= 1.25 * ec + 5e-3, (es, ec)
    assert d <= 2e-2 * max(1.0, ref.abs().max().item())


def test_stage_kernel_repeated_and_graph_replay():
    """The kernel leaves its counters zeroed: repeated eager launches and CUDA-graph replays give identical results."""
    C, reso, heads, split, B, n = 256, 14, 8, 7, 24, 4
    blocks, _ = _blocks(C, reso, heads, split, False, n, tag="stage_rep")
    x = torch.from_numpy(synth.synth_tensor("stage_rep_in", (B, reso * reso, C), 23)).bfloat16().to(DEV)
    old = modules.STAGE_EXEC_DIMS
    modules.STAGE_EXEC_DIMS = (C,)
    try:
        with torch.no_grad():
            y0 = modules.run_stage(blocks, x.clone()).clone()
            for _ in range(3):
                y = modules.run_stage(blocks, x.clone())
                assert torch.equal(y, y0)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                xin = x.clone()
                modules.run_stage(blocks, xin.clone())                # warm-up on the capture stream (workspace allocation)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    buf = xin.clone()
                    yg = modules.run_stage(blocks, buf)
            for _ in range(3):
                g.replay()
                torch.cuda.synchronize()
                assert torch.equal(yg, y0)
    finally:
        modules.STAGE_EXEC_DIMS = old


def test_model_with_stage_kernel_on_every_stage(monkeypatch):
    """Whole cswin_tiny_224 forward (bf16) with every stage on the persistent kernel vs the composed path: same logits up to
    bf16 rounding of the row statistics' summation order, identical arg-max on decisive pixels."""
    m = cw.cswin_tiny_224(num_classes=9).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    m = m.to(DEV)
    m.compute_dtype = torch.bfloat16
    x = torch.from_numpy(synth.synth_image_batch(4, 3, 224, seed=0, kind="ct")).to(DEV)
    outs = {}
    for dims in ((), (64, 128, 256, 512)):
        monkeypatch.setattr(modules, "STAGE_EXEC_DIMS", dims)
        n0 = cw.launch_count()
        with torch.no_grad():
            outs[dims] = m(x).float().cpu()
        outs[(dims, "n")] = cw.launch_count() - n0
    a, b = outs[()], outs[(64, 128, 256, 512)]
    d = (a - b).abs().max().item()
    top2 = a.topk(2, dim=1).values
    decisive = (top2[:, 0] - top2[:, 1]) > 0.05
    agree = (a.argmax(1) == b.argmax(1))[decisive].float().mean().item()
    print(f"[model, stage kernel everywhere] launches {outs[((64, 128, 256, 512), 'n')]} vs composed {outs[((), 'n')]}; logits max-abs diff {d:.3e}; "
          f"arg-max agreement on decisive pixels {agree:.6f}")
    assert outs[((64, 128, 256, 512), "n")] < outs[((), "n")] - 80
    assert d <= 6e-2 and agree >= 0.9999, (d, agree)       # |logit| ~ 25: a few bf16 ulps of difference in the folded-LN row statistics
