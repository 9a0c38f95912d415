"""GPU tests of the fused block kernels (through the C ABI): csrc/qkv_attn_tc.cu ([LayerNorm -> qkv Linear -> LePE attention of
both branches] in one launch; cswin_unet.py:168-176 with :82-109) against the composed native path (cswin_linear_fwd +
cswin_lepe_attention_fwd), against the CPU oracle in fp64, and through CSWinBlock / the whole model."""
import numpy as np
import pytest
import torch

import cswin_unet_b200 as cw
from cswin_unet_b200 import modules, ops, synth
from oracle import cswin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (C, reso, heads, split, last_stage, B): the three two-branch stages of cswin_tiny_224 (batch sizes that leave ragged window
# tiles), a one-branch 8x8 window (two-windows-per-tile mode with ONE window: np < slots), a 6-head block (head groups of 2 over 3
# heads per branch is NOT supported -> must fall back), an odd number of windows per branch
FUSED_CASES = [(64, 56, 2, 1, False, 2), (128, 28, 4, 2, False, 3), (256, 14, 8, 7, False, 5), (256, 14, 8, 7, False, 24),
               (128, 8, 4, 8, True, 3), (128, 12, 4, 4, False, 2), (192, 14, 6, 7, False, 2), (64, 8, 2, 8, True, 4)]


def _block(C, reso, heads, split, last):
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).eval()
    sd = {k: torch.from_numpy(synth.synth_tensor(f"fused/{C}/{reso}/" + k, tuple(v.shape), 11)) for k, v in blk.state_dict().items()}
    blk.load_state_dict(sd, strict=True)
    return blk.to(DEV), sd


def _branch_args(descs):
    return [dict(conv_w=d["conv_w"], conv_b=d["conv_b"], heads=d["heads"], H_sp=d["H_sp"], W_sp=d["W_sp"]) for d in descs]


@pytest.mark.parametrize("C,reso,heads,split,last,B", FUSED_CASES)
def test_qkv_attention_fused_equals_composed_and_oracle(C, reso, heads, split, last, B):
    blk, sd = _block(C, reso, heads, split, last)
    L = reso * reso
    x = torch.from_numpy(synth.synth_tensor(f"fused_in/{C}/{reso}", (B, L, C), 12)).bfloat16()
    xd = x.to(DEV)
    brs = [(a.num_heads, a.H_sp, a.W_sp) for a in blk.attns]
    supported = ops.qkv_attention_supported(C, reso, brs)
    assert supported == (C != 192), (C, supported)          # 3 heads per branch: no head groups of 2
    dt = torch.bfloat16
    st = ops.row_stats(xd)
    wq, csq, bq = blk._folded("qkv", blk.qkv, blk.norm1)
    # composed: folded-LN Linear -> (B, L, 3C) -> attention kernel
    qkv = ops.linear(xd, wq, None, ln_fold=(st, csq, blk.norm1.eps), bias_f32=bq)
    att_c = torch.empty((B, L, C), dtype=dt, device=DEV)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    descs, off = [], 0
    for a in blk.attns:
        sl = slice(off, off + a.dim)
        descs.append(a.branch_desc(q[..., sl], k[..., sl], v[..., sl], att_c[..., sl]))
        off += a.dim
    ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), dt)
    if not supported:
        with pytest.raises(cw.CswinError):
            ops.qkv_lepe_attention(xd, wq, bq, (st, csq, blk.norm1.eps), _branch_args(descs), reso, float(blk.attns[0].scale))
        return
    n0 = cw.tc_launch_count()
    att_f = ops.qkv_lepe_attention(xd, wq, bq, (st, csq, blk.norm1.eps), _branch_args(descs), reso, float(blk.attns[0].scale))
    assert cw.tc_launch_count() == n0 + 1
    # fp64 truth from the oracle's pieces on the same bf16 input
    xs = x.double()
    u = O._ln(xs, sd["norm1.weight"].double(), sd["norm1.bias"].double(), 1e-5)
    z = u @ sd["qkv.weight"].double().T + sd["qkv.bias"].double()
    ref = torch.empty(B, L, C, dtype=torch.float64)
    off = 0
    for i, a in enumerate(blk.attns):
        sl = slice(off, off + a.dim)
        ref[..., sl] = O.lepe_attention(z[..., :C][..., sl], z[..., C:2 * C][..., sl], z[..., 2 * C:][..., sl],
                                        sd[f"attns.{i}.get_v.weight"].double(), sd[f"attns.{i}.get_v.bias"].double(), reso,
                                        a.idx, a.split_size, a.num_heads, float(a.scale))
        off += a.dim
    ef = (att_f.float().cpu().double() - ref).abs().max().item()
    ec = (att_c.float().cpu().double() - ref).abs().max().item()
    d = (att_f.float() - att_c.float()).abs().max().item()
    print(f"[qkv+attn fused C={C} reso={reso} B={B}] max-abs vs fp64: fused {ef:.3e}, composed {ec:.3e}; fused vs composed {d:.3e}")
    assert torch.isfinite(att_f.float()).all()
    assert ef <= 1.25 * ec + 4e-3, (ef, ec)                   # not less accurate than the composed bf16 path
    assert d <= 3e-2 * max(1.0, ref.abs().max().item())


def test_qkv_attention_fused_without_layernorm_fold():
    """ln_stats = NULL: x is consumed as it is (caller normalised it) and the bias is the plain qkv bias."""
    C, reso, heads, split, B = 128, 28, 4, 2, 2
    blk, sd = _block(C, reso, heads, split, False)
    x = torch.from_numpy(synth.synth_tensor("fused_in/plain", (B, reso * reso, C), 13)).bfloat16().to(DEV)
    w = sd["qkv.weight"].bfloat16().to(DEV)
    bias = sd["qkv.bias"].float().to(DEV)
    brs = [dict(conv_w=a.get_v.weight.detach().bfloat16(), conv_b=a.get_v.bias.detach().bfloat16(), heads=a.num_heads, H_sp=a.H_sp, W_sp=a.W_sp)
           for a in blk.attns]
    att_f = ops.qkv_lepe_attention(x, w, bias, None, brs, reso, float(blk.attns[0].scale))
    qkv = ops.linear(x, w, sd["qkv.bias"].bfloat16().to(DEV))
    att_c = torch.empty_like(att_f)
    descs = []
    for i, a in enumerate(blk.attns):
        sl = slice(i * 64, (i + 1) * 64)
        descs.append(dict(q=qkv[..., :C][..., sl], k=qkv[..., C:2 * C][..., sl], v=qkv[..., 2 * C:][..., sl], out=att_c[..., sl],
                          conv_w=brs[i]["conv_w"], conv_b=brs[i]["conv_b"], heads=a.num_heads, H_sp=a.H_sp, W_sp=a.W_sp))
    ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), torch.bfloat16)
    d = (att_f.float() - att_c.float()).abs().max().item()
    print(f"[qkv+attn fused, no fold] fused vs composed {d:.3e}")
    assert d <= 3e-2


@pytest.mark.parametrize("C,reso,heads,split,last,B", [(64, 56, 2, 1, False, 2), (128, 28, 4, 2, False, 2), (256, 14, 8, 7, False, 3)])
def test_block_with_fused_kernels_matches_oracle(C, reso, heads, split, last, B, monkeypatch):
    blk, sd = _block(C, reso, heads, split, last)
    x = torch.from_numpy(synth.synth_tensor(f"fusedblk_in/{C}", (B, reso * reso, C), 14)).bfloat16()
    ref = O.cswin_block({k: v.double() for k, v in sd.items()}, "", x.double(), reso, heads, split, last)
    outs = {}
    for fused in (True, False):
        monkeypatch.setattr(modules, "FUSE_QKV_ATTN", fused)
        n0 = cw.launch_count()
        with torch.no_grad():
            outs[fused] = blk(x.to(DEV)).float().cpu().double()
        outs[(fused, "n")] = cw.launch_count() - n0
    ef = (outs[True] - ref).abs().max().item()
    ec = (outs[False] - ref).abs().max().item()
    print(f"[block C={C}] launches fused {outs[(True, 'n')]} vs composed {outs[(False, 'n')]}; max-abs vs fp64 fused {ef:.3e} composed {ec:.3e}")
    assert outs[(True, "n")] < outs[(False, "n")]
    assert ef <= 1.25 * ec + 5e-3


@pytest.mark.parametrize("B,S,in_dtype", [(2, 224, torch.float32), (3, 224, torch.bfloat16), (1, 512, torch.float32), (5, 172, torch.float32), (2, 96, torch.float32)])
def test_stem_fused_equals_composed_and_oracle(B, S, in_dtype, monkeypatch):
    """csrc/stem_tc.cu: Conv2d(3, 64, 7, 4, 2) + token layout + LayerNorm(64) (cswin_unet.py:338-342) as one implicit-GEMM launch,
    against the composed path (im2col + Linear + LayerNorm) and against torch in fp64; ragged last tile (B = 3, 5), tiles that
    span two images, a 512^2 input (128 pixels per output row) and small images (43 and 24 pixels per output row)."""
    from cswin_unet_b200 import model as model_mod
    m = cw.cswin_tiny_224(num_classes=9).eval()                           # (the stem does not depend on img_size)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=5).items()}
    sd["stage1_conv_embed.2.weight"] = 1.0 + 0.3 * torch.from_numpy(synth.synth_tensor("stem/g", (64,), 5))
    sd["stage1_conv_embed.2.bias"] = 0.2 * torch.from_numpy(synth.synth_tensor("stem/b", (64,), 6))
    sd["stage1_conv_embed.0.bias"] = 0.1 * torch.from_numpy(synth.synth_tensor("stem/cb", (64,), 7))
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    x = torch.from_numpy(synth.synth_image_batch(B, 3, S, seed=3, kind="randn")).to(in_dtype).to(DEV)
    outs = {}
    for fused in (True, False):
        monkeypatch.setattr(model_mod, "FUSE_STEM", fused)
        n0 = cw.launch_count()
        with torch.no_grad():
            y = m._stem(x, torch.bfloat16)
        outs[fused] = (y.float().cpu(), y._cswin_stats.cpu(), cw.launch_count() - n0)
    assert outs[True][2] == 1 and outs[False][2] == 3, (outs[True][2], outs[False][2])
    xr = x.float().cpu().double()
    ref = torch.nn.functional.conv2d(xr, sd["stage1_conv_embed.0.weight"].bfloat16().double(), sd["stage1_conv_embed.0.bias"].double(), stride=4, padding=2)
    ref = ref.flatten(2).transpose(1, 2)
    ref = torch.nn.functional.layer_norm(ref, (64,), sd["stage1_conv_embed.2.weight"].double(), sd["stage1_conv_embed.2.bias"].double(), 1e-5)
    ef, ec = (outs[True][0].double() - ref).abs().max().item(), (outs[False][0].double() - ref).abs().max().item()
    d = (outs[True][0] - outs[False][0]).abs().max().item()
    # the statistics side channel must describe the bf16 rows that were written
    yb = outs[True][0].double()
    st = outs[True][1].double().view(-1, 2)
    es = max((st[:, 0] - yb.sum(-1).flatten()).abs().max().item(), (st[:, 1] - (yb * yb).sum(-1).flatten()).abs().max().item())
    print(f"[stem fused B={B} S={S} {in_dtype}] max-abs vs fp64: fused {ef:.3e}, composed {ec:.3e}; fused vs composed {d:.3e}; stats error {es:.2e}")
    assert outs[True][0].shape == ref.shape and torch.isfinite(outs[True][0]).all()
    assert ef <= 1.25 * ec + 2e-3 and d <= 6e-2 and es <= 2e-3


# B, H (= W), C, N, stride, view: Merge_Block shapes of the 224^2 and 512^2 configurations (stride 2), CARAFE.encoder shapes
# (stride 1), M tiles of whole images (7x7 outputs: two images per tile, odd B leaves a ragged tile), of 7 / 4 / 2 output rows,
# full 128-row tiles (TMA-store epilogue), an input that is a column view of a wider buffer (CARAFE's [down | z] Linear)
CONV_CASES = [(3, 56, 64, 128, 2, False), (2, 28, 128, 256, 2, False), (5, 14, 256, 512, 2, False), (24, 14, 256, 512, 2, False),
              (3, 14, 64, 36, 1, True), (5, 7, 128, 36, 1, True), (1, 128, 64, 128, 2, False), (2, 64, 128, 256, 2, False),
              (2, 32, 256, 512, 2, False), (1, 20, 64, 72, 1, False)]


@pytest.mark.parametrize("B,H,C,N,stride,view", CONV_CASES)
def test_conv_tokens_implicit_gemm_equals_im2col_path_and_fp64(B, H, C, N, stride, view):
    """cswin_conv_tokens_fwd (gemm_tc.cu, implicit GEMM: the A operand is fetched as strided 4-D TMA boxes of the token image) for
    Merge_Block.conv (cswin_unet.py:214-216) and CARAFE.encoder (:240-241): bit-identical to cswin_im2col_tokens + cswin_linear_fwd
    (same MMA order) and within bf16 rounding of torch's fp64 conv2d."""
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + C + N)
    xw = torch.randn(B, H * H, C + (64 if view else 0), generator=g).bfloat16().to(DEV)
    x = xw[..., :C]
    w4 = (torch.randn(N, C, 3, 3, generator=g) / (9 * C) ** 0.5).bfloat16()
    bias = (0.1 * torch.randn(N, generator=g)).bfloat16()
    wk = w4.permute(0, 2, 3, 1).reshape(N, -1).contiguous().to(DEV)
    n0, t0 = cw.launch_count(), cw.tc_launch_count()
    y = ops.conv_tokens(x, H, H, wk, bias.to(DEV), 3, 3, stride, 1)
    assert y is not None and cw.launch_count() == n0 + 1 and cw.tc_launch_count() == t0 + 1      # one tcgen05 launch, no gather
    col = ops.im2col_tokens(x, H, H, 3, 3, stride, 1)
    y2 = ops.linear(col, wk, bias.to(DEV))
    OH = (H + 2 - 3) // stride + 1
    assert y.shape == y2.shape == (B * OH * OH, N)
    assert torch.equal(y, y2), (y.float() - y2.float()).abs().max().item()
    xi = x.float().cpu().double().view(B, H, H, C).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(xi, w4.double(), bias.double(), stride=stride, padding=1).permute(0, 2, 3, 1).reshape(B * OH * OH, N)
    err = (y.float().cpu().double() - ref).abs().max().item()
    assert err <= 3e-2, err


def test_conv_tokens_outside_the_envelope_is_declined_not_faked():
    x = torch.randn(2, 28 * 28, 32, device=DEV).bfloat16()                      # 32 channels: a 64-wide K block would span two taps
    wk = torch.randn(36, 9 * 32, device=DEV).bfloat16()
    n0 = cw.launch_count()
    assert ops.conv_tokens(x, 28, 28, wk, None, 3, 3, 1, 1) is None and cw.launch_count() == n0
    assert ops.conv_tokens(x.float(), 28, 28, wk.float(), None, 3, 3, 1, 1) is None  # fp32: the SIMT path composes im2col + Linear


def test_merge_block_and_carafe_use_the_implicit_conv(monkeypatch):
    """Merge_Block / CARAFE forward with the implicit-GEMM conv vs the im2col path: identical outputs, fewer launches."""
    for mod, reso in ((cw.Merge_Block(64, 128), 56), (cw.Merge_Block(256, 512), 14), (cw.CARAFE(256, 128), 14), (cw.CARAFE(512, 256), 7)):
        mod = mod.to(DEV).eval()
        x = torch.randn(3, reso * reso, mod.conv.in_channels if hasattr(mod, "conv") else mod.down.in_channels, device=DEV).bfloat16()
        res = {}
        for on in (True, False):
            monkeypatch.setattr(modules, "IMPLICIT_CONV", on)
            n0 = cw.launch_count()
            with torch.no_grad():
                res[on] = (mod(x), cw.launch_count() - n0)
        assert res[True][1] == res[False][1] - 1
        assert torch.equal(res[True][0], res[False][0])


def test_carafe_with_few_compressed_channels_pads_them_for_the_implicit_conv(monkeypatch):
    """C/4 = 16 / 32 compressed channels (CARAFE4 at 56^2, CARAFE at 28^2): `down` gets zero rows up to 64 outputs and the encoder zero
    input channels, so the 3x3 encoder is an implicit GEMM too.  Same values up to the accumulation order (the zero channels add
    exact zeros), one launch fewer than the im2col path."""
    for mod, reso in ((cw.CARAFE(128, 64), 28), (cw.CARAFE4(64, 64), 56)):
        mod = mod.to(DEV).eval()
        sd = {k: torch.from_numpy(synth.synth_tensor(f"carafe_pad/{reso}/" + k, tuple(v.shape), 3)) for k, v in mod.state_dict().items()}
        mod.load_state_dict(sd, strict=True)
        x = torch.randn(2, reso * reso, mod.down.in_channels, device=DEV).bfloat16()
        res = {}
        for on in (True, False):
            monkeypatch.setattr(modules, "IMPLICIT_CONV", on)
            n0 = cw.launch_count()
            with torch.no_grad():
                res[on] = (mod(x).float(), cw.launch_count() - n0)
        assert res[True][1] == res[False][1] - 1
        d = (res[True][0] - res[False][0]).abs().max().item()
        assert d <= 2e-2 * max(1.0, res[False][0].abs().max().item()), d
