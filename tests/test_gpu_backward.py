"""GPU tests of the native backward kernels (through the C ABI via cswin_unet_b200.autograd).  Truth = the reference's own
autograd where golden vectors exist (LePEAttention: tests/golden/lepe_*.npz) and torch.autograd through the fp64 CPU oracle
elsewhere.  fp32 path: relative L2 error per gradient tensor <= 1e-4; bf16 path: cosine similarity with the fp64 gradient."""
import numpy as np
import pytest
import torch

import cswin_unet_b200 as cw
from cswin_unet_b200 import autograd as ag
from cswin_unet_b200 import synth
from oracle import cswin_oracle as O
from tests import golden_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda"
LEPE_EXTRA = ((32, 16, 0, 2, 1), (64, 16, 1, 2, 2), (64, 8, -1, 8, 2), (128, 16, 0, 8, 4), (48, 12, 1, 3, 3))
# BASELINE configs[4] (512^2) windows with golden vectors from the unmodified reference (tests/golden/lepe_wide.npz)
LEPE_WIDE_GOLD = ((128, 32, 0, 8, 4), (128, 32, 1, 8, 4), (512, 16, -1, 8, 16), (64, 24, 0, 8, 2), (64, 14, -1, 14, 2))


def T(a, dtype=torch.float32, device=DEV):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a = a.detach().double().cpu().reshape(-1); b = b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("tag,cfgs,B", [("t224", synth.LEPE_CONFIGS_T224, 1), ("extra", LEPE_EXTRA, 2), ("wide", LEPE_WIDE_GOLD, 1)])
def test_lepe_attention_backward_fp32_vs_reference_autograd(tag, cfgs, B):
    z = G.load(f"lepe_{tag}")
    for (cb, reso, idx, split, heads) in cfgs:
        full_c = cb if idx == -1 else 2 * cb
        base = T(synth.synth_qkv(B, reso, full_c, seed=0)).requires_grad_(True)
        off = cb if idx == 1 else 0
        qkv = base.permute(2, 0, 1, 3)[..., off:off + cb]
        m = cw.LePEAttention(cb, resolution=reso, idx=idx, split_size=split, num_heads=heads).to(DEV)
        with torch.no_grad():
            m.get_v.weight.copy_(T(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.weight", (cb, 1, 3, 3), 1)))
            m.get_v.bias.copy_(T(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.bias", (cb,), 1)))
        y = m(qkv)
        key = f"c{cb}_r{reso}_i{idx}_s{split}_h{heads}"
        gup = T(synth.synth_tensor(f"lepe_grad/{key}", tuple(y.shape), 2))
        gb, gw, gbias = torch.autograd.grad(y, [base, m.get_v.weight, m.get_v.bias], gup)
        dqkv = gb[..., off:off + cb].permute(2, 0, 1, 3).reshape(3 * B, reso * reso, cb)
        G.compare(z, key + "_dqkv", dqkv.cpu().numpy(), atol=2e-5, rtol=1e-4)
        G.compare(z, key + "_dw", gw.reshape(cb, 9).cpu().numpy(), atol=1e-3, rtol=1e-4)
        G.compare(z, key + "_db", gbias.reshape(1, cb).cpu().numpy(), atol=1e-3, rtol=1e-4)


def test_elementary_backward_ops_fp32_vs_oracle_autograd():
    g = torch.Generator().manual_seed(5)
    M, C, N = 203, 96, 72
    x = torch.randn(M, C, generator=g, dtype=torch.float64)
    gam = 1 + 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
    bet = 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
    w = torch.randn(N, C + 40, generator=g, dtype=torch.float64) / 10
    b = 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    a2 = torch.randn(M, 40, generator=g, dtype=torch.float64)
    res = torch.randn(M, N, generator=g, dtype=torch.float64)
    ss = (torch.bernoulli(torch.full((7,), 0.7), generator=g) / 0.7).double()
    up = torch.randn(M, N, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (x, gam, bet, w, b, a2, res)]
    xo, go, bo, wo, bbo, a2o, ro = leaves
    u = O._ln(xo, go, bo, 1e-5)
    zz = torch.cat([u, a2o], -1) @ wo.T + bbo
    out = ro + ss.repeat_interleave(29).view(M, 1) * O._gelu(zz)
    ref = torch.autograd.grad(out, leaves, up)
    # native: LN -> Linear (two sources) -> GELU -> scaled residual add expressed with the fused Functions
    dl = [t.detach().float().to(DEV).requires_grad_(True) for t in (x, gam, bet, w, b, a2, res)]
    xn, gn, bn, wn, bbn, a2n, rn = dl
    un = ag.LayerNormFn.apply(xn, gn, bn, 1e-5)
    zn = ag.linear(un, wn, bbn, a2=a2n)
    hn = ag.GeluFn.apply(zn)
    eye = torch.eye(N, device=DEV)
    outn = ag.linear(hn, eye, None, residual=rn, sample_scale=ss.float().to(DEV), rps=29)      # res + s * h
    got = torch.autograd.grad(outn, dl, up.float().to(DEV))
    for name, a_, b_ in zip(("x", "gamma", "beta", "w", "b", "a2", "res"), got, ref):
        assert rel(a_, b_) <= 1e-4, (name, rel(a_, b_))


BLOCKS = ((64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True))


@pytest.mark.parametrize("dim,reso,heads,split,last", BLOCKS)
def test_block_backward_fp32_vs_oracle_autograd(dim, reso, heads, split, last):
    m = cw.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).train()
    sd64 = {k: torch.from_numpy(synth.synth_tensor(f"block/{dim}/" + k, tuple(v.shape), 3)).double() for k, v in m.state_dict().items()}
    m.load_state_dict({k: v.float() for k, v in sd64.items()}, strict=True)
    B = 2
    x64 = torch.from_numpy(synth.synth_tensor(f"block_in/{dim}", (B, reso * reso, dim), 4)).double()
    gup = torch.from_numpy(synth.synth_tensor(f"block_gup/{dim}", (B, reso * reso, dim), 6)).double()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
    xo = x64.clone().requires_grad_(True)
    yo = O.cswin_block(leaves, "", xo, reso, heads, split, last)
    ref = torch.autograd.grad(yo, [xo] + list(leaves.values()), gup)
    xn = x64.float().to(DEV).requires_grad_(True)
    yn = m(xn)
    assert rel(yn, yo) <= 1e-5
    got = torch.autograd.grad(yn, [xn] + [dict(m.named_parameters())[k] for k in leaves], gup.float().to(DEV))
    for name, a_, b_ in zip(["x"] + list(leaves), got, ref):
        assert rel(a_, b_) <= 1e-4, (name, rel(a_, b_))


def test_merge_and_carafe_backward_fp32_vs_oracle_autograd():
    for kind, mk, dim, reso in (("merge", lambda: cw.Merge_Block(64, 128), 64, 28), ("carafe", lambda: cw.CARAFE(128, 64), 128, 14),
                                ("carafe4", lambda: cw.CARAFE4(64, 64), 64, 14)):
        m = mk().to(DEV).train()
        sd64 = {k: torch.from_numpy(synth.synth_tensor(f"bw/{kind}/" + k, tuple(v.shape), 5)).double() for k, v in m.state_dict().items()}
        m.load_state_dict({k: v.float() for k, v in sd64.items()}, strict=True)
        x64 = torch.from_numpy(synth.synth_tensor(f"bw/{kind}/x", (2, reso * reso, dim), 6)).double()
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
        xo = x64.clone().requires_grad_(True)
        yo = O.merge_block(leaves, "", xo) if kind == "merge" else O.carafe(leaves, "", xo, 2 if kind == "carafe" else 4)
        gup = torch.from_numpy(synth.synth_tensor(f"bw/{kind}/g", tuple(yo.shape), 7)).double()
        ref = torch.autograd.grad(yo, [xo] + list(leaves.values()), gup)
        xn = x64.float().to(DEV).requires_grad_(True)
        yn = m(xn)
        assert rel(yn, yo) <= 1e-5, kind
        got = torch.autograd.grad(yn, [xn] + [dict(m.named_parameters())[k] for k in leaves], gup.float().to(DEV))
        for name, a_, b_ in zip(["x"] + list(leaves), got, ref):
            assert rel(a_, b_) <= 1e-4, (kind, name, rel(a_, b_))


def _model_and_oracle_grads(compute_dtype):
    m = cw.cswin_tiny_224(num_classes=9, drop_path_rate=0.0).train()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sdn = synth.synth_state_dict(shapes, seed=1234)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sdn.items()}, strict=True)
    m = m.to(DEV)
    m.compute_dtype = compute_dtype
    x = torch.from_numpy(synth.synth_image_batch(2, 3, 224, seed=0, kind="ct"))
    y = torch.from_numpy(synth.synth_labels(2, 224, 9, seed=0))
    logits = m(x.to(DEV))
    loss = O.seg_loss(logits.float(), y.to(DEV), 9)          # trainer.py:55-57 (harness-level loss, plain torch)
    loss.backward()
    return m, sdn, x, y, float(loss)


def test_full_model_train_step_gradients_fp32_vs_oracle_autograd():
    m, sdn, x, y, loss = _model_and_oracle_grads(torch.float32)
    leaves = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in sdn.items()}
    lo = O.seg_loss(O.cswin_unet_forward(leaves, x.double()), y, 9)
    lo.backward()
    assert abs(loss - float(lo)) <= 1e-5
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        r = rel(p.grad, leaves[k].grad)
        if r > worst[1]:
            worst = (k, r)
        assert r <= 2e-3, (k, r)
    print(f"[train fp32] loss {loss:.6f} (oracle {float(lo):.6f}); worst relative gradient error {worst[1]:.2e} at {worst[0]}")


def test_full_model_train_step_gradients_bf16_vs_oracle_autograd():
    m, sdn, x, y, loss = _model_and_oracle_grads(torch.bfloat16)
    leaves = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in sdn.items()}
    lo = O.seg_loss(O.cswin_unet_forward(leaves, x.double()), y, 9)
    lo.backward()
    assert abs(loss - float(lo)) <= 2e-2
    cs = {k: cos(p.grad, leaves[k].grad) for k, p in m.named_parameters() if p.numel() >= 4096}
    worst = min(cs, key=cs.get)
    print(f"[train bf16] loss {loss:.5f} (oracle {float(lo):.5f}); min cosine(grad) over weight matrices {cs[worst]:.4f} at {worst}")
    assert cs[worst] >= 0.98
    assert all(p.grad.dtype == p.dtype and torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("M,N,K", [(3136, 192, 64), (784, 512, 128), (4704, 256, 1024), (1176, 36, 1152), (201, 64, 152), (98, 9, 64), (75264, 64, 256)])
def test_linear_wgrad_bf16_tensor_core_vs_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    dz = torch.randn(M, N, generator=g).bfloat16()
    a = torch.randn(M, K, generator=g).bfloat16()
    dw = torch.zeros(N, K, device=DEV)
    db = torch.zeros(N, device=DEV)
    t0 = cw.tc_launch_count()
    cw.ops.linear_wgrad(dz.to(DEV), a.to(DEV), dw, db)
    if (N * 2) % 16 == 0 and (K * 2) % 16 == 0:       # TMA needs 16-byte row pitches; other shapes use the SIMT kernel
        assert cw.tc_launch_count() == t0 + 1, "bf16 wgrad must run on the tcgen05 kernel"
    ref_w = dz.double().T @ a.double()
    ref_b = dz.double().sum(0)
    assert rel(dw, ref_w) <= 1e-5 and rel(db, ref_b) <= 1e-5, (rel(dw, ref_w), rel(db, ref_b))


@pytest.mark.parametrize("cfgs,B", [(synth.LEPE_CONFIGS_T224, 2), (LEPE_EXTRA, 2)])
def test_lepe_attention_backward_bf16_tensor_core_vs_fp64(cfgs, B):
    for (cb, reso, idx, split, heads) in cfgs:
        full_c = cb if idx == -1 else 2 * cb
        base64 = torch.from_numpy(synth.synth_qkv(B, reso, full_c, seed=0)).bfloat16().double()     # bf16-representable inputs
        off = cb if idx == 1 else 0
        w64 = torch.from_numpy(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.weight", (cb, 1, 3, 3), 1)).bfloat16().double()
        b64 = torch.from_numpy(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.bias", (cb,), 1)).bfloat16().double()
        key = f"c{cb}_r{reso}_i{idx}_s{split}_h{heads}"
        gup64 = torch.from_numpy(synth.synth_tensor(f"lepe_grad/{key}", (B, reso * reso, cb), 2)).bfloat16().double()
        # fp64 truth through the oracle
        leaves = [t.clone().requires_grad_(True) for t in (base64, w64, b64)]
        v = leaves[0].permute(2, 0, 1, 3)[..., off:off + cb]
        yo = O.lepe_attention(v[0], v[1], v[2], leaves[1], leaves[2], reso, idx, split, heads)
        ref = torch.autograd.grad(yo, leaves, gup64)
        # native bf16
        base = base64.to(DEV).bfloat16().requires_grad_(True)
        m = cw.LePEAttention(cb, resolution=reso, idx=idx, split_size=split, num_heads=heads).to(DEV)
        with torch.no_grad():
            m.get_v.weight.copy_(w64.float()); m.get_v.bias.copy_(b64.float())
        t0 = cw.tc_launch_count()
        y = m(base.permute(2, 0, 1, 3)[..., off:off + cb])
        gb, gw, gbias = torch.autograd.grad(y, [base, m.get_v.weight, m.get_v.bias], gup64.to(DEV).bfloat16())
        if cb // heads == 32 and m.H_sp * m.W_sp <= 128:
            assert cw.tc_launch_count() == t0 + 2, f"{key}: forward and backward must both run on tcgen05 kernels"
        sl = (Ellipsis, slice(off, off + cb))
        assert rel(gb[sl], ref[0][sl]) <= 2e-2, (key, "dqkv", rel(gb[sl], ref[0][sl]))
        assert rel(gw, ref[1]) <= 2e-2 and rel(gbias, ref[2]) <= 2e-2, (key, rel(gw, ref[1]), rel(gbias, ref[2]))


def test_native_sgd_matches_torch_optim_sgd():
    """cswin_sgd_momentum_step == torch.optim.SGD(lr, momentum .9, weight_decay 1e-4).step() (trainer.py:42, :61) over ragged,
    unaligned tensors for 4 steps with a changing learning rate, and leaves bf16(param) in the shadow copies."""
    from cswin_unet_b200 import ops
    g = torch.Generator().manual_seed(3)
    shapes = [(7,), (64, 3, 7, 7), (131072 + 5,), (192, 64), (9, 64, 1, 1), (1,)]
    ps = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.SGD(ref, lr=0.05, momentum=0.9, weight_decay=1e-4)
    mom = [torch.zeros_like(p) for p in ps]
    flat = torch.zeros(sum((p.numel() + 7) // 8 * 8 for p in ps), dtype=torch.bfloat16, device=DEV)
    sh, off = [], 0
    for p in ps:
        sh.append(flat[off:off + p.numel()].view(p.shape))
        off += (p.numel() + 7) // 8 * 8
    lr_dev = torch.tensor([0.05], device=DEV)
    for step in range(4):
        grads = [torch.randn(s, generator=g).to(DEV) for s in shapes]
        lr = 0.05 * (1 - step / 4) ** 0.9
        for gp in opt.param_groups:
            gp["lr"] = lr
        lr_dev.fill_(lr)
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        opt.step()
        tbl = ops.sgd_chunk_table(ps, grads, mom, sh).to(DEV)
        assert tbl.shape[0] == sum((p.numel() + ops.SGD_CHUNK - 1) // ops.SGD_CHUNK for p in ps)
        ops.sgd_momentum_step(tbl, lr_dev, 0.9, 1e-4)
        for p, r, s_ in zip(ps, ref, sh):
            assert (p - r.detach()).abs().max().item() <= 1e-6 * max(1.0, r.abs().max().item()), step
            assert torch.equal(s_, p.bfloat16())


def test_train_step_native_sgd_follows_torch_sgd_trajectory(monkeypatch):
    """TrainStep (CUDA-graph replay, native fused SGD, lr schedule through device memory) against the same TrainStep driven by
    torch.optim.SGD: identical loss trajectory over 5 steps within bf16 noise, and parameters that moved the same way."""
    x = torch.from_numpy(synth.synth_image_batch(2, 3, 224, seed=0, kind="ct")).to(DEV)
    y = torch.from_numpy(synth.synth_labels(2, 224, 9, seed=0)).to(DEV)
    out = {}
    for mode in ("native", "torch"):
        if mode == "torch":
            monkeypatch.setenv("CSWIN_TORCH_SGD", "1")
        torch.manual_seed(0)
        m = cw.cswin_tiny_224(num_classes=9, drop_path_rate=0.0)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
        m = m.to(DEV)
        step = cw.TrainStep(m, lr=0.05, warmup=1)
        assert step._native_sgd == (mode == "native")
        losses = []
        for i in range(5):
            step.lr = 0.05 * (1 - i / 5) ** 0.9
            losses.append(float(step(x, y)))
        out[mode] = (losses, torch.cat([p.detach().flatten() for p in m.parameters()]).clone())
    ln, lt = out["native"][0], out["torch"][0]
    print(f"[native sgd] losses {ln} vs torch {lt}")
    assert ln[-1] < ln[0]
    assert max(abs(a - b) for a, b in zip(ln, lt)) <= 3e-2
    assert cos(out["native"][1], out["torch"][1]) >= 0.9999


@pytest.mark.parametrize("nc,ldt", [(9, torch.int64), (9, torch.uint8), (4, torch.int32), (3, torch.int64), (2, torch.uint8)])
def test_native_seg_loss_matches_oracle(nc, ldt):
    """cswin_seg_loss_fwd/bwd == 0.4 CE + 0.6 DiceLoss(softmax=True) of trainer.py:55-57 / utils.py:9-45 (oracle, fp64 autograd),
    including classes that never occur in the labels and an upstream gradient != 1."""
    from cswin_unet_b200 import train as T_
    g = torch.Generator().manual_seed(nc)
    logits = (torch.randn(3, nc, 40, 56, generator=g) * 2.0)
    labels = torch.randint(0, max(nc - 1, 1), (3, 40, 56), generator=g)          # the last class never occurs (for nc > 1)
    lo = logits.double().requires_grad_(True)
    ref = O.seg_loss(lo, labels, nc)
    (ref * 1.7).backward()
    ln = logits.to(DEV).requires_grad_(True)
    got = T_.seg_loss(ln, labels.to(ldt).to(DEV), nc)
    (got * 1.7).backward()
    assert abs(float(got) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert rel(ln.grad, lo.grad) <= 2e-5


@pytest.mark.parametrize("M,K,N", [(4704, 256, 1024), (1176, 512, 2048), (300, 64, 256), (18816, 128, 512)])
def test_linear_training_epilogues_match_the_composed_kernels(M, K, N):
    """cswin_linear_fwd training epilogues (tcgen05 path): aux_out = pre-activation next to GELU(z) (Mlp fc1, cswin_unet.py:22-23),
    act 2 = data gradient x GELU'(z).  Checked against the separate kernels they replace and against fp64 math."""
    from cswin_unet_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    u = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w1 = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(DEV)
    b1 = (torch.randn(N, generator=g) * 0.1).bfloat16().to(DEV)
    z_ref = ops.linear(u, w1, b1)
    h_ref = ops.act_fwd(z_ref, act=1)
    z = torch.empty_like(z_ref)
    t0 = cw.tc_launch_count()
    h = ops.linear(u, w1, b1, act=1, aux_out=z)
    assert cw.tc_launch_count() == t0 + 1
    assert torch.equal(z, z_ref), "aux_out must be bit-identical to the plain Linear"
    z64 = u.double() @ w1.double().T + b1.double()
    h64 = torch.nn.functional.gelu(z64)
    assert (h.double() - h64).abs().max().item() <= 2e-2 and rel(h, h64) <= 6e-3
    assert rel(h, h_ref.double()) <= 6e-3
    # backward epilogue: dz = (dh_src @ w2) * GELU'(z)
    C2 = K
    dz2 = torch.randn(M, C2, generator=g).bfloat16().to(DEV)
    w2 = (torch.randn(C2, N, generator=g) * N ** -0.5).bfloat16().to(DEV)          # fc2.weight (C, hidden), read in place as (K', N')
    dh = ops.linear(dz2, w2, w_kn=True)
    dz_ref = ops.act_bwd(dh, z, None, 0, act=1)
    dz = ops.linear(dz2, w2, w_kn=True, act=2, residual=z)
    x = z.double()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * np.pi) ** 0.5
    dz64 = (dz2.double() @ w2.double()) * gp
    assert rel(dz, dz64) <= 6e-3, rel(dz, dz64)
    assert rel(dz, dz_ref.double()) <= 8e-3, rel(dz, dz_ref.double())


LEPE_WIDE_BWD = ((128, 32, 0, 8, 4), (128, 32, 1, 8, 4), (512, 16, -1, 8, 16), (64, 24, 0, 8, 2), (64, 14, -1, 14, 2))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_lepe_attention_backward_wide_windows_vs_fp64(dtype, tol):
    """BASELINE configs[4] (512^2): 256-, 192- and 196-token windows.  bf16: forward and backward on the wide tcgen05 kernels
    (two query tiles x two key halves per (window, head)); fp32: the general kernels.  Both with the get_v parameter
    gradients from cswin_lepe_param_grad on the second stream where it applies — compared with autograd through the fp64 oracle."""
    B = 2
    for (cb, reso, idx, split, heads) in LEPE_WIDE_BWD:
        full_c = cb if idx == -1 else 2 * cb
        base64 = torch.from_numpy(synth.synth_qkv(B, reso, full_c, seed=0)).to(dtype).double()
        off = cb if idx == 1 else 0
        w64 = torch.from_numpy(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.weight", (cb, 1, 3, 3), 1)).to(dtype).double()
        b64 = torch.from_numpy(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.bias", (cb,), 1)).to(dtype).double()
        key = f"c{cb}_r{reso}_i{idx}_s{split}_h{heads}"
        gup64 = torch.from_numpy(synth.synth_tensor(f"lepe_grad/{key}", (B, reso * reso, cb), 2)).to(dtype).double()
        leaves = [t.clone().requires_grad_(True) for t in (base64, w64, b64)]
        v = leaves[0].permute(2, 0, 1, 3)[..., off:off + cb]
        yo = O.lepe_attention(v[0], v[1], v[2], leaves[1], leaves[2], reso, idx, split, heads)
        ref = torch.autograd.grad(yo, leaves, gup64)
        base = base64.to(DEV).to(dtype).requires_grad_(True)
        m = cw.LePEAttention(cb, resolution=reso, idx=idx, split_size=split, num_heads=heads).to(DEV)
        with torch.no_grad():
            m.get_v.weight.copy_(w64.float()); m.get_v.bias.copy_(b64.float())
        t0 = cw.tc_launch_count()
        y = m(base.permute(2, 0, 1, 3)[..., off:off + cb])
        assert rel(y, yo) <= tol, (key, "forward", rel(y, yo))
        gb, gw, gbias = torch.autograd.grad(y, [base, m.get_v.weight, m.get_v.bias], gup64.to(DEV).to(dtype))
        if dtype == torch.bfloat16:           # forward: wide tcgen05 kernel when 128 % W_sp == 0; backward: wide tcgen05 kernel always
            assert cw.tc_launch_count() == t0 + (2 if 128 % m.W_sp == 0 else 1), (key, cw.tc_launch_count() - t0)
        sl = (Ellipsis, slice(off, off + cb))
        assert rel(gb[sl], ref[0][sl]) <= tol, (key, "dqkv", rel(gb[sl], ref[0][sl]))
        assert rel(gw, ref[1]) <= tol and rel(gbias, ref[2]) <= tol, (key, rel(gw, ref[1]), rel(gbias, ref[2]))


def test_eval_after_native_training_sees_the_trained_weights():
    """the fused SGD writes parameters through raw pointers (no `_version` bump); eval-mode forwards and a
    SliceEngine captured BEFORE training must run with the trained weights afterwards (interleaved validation is a reference
    workflow: universal_train.py:644/868)."""
    torch.manual_seed(0)
    m = cw.cswin_tiny_224(num_classes=4, drop_path_rate=0.0).to(DEV)
    x = T(synth.synth_image_batch(2, 3, 224, seed=3, kind="ct"))
    y = torch.randint(0, 4, (2, 224, 224), device=DEV)
    m.eval()
    m.compute_dtype = torch.bfloat16
    with torch.no_grad():
        before = m(x).float()
    eng = cw.SliceEngine(m, batch=2)
    lab_before = eng.predict(x.cpu())
    step = cw.TrainStep(m, lr=0.5, graph=True, warmup=1)
    for _ in range(4):
        step(x, y)
    torch.cuda.synchronize()
    m.eval()
    with torch.no_grad():
        after = m(x).float()
    fresh = cw.cswin_tiny_224(num_classes=4, drop_path_rate=0.0).to(DEV).eval()
    fresh.load_state_dict(m.state_dict(), strict=True)
    fresh.compute_dtype = torch.bfloat16
    with torch.no_grad():
        want = fresh(x).float()
    assert (after - before).abs().max().item() > 1e-2, "training did not change the logits: test is vacuous"
    assert torch.equal(after, want), f"stale derived weights after training: {(after - want).abs().max().item():.3e}"
    lab_after = eng.predict(x.cpu())                       # engine captured before training: must have re-captured itself
    lab_want = cw.SliceEngine(fresh, batch=2).predict(x.cpu())
    assert torch.equal(lab_after, lab_want)
    step.close()
