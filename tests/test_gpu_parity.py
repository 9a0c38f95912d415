"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every CUDA op is called through the C ABI
(cswin_unet_b200.ops / modules -> ctypes -> libcswin_b200.so) and compared with

  (i)  the golden vectors generated from the unmodified reference (tests/golden/*.npz), and
  (ii) the CPU oracle (oracle/cswin_oracle.py) on the same seeded inputs.

Tolerances (north_star): fp32 path max-abs <= 1e-4 against the fp32/fp64 reference; bf16 path max-abs <= 2e-2 on
logits of the reference-init model; op-level bf16 tolerances are stated at each test and are relative to O(1)
activations whose inputs were first rounded to bf16 (so only in-kernel rounding is measured).
"""
import numpy as np
import pytest
import torch

import cswin_unet_b200 as cw
from cswin_unet_b200 import ops, synth
from oracle import cswin_oracle as O
from tests import golden_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda"
LEPE_EXTRA = ((32, 16, 0, 2, 1), (64, 16, 1, 2, 2), (64, 8, -1, 8, 2), (128, 16, 0, 8, 4), (48, 12, 1, 3, 3))
# BASELINE configs[4] (512^2) windows with golden vectors from the unmodified reference (tests/golden/lepe_wide.npz)
LEPE_WIDE_GOLD = ((128, 32, 0, 8, 4), (128, 32, 1, 8, 4), (512, 16, -1, 8, 16), (64, 24, 0, 8, 2), (64, 14, -1, 14, 2))


def T(a, dtype=torch.float32, device=DEV):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


def test_extension_is_loaded_and_counts_launches():
    n0 = cw.launch_count()
    x = torch.randn(64, 64, device=DEV)
    ops.layernorm(x, torch.ones(64, device=DEV), torch.zeros(64, device=DEV))
    assert cw.launch_count() == n0 + 1


# ---------------------------------------------------------------------------------------------------
# LePE attention
# ---------------------------------------------------------------------------------------------------
def lepe_case(cb, reso, idx, split, heads, B, dtype):
    full_c = cb if idx == -1 else 2 * cb
    base = T(synth.synth_qkv(B, reso, full_c, seed=0), dtype)
    off = cb if idx == 1 else 0
    qkv = base.permute(2, 0, 1, 3)[..., off:off + cb]              # the strided view the reference block passes
    m = cw.LePEAttention(cb, resolution=reso, idx=idx, split_size=split, num_heads=heads).to(DEV).eval()
    with torch.no_grad():
        m.get_v.weight.copy_(T(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.weight", (cb, 1, 3, 3), 1)))
        m.get_v.bias.copy_(T(synth.synth_tensor(f"lepe/{cb}/{reso}/{idx}/get_v.bias", (cb,), 1)))
    return m, qkv


@pytest.mark.parametrize("tag,cfgs,B", [("t224", synth.LEPE_CONFIGS_T224, 1), ("extra", LEPE_EXTRA, 2), ("wide", LEPE_WIDE_GOLD, 1)])
def test_lepe_attention_fp32_vs_golden(tag, cfgs, B):
    z = G.load(f"lepe_{tag}")
    for (cb, reso, idx, split, heads) in cfgs:
        m, qkv = lepe_case(cb, reso, idx, split, heads, B, torch.float32)
        assert not qkv.is_contiguous()
        with torch.no_grad():
            y = m(qkv)
        G.compare(z, f"c{cb}_r{reso}_i{idx}_s{split}_h{heads}", y.cpu().numpy(), atol=1e-4)


@pytest.mark.parametrize("tag,cfgs,B", [("t224", synth.LEPE_CONFIGS_T224, 1), ("extra", LEPE_EXTRA, 2)])
def test_lepe_attention_bf16_vs_oracle(tag, cfgs, B):
    for (cb, reso, idx, split, heads) in cfgs:
        m, qkv = lepe_case(cb, reso, idx, split, heads, B, torch.bfloat16)
        t0 = cw.tc_launch_count()
        with torch.no_grad():
            y = m(qkv).float().cpu()
        if cb // heads == 32 and m.H_sp * m.W_sp <= 128:
            assert cw.tc_launch_count() == t0 + 1, "the tcgen05 attention kernel must serve head_dim 32, N <= 128"
        q, k, v = (qkv[i].float().cpu().double() for i in range(3))            # bf16-rounded inputs, exact in fp64
        w = m.get_v.weight.detach().bfloat16().double().cpu()
        b = m.get_v.bias.detach().bfloat16().double().cpu()
        ref = O.lepe_attention(q, k, v, w, b, reso, idx, split, heads)
        err = (y.double() - ref).abs().max().item()
        # |out| is O(1..4): bf16 output rounding (2^-9 relative) + bf16 P rounding in the tensor-core path
        assert err <= 3e-2, f"{(cb, reso, idx, split, heads)}: bf16 max-abs {err:.3e}"


LEPE_WIDE = ((128, 32, 0, 8, 4), (128, 32, 1, 8, 4), (512, 16, -1, 8, 16), (64, 24, 0, 8, 2), (32, 16, 0, 16, 1), (32, 48, 0, 4, 1))


@pytest.mark.parametrize("B", [1, 3])
def test_lepe_attention_bf16_wide_windows_on_tcgen05(B):
    """BASELINE configs[4] (512^2, split [1,2,8,8]): stripe windows of 256 tokens (32x8, 8x32, 16x16) — and ragged ones of 192
    (24x8, 48x4: the second query tile is half empty) — run on the wide variant of the tcgen05 kernel (two 128-row query
    tiles per (window, head), S = 128 x 256 in TMEM).  Checked against the fp64 oracle on the bf16-rounded inputs."""
    for (cb, reso, idx, split, heads) in LEPE_WIDE:
        m, qkv = lepe_case(cb, reso, idx, split, heads, B, torch.bfloat16)
        assert 128 < m.H_sp * m.W_sp <= 256 and 128 % m.W_sp == 0
        t0 = cw.tc_launch_count()
        with torch.no_grad():
            y = m(qkv).float().cpu()
        assert cw.tc_launch_count() == t0 + 1, f"{(cb, reso, idx, split, heads)}: wide windows must run on the tcgen05 kernel"
        q, k, v = (qkv[i].float().cpu().double() for i in range(3))
        w = m.get_v.weight.detach().bfloat16().double().cpu()
        b = m.get_v.bias.detach().bfloat16().double().cpu()
        ref = O.lepe_attention(q, k, v, w, b, reso, idx, split, heads)
        err = (y.double() - ref).abs().max().item()
        assert err <= 3e-2, f"{(cb, reso, idx, split, heads)} B={B}: bf16 max-abs {err:.3e}"


def test_lepe_attention_batch24_matches_oracle_fp32():
    """BASELINE config 2 at the bench batch size: both branches of a stage-2 block in ONE launch."""
    B, reso, C, heads, split = 24, 28, 128, 4, 2
    base = T(synth.synth_qkv(B, reso, C, seed=3))
    q, k, v = base[:, :, 0], base[:, :, 1], base[:, :, 2]
    blk = cw.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, qkv_bias=True).to(DEV).eval()
    out = torch.empty(B, reso * reso, C, device=DEV)
    h = C // 2
    descs = [a.branch_desc(q[..., i * h:(i + 1) * h], k[..., i * h:(i + 1) * h], v[..., i * h:(i + 1) * h],
                           out[..., i * h:(i + 1) * h]) for i, a in enumerate(blk.attns)]
    ops.lepe_attention_fwd(descs, B, reso, float(blk.attns[0].scale), torch.float32)
    qc, kc, vc = (t.cpu() for t in (q, k, v))
    for i, a in enumerate(blk.attns):
        sl = slice(i * h, (i + 1) * h)
        ref = O.lepe_attention(qc[..., sl], kc[..., sl], vc[..., sl], a.get_v.weight.detach().cpu(),
                               a.get_v.bias.detach().cpu(), reso, i, split, heads // 2)
        assert (out[..., sl].cpu() - ref).abs().max().item() <= 1e-4


def test_lepe_attention_rejects_bad_shapes():
    m = cw.LePEAttention(32, resolution=8, idx=0, split_size=3, num_heads=1).to(DEV).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="not divisible"):
        m(torch.zeros(3, 1, 64, 32, device=DEV))
    m = cw.LePEAttention(32, resolution=8, idx=0, split_size=2, num_heads=1).to(DEV).eval()
    with torch.no_grad():
        assert m(torch.zeros(3, 0, 64, 32, device=DEV)).shape == (0, 64, 32)          # empty batch is a no-op


# ---------------------------------------------------------------------------------------------------
# LayerNorm / Linear
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("M,C", [(1000, 64), (333, 128), (77, 256), (49, 512), (5, 96)])
def test_layernorm(dtype, tol, M, C):
    g = torch.Generator().manual_seed(M * 7 + C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(dtype)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(dtype)
    b = (0.1 * torch.randn(C, generator=g)).to(dtype)
    y = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5)
    ref = O._ln(x.double(), w.double(), b.double(), 1e-5)
    assert (y.cpu().double() - ref).abs().max().item() <= tol


LINEAR_CASES = [
    # M, N, K1, K2, ln, act, res, drop
    (3136, 192, 64, 0, True, 0, False, False),      # stage-1 LN+qkv
    (3136, 64, 64, 0, False, 0, True, True),        # proj + residual + DropPath
    (784, 512, 128, 0, True, 1, False, False),      # LN + fc1 + GELU
    (784, 128, 512, 0, False, 0, True, False),      # fc2 + residual
    (196, 256, 256, 256, False, 0, False, False),   # concat_linear on two sources
    (98, 9, 64, 0, False, 0, False, False),         # folded head: N = 9 classes
    (130, 36, 288, 0, False, 0, False, False),      # CARAFE encoder as GEMM (ragged M, N)
    (201, 64, 152, 0, False, 0, False, False),      # stem conv as GEMM, K padded 147 -> 152
    (49, 2048, 512, 0, True, 1, False, False),      # stage-4 fc1
    (49, 512, 2048, 0, False, 0, True, True),       # stage-4 fc2
]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("M,N,K1,K2,ln,act,res,drop", LINEAR_CASES)
def test_linear_fused(dtype, tol, M, N, K1, K2, ln, act, res, drop):
    g = torch.Generator().manual_seed(M + N + K1)
    rps = 49 if M % 49 == 0 else M
    a = torch.randn(M, K1, generator=g).to(dtype)
    a2 = torch.randn(M, K2, generator=g).to(dtype) if K2 else None
    w = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).to(dtype)
    bias = (0.1 * torch.randn(N, generator=g)).to(dtype)
    gam = (1 + 0.1 * torch.randn(K1, generator=g)).to(dtype)
    bet = (0.1 * torch.randn(K1, generator=g)).to(dtype)
    r = torch.randn(M, N, generator=g).to(dtype) if res else None
    ss = (torch.bernoulli(torch.full((M // rps,), 0.8), generator=g) / 0.8).float() if drop else None
    y = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), a2=None if a2 is None else a2.to(DEV),
                   ln=(gam.to(DEV), bet.to(DEV), 1e-5) if ln else None, act=act,
                   residual=None if r is None else r.to(DEV), sample_scale=None if ss is None else ss.to(DEV),
                   rows_per_sample=rps)
    A = a.double()
    if ln:
        A = O._ln(A, gam.double(), bet.double(), 1e-5)
        if dtype == torch.bfloat16:
            A = A.bfloat16().double()                     # the tensor-core path rounds the normalised operand to bf16
    if a2 is not None:
        A = torch.cat([A, a2.double()], -1)
    t = A @ w.double().T + bias.double()
    if act:
        t = O._gelu(t)
    if ss is not None:
        t = t * ss.double().repeat_interleave(rps).view(M, 1)
    if r is not None:
        t = t + r.double()
    err = (y.cpu().double() - t).abs().max().item()
    assert y.shape == (M, N) and err <= tol, f"max-abs {err:.3e}"


PERSIST_CASES = [
    # M, N, K1, K2, act, res, stats : more than one wave of 128-row tiles -> the persistent form of linear_tc_kernel (gemm_tc.cu,
    # kPersist: tiles walked by resident CTAs, accumulator double-buffered in TMEM, operand ring running across tiles)
    (38457, 64, 64, 0, 0, True, True),
    (38457, 16, 64, 0, 0, False, False),       # BN = 16: half of the epilogue warps own no column unit
    (38457, 96, 152, 0, 0, False, False),      # ragged K (TMA zero fill), BN = 96
    (38457, 128, 512, 0, 0, True, True),       # deep K loop, residual + row statistics
    (38457, 128, 128, 128, 0, False, False),   # two-source K loop (decoder concat Linear)
    (38457, 72, 64, 0, 0, True, False),        # ragged N
]


@pytest.mark.parametrize("M,N,K1,K2,act,res,stats", PERSIST_CASES)
def test_linear_persistent_form_multiwave(M, N, K1, K2, act, res, stats):
    g = torch.Generator().manual_seed(M + N + K1 + K2)
    a = torch.randn(M, K1, generator=g).bfloat16()
    a2 = torch.randn(M, K2, generator=g).bfloat16() if K2 else None
    w = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).bfloat16()
    bias = (0.1 * torch.randn(N, generator=g)).bfloat16()
    r = torch.randn(M, N, generator=g).bfloat16() if res else None
    n0 = cw.tc_launch_count()
    out = ops.linear(a.to(DEV), w.to(DEV), bias.to(DEV), a2=None if a2 is None else a2.to(DEV), act=act,
                     residual=None if r is None else r.to(DEV), want_stats=stats)
    assert cw.tc_launch_count() == n0 + 1                  # the tcgen05 kernel, not the SIMT fallback
    y, st = out if stats else (out, None)
    A = a.double() if a2 is None else torch.cat([a.double(), a2.double()], -1)
    ref = A @ w.double().T + bias.double()
    if res:
        ref = ref + r.double()
    yd = y.float().cpu().double()
    err = (yd - ref).abs().max().item()
    assert y.shape == (M, N) and err <= 4e-2, f"max-abs {err:.3e}"
    if stats:                                              # side channel == row sums of the bf16 values actually stored
        s2 = st.cpu().double().sum(1)
        assert (s2[:, 0] - yd.sum(1)).abs().max().item() <= 2e-3
        assert ((s2[:, 1] - (yd * yd).sum(1)).abs() / (yd * yd).sum(1)).max().item() <= 1e-5


# ---------------------------------------------------------------------------------------------------
# block / merge / carafe vs golden
# ---------------------------------------------------------------------------------------------------
def load_named(module, prefix, seed):
    sd = {k: T(synth.synth_tensor(prefix + k, tuple(v.shape), seed)) for k, v in module.state_dict().items()}
    module.load_state_dict(sd, strict=True)


BLOCKS = ((64, 56, 2, 1, False), (128, 28, 4, 2, False), (256, 14, 8, 7, False), (512, 7, 16, 7, True))


# bf16: inputs, weights and 5 intermediate activations of a block are rounded to bf16 (2^-9 relative each).  The bound is RELATIVE to
# the largest golden activation of the block: measured 0.6-0.7e-2 of it on B200 (smoke() prints it) — 2.5e-2 leaves room for other seeds
# and still fails a kernel that is wrong in any visible way (the old absolute 1.5e-1 did not).
BF16_BLOCK_REL = 2.5e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, None)])
def test_block_vs_golden(dtype, tol):
    z = G.load("block")
    for (dim, reso, heads, split, last) in BLOCKS:
        m = cw.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
        load_named(m, f"block/{dim}/", 3)
        x = T(synth.synth_tensor(f"block_in/{dim}", (2, reso * reso, dim), 4), dtype)
        with torch.no_grad():
            y = m(x)
        assert y.dtype == dtype
        scale = float(np.abs(z[f"d{dim}.rows"].astype(np.float64)).max())
        err = G.compare(z, f"d{dim}", y.float().cpu().numpy(), atol=tol if tol is not None else BF16_BLOCK_REL * scale)
        print(f"[block {dim} {dtype}] max-abs {err:.3e}, max |golden| {scale:.2f}, relative {err / scale:.2e}")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 6e-2)])
def test_merge_and_carafe_vs_golden(dtype, tol):
    z = G.load("merge_carafe")
    for (dim, reso) in ((64, 56), (128, 28), (256, 14)):
        m = cw.Merge_Block(dim, dim * 2).to(DEV).eval()
        load_named(m, f"merge/{dim}/", 5)
        x = T(synth.synth_tensor(f"merge_in/{dim}", (2, reso * reso, dim), 6), dtype)
        with torch.no_grad():
            y = m(x)
        G.compare(z, f"merge_d{dim}", y.float().cpu().numpy(), atol=tol)
    for (cls, dim, dout, reso, up) in ((cw.CARAFE, 512, 256, 7, 2), (cw.CARAFE, 128, 64, 28, 2), (cw.CARAFE4, 64, 64, 14, 4)):
        m = cls(dim, dout).to(DEV).eval()
        load_named(m, f"carafe/{dim}/{up}/", 7)
        x = T(synth.synth_tensor(f"carafe_in/{dim}/{up}", (2, reso * reso, dim), 8), dtype)
        with torch.no_grad():
            y = m(x)
        G.compare(z, f"carafe_d{dim}_u{up}", y.float().cpu().numpy(), atol=tol)


# ---------------------------------------------------------------------------------------------------
# whole model
# ---------------------------------------------------------------------------------------------------
def build_model(mode):
    m = cw.cswin_tiny_224(num_classes=9).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234, mode=mode).items()}, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("mode,gold", [("alive", "model_t224"), ("refinit", "model_refinit")])
def test_full_model_fp32_logits(mode, gold):
    z = G.load(gold)
    m = build_model(mode)
    for kind in ("randn", "ct"):
        x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind=kind))
        with torch.no_grad():
            logits = m(x)
        assert logits.shape == (2, 9, 224, 224) and logits.dtype == torch.float32
        err = G.compare(z, f"logits_{kind}", logits.permute(0, 2, 3, 1).cpu().numpy(), atol=1e-4)
        agree = (logits.argmax(1).cpu().numpy() == z[f"argmax_{kind}"]).mean()
        print(f"[fp32 {mode}/{kind}] max-abs {err:.2e} argmax agreement {agree:.6f}")
        assert agree >= 0.999


def test_full_model_bf16_logits_reference_init():
    """north_star: bf16 logits max-abs <= 2e-2 vs the fp32 reference (random-init distributions)."""
    z = G.load("model_refinit")
    m = build_model("refinit")
    m.compute_dtype = torch.bfloat16
    for kind in ("randn", "ct"):
        x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind=kind))
        with torch.no_grad():
            logits = m(x).float()
        err = G.compare(z, f"logits_{kind}", logits.permute(0, 2, 3, 1).cpu().numpy(), atol=2e-2)
        am = logits.argmax(1).cpu().numpy()
        agree = (am == z[f"argmax_{kind}"]).mean()
        decisive = z[f"margin_{kind}"].astype(np.float32) >= 0.01
        agree_dec = (am == z[f"argmax_{kind}"])[decisive].mean()
        print(f"[bf16 refinit/{kind}] max-abs {err:.2e} argmax agreement {agree:.5f} "
              f"(reference's own bf16: {float(z[f'refbf16_argmax_agree_{kind}']):.5f}, max-abs "
              f"{float(z[f'refbf16_maxabs_{kind}']):.2e}); on margin>=0.01 pixels {agree_dec:.6f}")
        assert agree_dec >= 0.999
        assert agree >= float(z[f"refbf16_argmax_agree_{kind}"]) - 0.003


def test_full_model_bf16_logits_alive_weights():
    z = G.load("model_t224")
    m = build_model("alive")
    m.compute_dtype = torch.bfloat16
    x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind="randn"))
    with torch.no_grad():
        logits = m(x).float()
    # logits here have std ~1 and |max| ~2.8 (7x the reference-init scale): tolerance scaled accordingly
    err = G.compare(z, "logits_randn", logits.permute(0, 2, 3, 1).cpu().numpy(), atol=1.5e-1)
    am = logits.argmax(1).cpu().numpy()
    decisive = z["margin_randn"].astype(np.float32) >= 0.05
    print(f"[bf16 alive] max-abs {err:.2e} argmax agreement {(am == z['argmax_randn']).mean():.5f}")
    assert (am == z["argmax_randn"])[decisive].mean() >= 0.999


def test_batch_independence_and_cuda_graph_replay():
    """Slices are independent (what makes slice-sharding legal) and the forward is CUDA-graph capturable."""
    m = build_model("alive")
    x = T(synth.synth_image_batch(4, 3, 224, seed=5, kind="ct"))
    with torch.no_grad():
        full = m(x)
        part = m(x[1:3])
    assert torch.equal(full[1:3], part)
    static_x = x.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(2):
            m(static_x)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        static_y = m(static_x)
    static_x.copy_(torch.flip(x, dims=(0,)))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_y, torch.flip(full, dims=(0,)))


def test_droppath_training_forward_matches_oracle():
    blk = cw.CSWinBlock(dim=128, reso=28, num_heads=4, split_size=2, qkv_bias=True, drop_path=0.3).to(DEV).train()
    load_named(blk, "block/128/", 3)
    x = T(synth.synth_tensor("block_in/128", (6, 784, 128), 4))
    torch.manual_seed(11)
    with torch.no_grad():
        y = blk(x)
    torch.manual_seed(11)
    s1 = blk.drop_path.sample_scale(x); s2 = blk.drop_path.sample_scale(x)
    sd = {k: v.detach().cpu() for k, v in blk.state_dict().items()}
    xc = x.cpu()
    # oracle with the two masks applied (x + s1*proj(.) ; x1 + s2*mlp(.)) — restated here because the oracle takes one mask
    ref1 = O.cswin_block(sd, "", xc, 28, 4, 2, False, sample_scale=None)
    assert not torch.allclose(y.cpu(), ref1, atol=1e-3)            # masks really were applied
    u = O._ln(xc, sd["norm1.weight"], sd["norm1.bias"], 1e-5) @ sd["qkv.weight"].T + sd["qkv.bias"]
    q, k, v = u[..., :128], u[..., 128:256], u[..., 256:]
    a = torch.cat([O.lepe_attention(q[..., i * 64:(i + 1) * 64], k[..., i * 64:(i + 1) * 64], v[..., i * 64:(i + 1) * 64],
                                    sd[f"attns.{i}.get_v.weight"], sd[f"attns.{i}.get_v.bias"], 28, i, 2, 2) for i in range(2)], -1)
    x1 = xc + s1.cpu().view(-1, 1, 1) * (a @ sd["proj.weight"].T + sd["proj.bias"])
    hdn = O._gelu(O._ln(x1, sd["norm2.weight"], sd["norm2.bias"], 1e-5) @ sd["mlp.fc1.weight"].T + sd["mlp.fc1.bias"])
    ref = x1 + s2.cpu().view(-1, 1, 1) * (hdn @ sd["mlp.fc2.weight"].T + sd["mlp.fc2.bias"])
    assert (y.cpu() - ref).abs().max().item() <= 1e-4


# ---------------------------------------------------------------------------------------------------
# volume inference (test_single_volume loop, utils.py:61-90) through the SliceEngine: Dice / HD95 parity
# ---------------------------------------------------------------------------------------------------
def _synthetic_volume(D=6, S=160, seed=0):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32) / S
    vol = np.zeros((D, S, S), np.float32)
    for d in range(D):
        for _ in range(5):
            cy, cx, r, a = g.uniform(0.2, 0.8), g.uniform(0.2, 0.8), g.uniform(0.05, 0.2), g.uniform(0.3, 1.0)
            vol[d] += a * np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
        vol[d] = np.clip(vol[d] / vol[d].max() + 0.02 * g.standard_normal((S, S)), 0, 1)
    return vol


@pytest.mark.parametrize("dtype,dice_tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
def test_volume_inference_dice_hd95_vs_oracle(dtype, dice_tol):
    from scipy.ndimage import zoom
    vol = _synthetic_volume()
    D, S, _ = vol.shape
    m = build_model("alive")
    eng = cw.SliceEngine(m, batch=4, compute_dtype=dtype)
    pred, rng = cw.predict_volume(eng, vol)
    assert pred.shape == (D, S, S) and pred.dtype == np.uint8 and list(rng) == list(range(D))
    # the same loop with the CPU oracle as the network (utils.py:61-80)
    shapes = O.state_dict_shapes()
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}
    ref = np.zeros((D, S, S), np.uint8)
    with torch.no_grad():
        for d in range(D):
            sl = zoom(vol[d], (224 / S, 224 / S), order=3)
            x = torch.from_numpy(sl)[None, None].float().repeat(1, 3, 1, 1)
            out = O.cswin_unet_forward(sd, x).softmax(1).argmax(1)[0].numpy()
            ref[d] = zoom(out, (S / 224, S / 224), order=0)
    agree = (pred == ref).mean()
    # "ground truth" for the metric comparison: the oracle's own labels shifted by one voxel (any fixed label volume works:
    # the criterion is that BOTH predictions score the same against it)
    gt = np.roll(ref, 1, axis=2)
    worst = 0.0
    for c in range(1, 9):
        d_ref, h_ref = O.dice_hd95_percase(ref == c, gt == c)
        d_new, h_new = O.dice_hd95_percase(pred == c, gt == c)
        if dtype != torch.float32 and (ref == c).mean() < 0.005:
            continue          # bf16: Dice of a class covering < 0.5 % of the voxels flips with a handful of near-tie pixels
        worst = max(worst, abs(d_ref - d_new))
        if dtype == torch.float32:
            assert abs(d_ref - d_new) <= 1e-3 and abs(h_ref - h_new) <= 1e-3, (c, d_ref, d_new, h_ref, h_new)
    print(f"[volume {dtype}] label agreement {agree:.6f}, worst per-class Dice difference {worst:.2e}")
    assert worst <= dice_tol
    if dtype == torch.float32:
        assert agree >= 0.9999


def test_predict_volume_sharded_equals_unsharded():
    vol = _synthetic_volume(D=5, S=224, seed=3)
    m = build_model("alive")
    eng = cw.SliceEngine(m, batch=2, compute_dtype=torch.bfloat16)
    full, _ = cw.predict_volume(eng, vol)
    parts = [cw.predict_volume(eng, vol, shard=(r, 3))[0] for r in range(3)]
    assert np.array_equal(np.concatenate(parts, 0), full)


@pytest.mark.parametrize("inflight", [2, 3])
def test_engine_concurrent_forwards_equal_sequential(inflight):
    """SliceEngine(inflight = K) runs K captured forwards concurrently on K streams: label maps (stream API, ragged last batch) and
    resampled volumes must equal the strictly sequential engine's, batch by batch, over several rounds of slot reuse."""
    m = build_model("alive")
    g = torch.Generator().manual_seed(5)
    batches = [torch.rand(3 if i != 6 else 2, 3, 224, 224, generator=g) for i in range(7)]
    seq = list(cw.SliceEngine(m, batch=3, compute_dtype=torch.bfloat16, inflight=1).predict_stream(iter(batches)))
    eng = cw.SliceEngine(m, batch=3, compute_dtype=torch.bfloat16, inflight=inflight)
    # slots = forwards in flight + one batch copying in + one draining; exactly `inflight` compute streams
    assert len(eng.slots) == (2 if inflight == 1 else inflight + 2) and len({s["stream"].cuda_stream for s in eng.slots}) == inflight
    for _ in range(2):
        con = list(eng.predict_stream(iter(batches)))
        assert len(con) == len(seq) and all(torch.equal(a, b) for a, b in zip(con, seq))
    # single-channel host batches: one channel crosses PCIe, the 1 -> 3 repeat (vision_transformer.py:40-41) happens on the device
    one = [b[:, :1].contiguous() for b in batches]
    want1 = list(eng.predict_stream(iter([b.expand(-1, 3, -1, -1).contiguous() for b in one])))
    got1 = list(eng.predict_stream(iter(one)))
    assert all(torch.equal(a, b) for a, b in zip(got1, want1))
    vol = _synthetic_volume(D=8, S=256, seed=4)
    want, _ = cw.predict_volume(cw.SliceEngine(m, batch=3, compute_dtype=torch.bfloat16, inflight=1), vol, resample="gpu")
    got, _ = cw.predict_volume(eng, vol, resample="gpu")
    assert np.array_equal(np.asarray(got), np.asarray(want))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("M,N,K", [(3136, 64, 192), (784, 128, 512), (4704, 1024, 256), (201, 152, 64), (98, 40, 72)])
def test_linear_kn_weight_layout(dtype, tol, M, N, K):
    """w_layout = 1: out = a @ w with w (K, N) row-major — the data-gradient form dA = dZ W (a column slice of W works too)."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dtype)
    wfull = (torch.randn(K, N + 16, generator=g) / K ** 0.5).to(dtype)
    w = wfull[:, 8:8 + N]                                     # non-contiguous (K, N) view, 16-byte aligned rows
    y = ops.linear(a.to(DEV), wfull.to(DEV)[:, 8:8 + N], None, w_kn=True)
    ref = a.double() @ w.double()
    assert y.shape == (M, N) and (y.cpu().double() - ref).abs().max().item() <= tol


def test_512px_config_blocks_fp32_and_bf16_vs_oracle():
    """BASELINE configs[4]: 512^2 input with split [1,2,8,8] -> stripe windows of 128 / 128 / 256 / 256 tokens.  fp32 runs on
    the general SIMT kernel (exact); bf16 runs on the tcgen05 kernel (its wide variant for the 256-token windows)."""
    for (dim, reso, heads, split, last) in ((64, 128, 2, 1, False), (256, 32, 8, 8, False), (512, 16, 16, 8, True)):
        m = cw.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last).to(DEV).eval()
        load_named(m, f"block512/{dim}/", 3)
        x = T(synth.synth_tensor(f"block512_in/{dim}", (1, reso * reso, dim), 4))
        sd = {k: v.detach().cpu().double() for k, v in m.state_dict().items()}
        ref = O.cswin_block(sd, "", x.cpu().double(), reso, heads, split, last)
        with torch.no_grad():
            y32 = m(x).cpu().double()
            y16 = m(x.bfloat16()).float().cpu().double()
        assert (y32 - ref).abs().max().item() <= 1e-4, dim
        e16, scale = (y16 - ref).abs().max().item(), ref.abs().max().item()
        print(f"[512^2 block {dim}] bf16 max-abs {e16:.3e}, max |ref| {scale:.2f}, relative {e16 / scale:.2e}")
        assert e16 <= BF16_BLOCK_REL * scale, (dim, e16, scale)


@pytest.mark.parametrize("M,C,N,act,mean", [(3136 * 2, 64, 192, 0, 0.0), (784, 128, 512, 1, 0.7), (196 * 3, 256, 768, 0, -1.5),
                                            (49 * 5, 512, 2048, 1, 3.0), (130, 64, 256, 1, 0.2)])
def test_linear_folded_layernorm_chain(M, C, N, act, mean):
    """LayerNorm folded into the tcgen05 Linear (cswin_unet.py:168-169, :179): the producer Linear's epilogue emits the
    per-row (sum, sum^2) side channel, the consumer runs on the RAW rows with W o gamma and applies
    rstd * (acc - mean * colsum) + (b + W beta).  Checked against fp64 LN -> Linear on the same bf16 activations."""
    g = torch.Generator().manual_seed(M + C + N)
    a0 = torch.randn(M, C, generator=g).bfloat16()
    wp = (torch.randn(C, C, generator=g) / C ** 0.5).bfloat16()
    bp = (torch.randn(C, generator=g) * 0.1 + mean).bfloat16()
    res = torch.randn(M, C, generator=g).bfloat16()
    gam = (1 + 0.2 * torch.randn(C, generator=g)).float()
    bet = (0.2 * torch.randn(C, generator=g)).float()
    W = (torch.randn(N, C, generator=g) / C ** 0.5).float()
    b = (0.1 * torch.randn(N, generator=g)).float()
    x1, st = ops.linear(a0.to(DEV), wp.to(DEV), bp.to(DEV), residual=res.to(DEV), want_stats=True)
    # the side channel equals the row sums of the bf16 values actually stored
    xs = x1.float().cpu().double()
    s = st.cpu().double().sum(1)
    assert (s[:, 0] - xs.sum(1)).abs().max().item() <= 1e-3 * max(1.0, abs(mean) * C)
    assert ((s[:, 1] - (xs * xs).sum(1)).abs() / (xs * xs).sum(1)).max().item() <= 1e-5
    assert (ops.row_stats(x1).cpu().double()[:, 0] - s).abs().max().item() <= 1e-3 * max(1.0, abs(mean) * C)
    wf = (W * gam[None, :]).bfloat16()
    cs = wf.float().sum(1)
    bf = W @ bet + b
    y = ops.linear(x1, wf.to(DEV), None, ln_fold=(st, cs.to(DEV), 1e-5), bias_f32=bf.to(DEV), act=act)
    mu = xs.mean(1, keepdim=True)
    var = xs.var(1, unbiased=False, keepdim=True)
    ref = ((xs - mu) / (var + 1e-5).sqrt() * gam.double() + bet.double()) @ W.double().T + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (y.float().cpu().double() - ref).abs().max().item()
    assert err <= 4e-2, err                                  # bf16 operands + bf16 output rounding (|y| up to ~6)
    # against the unfused bf16 path (LayerNorm kernel -> Linear) the two agree to bf16 resolution
    y2 = ops.linear(x1, W.bfloat16().to(DEV), b.bfloat16().to(DEV), ln=(gam.bfloat16().to(DEV), bet.bfloat16().to(DEV), 1e-5), act=act)
    err2 = (y2.float().cpu().double() - ref).abs().max().item()
    assert err <= 2.0 * err2 + 1e-2, (err, err2)


def test_folded_layernorm_model_matches_unfolded():
    """Whole-model bf16 logits with the LN fold on (default) and off agree with the fp32 logits equally well."""
    from cswin_unet_b200 import modules as M_
    m = build_model("refinit")
    x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind="ct"))
    with torch.no_grad():
        ref = m(x).float()
        m.compute_dtype = torch.bfloat16
        old = M_.FOLD_LN
        try:
            M_.FOLD_LN = True
            n0 = cw._lib.launch_count()
            y_f = m(x).float()
            n_f = cw._lib.launch_count() - n0
            M_.FOLD_LN = False
            n0 = cw._lib.launch_count()
            y_u = m(x).float()
            n_u = cw._lib.launch_count() - n0
        finally:
            M_.FOLD_LN = old
    e_f, e_u = (y_f - ref).abs().max().item(), (y_u - ref).abs().max().item()
    assert e_f <= 2e-2 and e_u <= 2e-2, (e_f, e_u)
    assert abs((y_f.argmax(1) == ref.argmax(1)).float().mean().item() - (y_u.argmax(1) == ref.argmax(1)).float().mean().item()) <= 3e-3
    assert n_f <= n_u - 40, (n_f, n_u)                       # ~52 LayerNorm launches gone, 4 row_stats added, stage-1 MLPs fused                       # ~52 LayerNorm launches disappear, 4 row_stats appear


@pytest.mark.parametrize("M,C,mean", [(300, 64, 0.0), (75264, 64, 0.3), (18816, 128, -0.5), (4704, 256, 0.7), (130, 256, 0.0),
                                      (128 * 5 + 1, 128, 2.0)])
def test_fused_mlp_vs_fp64(M, C, mean):
    """cswin_mlp_fwd: x + fc2(GELU(fc1(LayerNorm(x)))) (cswin_unet.py:179, Mlp :22-26) in one tcgen05 launch, against fp64 on
    the same bf16 inputs and against the composed (LayerNorm kernel -> Linear+GELU -> Linear+residual) bf16 path."""
    g = torch.Generator().manual_seed(M + C)
    hid = 4 * C
    x = (torch.randn(M, C, generator=g) + mean).bfloat16()
    gam = (1 + 0.2 * torch.randn(C, generator=g)).float()
    bet = (0.2 * torch.randn(C, generator=g)).float()
    W1 = (torch.randn(hid, C, generator=g) / C ** 0.5).float()
    b1 = (0.1 * torch.randn(hid, generator=g)).float()
    W2 = (torch.randn(C, hid, generator=g) / hid ** 0.5).bfloat16()
    b2 = (0.1 * torch.randn(C, generator=g)).float()
    xd = x.to(DEV)
    st = ops.row_stats(xd)
    w1f = (W1 * gam[None, :]).bfloat16()
    y, st2 = ops.mlp_fused(xd, w1f.to(DEV), w1f.float().sum(1).to(DEV), (W1 @ bet + b1).to(DEV), W2.to(DEV), b2.to(DEV), st, 1e-5)
    xs = x.double()
    mu, var = xs.mean(1, keepdim=True), xs.var(1, unbiased=False, keepdim=True)
    u = (xs - mu) / (var + 1e-5).sqrt() * gam.double() + bet.double()
    h = torch.nn.functional.gelu(u @ W1.double().T + b1.double())
    ref = xs + h @ W2.double().T + b2.double()
    err = (y.float().cpu().double() - ref).abs().max().item()
    # composed bf16 path on the same operands
    u16 = ops.layernorm(xd, gam.bfloat16().to(DEV), bet.bfloat16().to(DEV), 1e-5)
    h16 = ops.linear(u16, W1.bfloat16().to(DEV), b1.bfloat16().to(DEV), act=1)
    y2 = ops.linear(h16, W2.to(DEV), b2.bfloat16().to(DEV), residual=xd)
    err2 = (y2.float().cpu().double() - ref).abs().max().item()
    print(f"[fused mlp M={M} C={C}] max-abs vs fp64: fused {err:.3e}, composed {err2:.3e}")
    assert err <= 1.5 * err2 + 2e-2, (err, err2)
    # the statistics side channel describes the stored rows
    ys = y.float().cpu().double()
    s = st2.cpu().double().sum(1)
    assert (s[:, 0] - ys.sum(1)).abs().max().item() <= 1e-3 * C * max(1.0, abs(mean))
    assert ((s[:, 1] - (ys * ys).sum(1)).abs() / (ys * ys).sum(1)).max().item() <= 1e-5


def test_fused_mlp_rejects_unsupported():
    x = torch.zeros(64, 512, device=DEV, dtype=torch.bfloat16)
    assert not ops.mlp_supported(512, 2048) and not ops.mlp_supported(256, 512)
    with pytest.raises(cw.CswinError):
        ops.mlp_fused(x, torch.zeros(2048, 512, device=DEV, dtype=torch.bfloat16), torch.zeros(2048, device=DEV),
                      torch.zeros(2048, device=DEV), torch.zeros(512, 2048, device=DEV, dtype=torch.bfloat16),
                      torch.zeros(512, device=DEV), ops.row_stats(x), 1e-5, want_stats=False)


@pytest.mark.parametrize("M,N,K", [(1176, 576, 36), (75264, 64, 9), (300, 40, 20)])
def test_linear_bf16_ragged_k_on_tensor_cores(M, N, K):
    """Contraction lengths that are not multiples of 8 (the CARAFE encoder's 36 / the head's 9 output channels seen from their
    data gradients) stay on the tcgen05 kernel as long as the row pitches are 16-byte multiples: TMA zero-fills the tail."""
    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 7) // 8 * 8
    a = torch.zeros(M, ld, dtype=torch.bfloat16); a[:, :K] = torch.randn(M, K, generator=g).bfloat16()
    a[:, K:] = 7.0                                            # poison the padding: it must never be read as data
    w = torch.zeros(N, ld, dtype=torch.bfloat16); w[:, :K] = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    w[:, K:] = -3.0
    wkn = (torch.randn(K, N, generator=g) / K ** 0.5).bfloat16()
    ad, wd = a.to(DEV)[:, :K], w.to(DEV)[:, :K]
    n0 = cw.tc_launch_count()
    y1 = ops.linear(ad, wd, None)
    y2 = ops.linear(ad, wkn.to(DEV), None, w_kn=True)
    assert cw.tc_launch_count() - n0 == 2
    r1 = a[:, :K].double() @ w[:, :K].double().T
    r2 = a[:, :K].double() @ wkn.double()
    assert (y1.float().cpu().double() - r1).abs().max().item() <= 3e-2
    assert (y2.float().cpu().double() - r2).abs().max().item() <= 3e-2


# ---------------------------------------------------------------------------------------------------
# slice resampling (scipy.ndimage.zoom of utils.py:69 / :77) on the GPU
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,H,W,P", [(3, 512, 512, 224), (2, 40, 36, 17), (1, 64, 64, 28), (2, 300, 200, 224), (1, 160, 144, 224)])
def test_zoom_cubic_matches_scipy(n, H, W, P):
    from scipy.ndimage import zoom
    rng = np.random.default_rng(n + H + P)
    x = rng.random((n, H, W)).astype(np.float32)
    ref = np.stack([zoom(s, (P / H, P / W), order=3) for s in x])
    out = ops.zoom_cubic(torch.from_numpy(x).to(DEV), (P, P), out=torch.zeros((n, 3, P, P), device=DEV))
    y = out.cpu().numpy()
    assert np.array_equal(y[:, 0], y[:, 1]) and np.array_equal(y[:, 0], y[:, 2])          # the 1 -> 3 channel repeat
    err = np.abs(y[:, 0] - ref).max()
    exact = (y[:, 0] == ref).mean()
    print(f"[zoom cubic {H}x{W}->{P}] max-abs {err:.2e}, bit-identical {exact:.6f}")
    assert err <= 1e-6 and exact >= 0.999
    if (H, W, P) == (512, 512, 224):                     # scipy's edge rule: 223 * (511 / 223) > 511 -> cval
        assert (ref[:, -1] == 0).all() and (y[:, 0, -1] == 0).all() and (y[:, 0, :, -1] == 0).all()


@pytest.mark.parametrize("n,P,H,W", [(3, 224, 512, 512), (2, 28, 64, 64), (2, 224, 300, 200), (1, 17, 40, 36)])
def test_zoom_nearest_matches_scipy(n, P, H, W):
    from scipy.ndimage import zoom
    rng = np.random.default_rng(n + H + P)
    lab = (rng.random((n, P, P)) * 9).astype(np.uint8)
    ref = np.stack([zoom(l, (H / P, W / P), order=0) for l in lab])
    y = ops.zoom_nearest_u8(torch.from_numpy(lab).to(DEV), (H, W)).cpu().numpy()
    assert y.shape == ref.shape and np.array_equal(y, ref)


def test_predict_volume_gpu_resampling_equals_host_resampling():
    """predict_volume(resample='gpu') (cswin_zoom_cubic_fwd -> forward -> cswin_zoom_nearest_u8, nothing but raw slices in and
    label maps out) gives the label maps of the reference's host-side scipy loop (utils.py:61-80)."""
    vol = _synthetic_volume(D=7, S=160, seed=5)
    m = build_model("alive")
    eng = cw.SliceEngine(m, batch=3, compute_dtype=torch.bfloat16)
    host, _ = cw.predict_volume(eng, vol)
    gpu, rng = cw.predict_volume(eng, vol, resample="gpu")
    assert gpu.shape == host.shape and gpu.dtype == np.uint8 and list(rng) == list(range(7))
    agree = (gpu == host).mean()
    print(f"[volume gpu-resample] label agreement with the scipy path {agree:.6f}")
    assert agree >= 0.9999
    part = cw.predict_volume(eng, vol, shard=(1, 2), resample="gpu")
    assert np.array_equal(part[0], gpu[list(part[1])])


def test_full_model_512px_config_logits_vs_golden():
    """BASELINE configs[4]: the native model at 512^2 (3 classes, split [1,2,8,8]) against the unmodified reference's logits
    (tests/golden/model_512.npz): fp32 <= 1e-4 and argmax agreement; the bf16 forward (wide tcgen05 attention kernels) is run
    and reported against the same vectors."""
    z = G.load("model_512")
    m = cw.cswin_tiny_224(num_classes=3, img_size=512, split_size=[1, 2, 8, 8]).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    m = m.to(DEV)
    x = T(synth.synth_image_batch(1, 3, 512, seed=0, kind="ct"))
    with torch.no_grad():
        logits = m(x)
        m.compute_dtype = torch.bfloat16
        t0 = cw.tc_launch_count()
        l16 = m(x).float()
        n_tc = cw.tc_launch_count() - t0
    assert logits.shape == (1, 3, 512, 512) and logits.dtype == torch.float32
    err = G.compare(z, "logits_ct", logits.permute(0, 2, 3, 1).cpu().numpy(), atol=1e-4)
    agree = (logits.argmax(1).cpu().numpy() == z["argmax_ct"]).mean()
    a16 = (l16.argmax(1).cpu().numpy() == z["argmax_ct"]).mean()
    e16 = (l16 - logits).abs().max().item()
    print(f"[512px fp32] max-abs {err:.2e} argmax agreement {agree:.6f}; [bf16] max-abs vs fp32 {e16:.2e}, argmax agreement {a16:.6f}, "
          f"{n_tc} tcgen05 launches")
    assert agree >= 0.999
    # bf16 numbers are reported only: with these "alive" synthetic weights and 3 classes the argmax of a smooth input sits on
    # near-ties (the bf16 criteria of the north star are asserted on the reference-init T224 model above, and the 512^2 bf16
    # kernels block by block in test_512px_config_blocks_* / test_lepe_attention_bf16_wide_windows_on_tcgen05)
    assert torch.isfinite(l16).all()


# ---------------------------------------------------------------------------------------------------
# GELU over the range trained checkpoints produce (the degree-7 tanh-argument fit turns over near |x| = 7.3)
# ---------------------------------------------------------------------------------------------------
def _wide_preacts(M, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = (torch.rand(M, N, generator=g) * 40.0 - 20.0)
    z[0, :8] = torch.tensor([-20.0, -8.0, -7.4, -5.5, 5.5, 7.4, 8.0, 20.0])
    return z.bfloat16()


def test_gelu_epilogue_wide_range_bf16():
    """Linear + GELU epilogue (tcgen05 fast path and predicated path), fused MLP and the activation kernel with pre-activations
    in [-20, 20]: bf16 result within one bf16 ulp of exact GELU(erf) of the same pre-activation (Mlp, cswin_unet.py:22-26)."""
    M, K = 256, 64
    z = _wide_preacts(M, K)
    eye = torch.eye(K).bfloat16()
    ref = torch.nn.functional.gelu(z.double()).float()
    tol = lambda r: 2.0 ** -8 * r.abs() + 2e-3                        # one bf16 ulp relative + the documented tanh.approx floor
    # (a) GEMM with identity weights: accumulator == z exactly, epilogue = GELU
    for n_out in (K, K - 4):                                          # N % 8 == 0 -> TMA-store path; ragged N -> predicated path
        y = ops.linear(z.to(DEV), eye[:n_out].contiguous().to(DEV), None, act=1).float().cpu()
        assert ((y - ref[:, :n_out]).abs() <= tol(ref[:, :n_out])).all(), (y - ref[:, :n_out]).abs().max()
        assert abs(y[0, 0].item()) <= 2e-3 and abs(y[0, 1].item()) <= 2e-3 and abs(y[0, 6].item() - 8.0) < 0.04 and abs(y[0, 7].item() - 20.0) < 0.1
    # (b) streaming activation kernel
    y = ops.act_fwd(z.to(DEV), act=1).float().cpu()
    assert ((y - ref).abs() <= tol(ref)).all()
    # (c) fused MLP: LayerNorm with a huge gamma pushes fc1 pre-activations far outside +-7
    C, hid = 64, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(300, C, generator=g).bfloat16()
    gam = torch.full((C,), 12.0)
    W1 = (torch.randn(hid, C, generator=g) / C ** 0.5)
    w1f = (W1 * gam[None, :]).bfloat16()
    W2 = (torch.randn(C, hid, generator=g) / hid ** 0.5).bfloat16()
    xd = x.to(DEV)
    yf, _ = ops.mlp_fused(xd, w1f.to(DEV), w1f.float().sum(1).to(DEV), torch.zeros(hid, device=DEV), W2.to(DEV),
                          torch.zeros(C, device=DEV), ops.row_stats(xd), 1e-5)
    xs = x.double()
    u = (xs - xs.mean(1, keepdim=True)) / (xs.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt()
    pre = u @ w1f.double().T
    assert pre.abs().max() > 20.0
    refm = xs + torch.nn.functional.gelu(pre) @ W2.double().T
    err = (yf.float().cpu().double() - refm).abs().max().item()
    assert err <= 2e-2 * refm.abs().max().item(), (err, refm.abs().max().item())


def test_gelu_grad_wide_range_bf16():
    """GELU' in the streaming kernel (cswin_act_bwd) and in the fc2 data-gradient epilogue (act = 2) for |z| up to 20."""
    M, N = 256, 64
    z = _wide_preacts(M, N, seed=1)
    dh = torch.ones(M, N).bfloat16()
    zd = z.double()
    ref = (0.5 * (1 + torch.erf(zd / 2 ** 0.5)) + zd * torch.exp(-zd * zd / 2) / (2 * torch.pi) ** 0.5).float()
    dz = ops.act_bwd(dh.to(DEV), z.to(DEV), None, 0, act=1).float().cpu()
    assert ((dz - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-3).all(), (dz - ref).abs().max()
    assert abs(dz[0, 0].item()) < 1e-6 and abs(dz[0, 1].item()) < 1e-6 and abs(dz[0, 6].item() - 1.0) < 5e-3 and abs(dz[0, 7].item() - 1.0) < 5e-3
    # act = 2 epilogue: dZ = (dH2 @ W) o GELU'(z) with W = I (w_kn layout) -> GELU'(z) itself
    eye = torch.eye(N).bfloat16()
    dz2 = ops.linear(dh.to(DEV), eye.to(DEV), w_kn=True, act=2, residual=z.to(DEV)).float().cpu()
    assert ((dz2 - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-3).all(), (dz2 - ref).abs().max()


def test_bf16_simt_fallback_is_counted_and_the_model_never_takes_it():
    """a bf16 call outside the tcgen05 envelope (here a 20x20 = 400-token window) runs on the general SIMT
    kernel, which is reported (counter + one warning line), and the cswin_tiny forward never falls back."""
    m, qkv = lepe_case(64, 20, -1, 20, 2, 1, torch.bfloat16)                 # one 20x20 = 400-token window: > 256, SIMT kernel
    n0 = cw.simt_fallback_count()
    with torch.no_grad():
        y = m(qkv)
    assert cw.simt_fallback_count() == n0 + 1 and torch.isfinite(y.float()).all()
    model = build_model("alive")
    model.compute_dtype = torch.bfloat16
    x = T(synth.synth_image_batch(2, 3, 224, seed=0, kind="ct"))
    n1 = cw.simt_fallback_count()
    with torch.no_grad():
        model(x)
        model.predict_labels(x)
    assert cw.simt_fallback_count() == n1, "the cswin_tiny_224 bf16 forward took a SIMT fallback"
