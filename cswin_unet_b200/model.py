"""Host graph of CSWin-UNet built from the native modules.

`CSWinTransformer` keeps the reference's constructor signature, sub-module names and the exact 463-key
state_dict of /root/reference/networks/cswin_unet.py:322-554 (checked against the reference's key list in
tests/golden/model_t224.npz), so checkpoints move both ways with `strict=True`.  The reference file cannot be
shipped to the GPU box, so this assembly is what bench.py and the parity tests run there; in a checkout of the
reference, `cswin_unet_b200.install()` swaps the hot-path classes into the reference's own assembly instead.

Differences in *how* (not what) the forward is computed, all exact in real arithmetic:
  * stem conv 7x7/4 as an im2col gather + Linear + LayerNorm (:338-342);
  * skip `torch.cat` + `concat_linear` as one Linear with two A sources (:509-510, :518-519, :526-527);
  * head: CARAFE4 `out` (64->64, bias) and `output` (64->classes, no bias) are folded into one 64->classes map
    applied at 56x56 before the re-assembly, which writes NCHW logits directly (:536-544).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import os

import torch
import torch.nn as nn

from . import autograd as ag
from . import ops
from .modules import CARAFE, CARAFE4, CSWinBlock, Merge_Block, _Native, _no_autograd, run_stage

Tensor = torch.Tensor


# Inference forwards of a batch are split into CSWIN_STREAMS sub-batches that run on their own CUDA streams (fork / join, also
# inside CUDA-graph capture).  At cswin_tiny sizes every kernel is at most one or two waves and latency-bound — barrier / TMEM
# set-up, the first TMA round trip and the epilogue drain of one launch leave the tensor cores idle — so two independent launch
# chains on the same SMs could hide each other's fill and drain.  Samples are independent (no batch statistics anywhere), results
# are bit-identical to the unsplit forward.  MEASURED on B200 at batch 24 (profiles/r02_sweep_streams.log): 18,984 slices/s with one
# chain, 18,018 / 16,703 / 16,348 with 2 / 3 / 4 — each launch is bound by its own latency chain (first TMA round trip, K loop at
# the L2 -> SM ingest limit, epilogue), which does not get shorter with half the rows, and two chains contend for the same SMs.
# Default 1 (off); kept as an A/B switch.
N_STREAMS = max(1, int(os.environ.get("CSWIN_STREAMS", "1")))

# stem conv + LayerNorm as one implicit-GEMM launch (csrc/stem_tc.cu) instead of im2col + Linear + LayerNorm (bf16 inference)
FUSE_STEM = os.environ.get("CSWIN_FUSE_STEM", "1") != "0"

# training head folded like the inference head (CSWIN_UNFOLDED_TRAIN_HEAD=1 restores the reference's unfolded structure)
FOLD_TRAIN_HEAD = os.environ.get("CSWIN_UNFOLDED_TRAIN_HEAD") != "1"


def _view_keep_stats(y: Tensor, shape) -> Tensor:
    v = y.view(shape)
    st = getattr(y, "_cswin_stats", None)
    if st is not None:
        v._cswin_stats = st
    return v


class CSWinTransformer(_Native):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=8, embed_dim=64, depth=[1, 2, 9, 1],
                 split_size=[1, 2, 7, 7], num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0, hybrid_backbone=None, norm_layer=nn.LayerNorm, use_chk=False):
        super().__init__()
        if use_chk:
            raise NotImplementedError("use_chk: activation checkpointing is never enabled by the reference wrapper")
        self.use_chk = use_chk
        self.compute_dtype: Optional[torch.dtype] = None   # None: follow the input dtype; torch.bfloat16: tcgen05 path
        self.img_size = img_size
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        heads = num_heads
        depth = list(depth)

        self.stage1_conv_embed = nn.Sequential(
            nn.Conv2d(in_chans, embed_dim, 7, 4, 2),
            nn.Identity(),                       # the reference's einops Rearrange('b c h w -> b (h w) c') (no params)
            nn.LayerNorm(embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, int(np.sum(depth)))]

        def blocks(dim, heads_i, reso, split, first, n, last=False):
            return nn.ModuleList([
                CSWinBlock(dim=dim, num_heads=heads_i, reso=reso, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                           qk_scale=qk_scale, split_size=split, drop=drop_rate, attn_drop=attn_drop_rate,
                           drop_path=dpr[first + i], norm_layer=norm_layer, last_stage=last) for i in range(n)])

        off = [0, depth[0], depth[0] + depth[1], depth[0] + depth[1] + depth[2]]
        d = embed_dim
        self.stage1 = blocks(d, heads[0], img_size // 4, split_size[0], off[0], depth[0])
        self.merge1 = Merge_Block(d, d * 2)
        self.stage2 = blocks(d * 2, heads[1], img_size // 8, split_size[1], off[1], depth[1])
        self.merge2 = Merge_Block(d * 2, d * 4)
        self.stage3 = blocks(d * 4, heads[2], img_size // 16, split_size[2], off[2], depth[2])
        self.merge3 = Merge_Block(d * 4, d * 8)
        self.stage4 = blocks(d * 8, heads[3], img_size // 32, split_size[-1], off[3], depth[-1], last=True)
        self.norm = norm_layer(d * 8)

        self.stage_up4 = blocks(d * 8, heads[3], img_size // 32, split_size[-1], off[3], depth[-1], last=True)
        self.upsample4 = CARAFE(d * 8, d * 4)
        self.concat_linear4 = nn.Linear(512, 256)          # hard-coded in the reference (:404) -> embed_dim is 64
        self.stage_up3 = blocks(d * 4, heads[2], img_size // 16, split_size[2], off[2], depth[2])
        self.upsample3 = CARAFE(d * 4, d * 2)
        self.concat_linear3 = nn.Linear(256, 128)
        self.stage_up2 = blocks(d * 2, heads[1], img_size // 8, split_size[1], off[1], depth[1])
        self.upsample2 = CARAFE(d * 2, d)
        self.concat_linear2 = nn.Linear(128, 64)
        self.stage_up1 = blocks(d, heads[0], img_size // 4, split_size[0], off[0], depth[0])
        self.upsample1 = CARAFE4(d, 64)
        self.norm_up = norm_layer(embed_dim)
        self.output = nn.Conv2d(in_channels=embed_dim, out_channels=self.num_classes, kernel_size=1, bias=False)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- pieces -------------------------------------------------------------------------------
    def _stem(self, x: Tensor, dt: torch.dtype) -> Tensor:
        conv, ln = self.stage1_conv_embed[0], self.stage1_conv_embed[2]
        B = x.shape[0]
        K = conv.weight[0].numel()
        Kp = (K + 7) // 8 * 8                               # row pitch multiple of 16 B for the bf16 TMA path
        if ag.needs_grad(*self.stage1_conv_embed.parameters()):
            col = ops.im2col_nchw(x, 7, 7, 4, 2, Kp, dt)                  # the network input needs no gradient
            wk = torch.nn.functional.pad(conv.weight.reshape(conv.weight.shape[0], -1), (0, Kp - K))
            y = ag.LayerNormFn.apply(ag.linear(col, wk, conv.bias), ln.weight, ln.bias, ln.eps)
            return y.view(B, -1, y.shape[-1])
        if (FUSE_STEM and dt == torch.bfloat16 and x.is_cuda and tuple(conv.weight.shape) == (64, 3, 7, 7) and conv.stride == (4, 4)
                and conv.padding == (2, 2) and conv.bias is not None):
            # conv + token layout + LayerNorm + row statistics in one tcgen05 launch (csrc/stem_tc.cu)
            f32 = torch.float32
            y = ops.stem_fused(x, self._w("stem.w192", conv.weight, dt, lambda t: torch.nn.functional.pad(t.reshape(t.shape[0], -1), (0, 192 - K))),
                               self._w("stem.b32", conv.bias, f32), self._w("stem.n.w32", ln.weight, f32), self._w("stem.n.b32", ln.bias, f32), ln.eps)
            if y is not None:
                return y
        col = ops.im2col_nchw(x, 7, 7, 4, 2, Kp, dt)
        wk = self._w("stem.w", conv.weight, dt, lambda t: torch.nn.functional.pad(t.reshape(t.shape[0], -1), (0, Kp - K)))
        y = ops.linear(col, wk, self._w("stem.b", conv.bias, dt))
        y = ops.layernorm_with_row_stats(y, self._w("stem.n.w", ln.weight, dt), self._w("stem.n.b", ln.bias, dt), ln.eps)
        return _view_keep_stats(y, (B, -1, y.shape[-1]))

    def _skip_linear(self, lin: nn.Linear, skip: Tensor, x: Tensor, key: str) -> Tensor:
        if ag.needs_grad(skip, x, lin.weight):
            return ag.linear(skip, lin.weight, lin.bias, a2=x)
        dt = x.dtype
        if dt == torch.bfloat16:                             # the epilogue also emits the row statistics the next block folds
            y, st = ops.linear(skip, self._w(key + ".w", lin.weight, dt), self._w(key + ".b", lin.bias, dt), a2=x, want_stats=True)
            y._cswin_stats = st
            return y
        return ops.linear(skip, self._w(key + ".w", lin.weight, dt), self._w(key + ".b", lin.bias, dt), a2=x)

    def _ln(self, ln: nn.LayerNorm, x: Tensor, key: str) -> Tensor:
        if ag.needs_grad(x, ln.weight):
            return ag.LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps)
        dt = x.dtype
        return ops.layernorm_with_row_stats(x, self._w(key + ".w", ln.weight, dt), self._w(key + ".b", ln.bias, dt), ln.eps)

    def forward_features(self, x: Tensor) -> Tensor:
        dt = self.compute_dtype or x.dtype
        if dt not in (torch.float32, torch.bfloat16):
            raise TypeError(f"compute dtype must be float32 or bfloat16, got {dt}")
        x = self._stem(x, dt)
        x = run_stage(self.stage1, x)
        self.x1 = x
        x = self.merge1(x)
        x = run_stage(self.stage2, x)
        self.x2 = x
        x = self.merge2(x)
        x = run_stage(self.stage3, x)
        self.x3 = x
        x = self.merge3(x)
        x = run_stage(self.stage4, x)
        return self._ln(self.norm, x, "norm")

    def forward_up_features(self, x: Tensor) -> Tensor:
        x = run_stage(self.stage_up4, x)
        x = self._skip_linear(self.concat_linear4, self.x3, self.upsample4(x), "cl4")
        x = run_stage(self.stage_up3, x)
        x = self._skip_linear(self.concat_linear3, self.x2, self.upsample3(x), "cl3")
        x = run_stage(self.stage_up2, x)
        x = self._skip_linear(self.concat_linear2, self.x1, self.upsample2(x), "cl2")
        x = run_stage(self.stage_up1, x)
        return self._ln(self.norm_up, x, "norm_up")

    def up_x4(self, x: Tensor, logits_dtype: Optional[torch.dtype] = None, want_logits: bool = True,
              want_labels: bool = False, out_logits: Optional[Tensor] = None, out_labels: Optional[Tensor] = None):
        """CARAFE4 + output conv with the two 1x1 maps folded (cswin_unet.py:536-544).  Returns NCHW logits
        (B, classes, 4H, 4W); with want_labels also / only the uint8 arg-max label map (utils.py:73-75)."""
        B, L, Cn = x.shape
        H = W = int(round(L ** 0.5))
        up = self.upsample1
        if (ag.needs_grad(x, self.output.weight, *up.parameters()) and FOLD_TRAIN_HEAD
                and ops.carafe_head_bwd_supported(self.num_classes, up.up_factor)):
            # training, folded like inference: W_f = W_output W_out and b_f = W_output b_out are formed ON THE TAPE (two tiny
            # fp32 matmuls), so autograd turns the head kernel's d z / d bias into the gradients of both 1x1 convs; the
            # (B, 16 L, 64) activation of the unfolded structure and its backward never exist
            nc = self.num_classes
            enc = up._kernel_logits_tape(x, H, W)
            wo2 = self.output.weight.reshape(nc, -1).float()
            wf = torch.nn.functional.pad(wo2 @ up.out.weight.reshape(up.out.weight.shape[0], -1).float(), (0, 0, 0, 16 - nc))
            bf = wo2 @ up.out.bias.float()
            z = ag.linear(x, wf, None)                                          # (B, L, 16); columns >= classes are 0
            lg = ag.CarafeHeadFn.apply(enc, z.view(B * L, -1), bf, B, H, W, up.up_factor, nc)   # fp32 NCHW
            return lg.to(logits_dtype) if logits_dtype is not None else lg
        if ag.needs_grad(x, self.output.weight, *up.parameters()):
            # training, CSWIN_UNFOLDED_TRAIN_HEAD=1 or an unsupported class count: the reference's unfolded structure (CARAFE4 with its own `out` conv, then the `output` conv) so that
            # every parameter receives its gradient from a native kernel; logits come back token-major and are viewed NCHW
            y = up(x)                                                           # (B, 16 L, 64)
            lg = ag.linear(y, self.output.weight.reshape(self.output.weight.shape[0], -1), None)
            lg = lg.view(B, H * up.up_factor, W * up.up_factor, -1).permute(0, 3, 1, 2)
            return lg.to(logits_dtype) if logits_dtype is not None else lg
        dt = x.dtype
        wo = self.output.weight
        npad = 16 if self.num_classes <= 16 else self.num_classes             # 16-column rows: vector loads in the head kernel

        def fold_w(o, u):
            w = o.reshape(o.shape[0], -1).float() @ u.reshape(u.shape[0], -1).float()
            return torch.nn.functional.pad(w, (0, 0, 0, npad - w.shape[0]))
        wf = self._w("head.w", (wo, up.out.weight), dt, fold_w)
        bf = self._w("head.b", (wo, up.out.bias), dt, lambda o, b: o.reshape(o.shape[0], -1).float() @ b.float())
        enc, z = up._kernel_logits(x, H, W, wf, z_key="head")                   # z (B, L, npad) = folded head map of x; columns >= classes are 0
        if self.num_classes <= 16:
            logits, labels = ops.carafe_head(enc, z.flatten(0, 1), bf, B, H, W, up.up_factor, want_logits=want_logits,
                                             want_labels=want_labels, logits_dtype=logits_dtype or dt,
                                             n_classes=self.num_classes, out_logits=out_logits, out_labels=out_labels)
        else:                                                                   # generic re-assembly, labels via torch
            logits = ops.carafe_reassemble(enc, z.flatten(0, 1), bf, B, H, W, up.up_factor, nchw_out=True,
                                           out_dtype=logits_dtype or dt)
            labels = logits.argmax(1).to(torch.uint8) if want_labels else None
            if out_logits is not None and want_logits:
                out_logits.copy_(logits); logits = out_logits
            if out_labels is not None and want_labels:
                out_labels.copy_(labels); labels = out_labels
        if want_labels:
            return (logits, labels) if want_logits else labels
        return logits

    def _run(self, x: Tensor, logits_dtype, want_logits: bool, want_labels: bool, out_logits=None, out_labels=None):
        x = self.forward_features(x)
        x = self.forward_up_features(x)
        return self.up_x4(x, logits_dtype=logits_dtype, want_logits=want_logits, want_labels=want_labels,
                          out_logits=out_logits, out_labels=out_labels)

    def _side_streams(self, device, n: int):
        key = "_streams_%s" % (device,)
        st = self.__dict__.get(key)
        if st is None or len(st) < n:
            st = [torch.cuda.Stream(device) for _ in range(n)]
            self.__dict__[key] = st
        return st[:n]

    def _run_split(self, x: Tensor, logits_dtype, want_logits: bool, want_labels: bool):
        """`_run` over N_STREAMS sub-batches on concurrent streams (inference only; see N_STREAMS)."""
        B = x.shape[0]
        n = min(N_STREAMS, B)
        if (n <= 1 or not x.is_cuda or self.num_classes > 16
                or ag.needs_grad(x, *self.parameters())):
            return self._run(x, logits_dtype, want_logits, want_labels)
        S = self.img_size
        dt = logits_dtype or self.compute_dtype or x.dtype
        logits = torch.empty((B, self.num_classes, S, S), dtype=dt, device=x.device) if want_logits else None
        labels = torch.empty((B, S, S), dtype=torch.uint8, device=x.device) if want_labels else None
        cur = torch.cuda.current_stream(x.device)
        side = self._side_streams(x.device, n - 1)
        fork = torch.cuda.Event()
        fork.record(cur)
        for i in range(n):
            lo, hi = B * i // n, B * (i + 1) // n
            s = cur if i == 0 else side[i - 1]
            if i:
                s.wait_event(fork)
            with torch.cuda.stream(s):
                self._run(x[lo:hi], dt, want_logits, want_labels, None if logits is None else logits[lo:hi],
                          None if labels is None else labels[lo:hi])
        for s in side:
            cur.wait_stream(s)
        if want_labels:
            return (logits, labels) if want_logits else labels
        return logits

    def forward(self, x: Tensor) -> Tensor:
        # logits leave in the dtype of the INPUT: an fp32 image through the bf16 compute path (compute_dtype) gets fp32 logits
        # straight from the head kernel's fp32 accumulators instead of a final rounding to bf16 (|logit| ~ 25 -> 0.1 absolute)
        in_dt = x.dtype
        return self._run_split(x, in_dt if in_dt in (torch.float32, torch.bfloat16) else None, True, False)

    @torch.no_grad()
    def predict_labels(self, x: Tensor) -> Tensor:
        """argmax(softmax(forward(x)), 1) as uint8 (B, H, W), computed inside the head kernel (no logits written)."""
        return self._run_split(x, None, False, True)


class CSwinUnet(nn.Module):
    """Counterpart of networks/vision_transformer.py:17-43 without the yacs dependency: takes either a config
    object with the reference's attribute tree or plain keyword hyper-parameters (cswin_tiny_224_lite defaults:
    configs/cswin_tiny_224_lite.yaml:4-10, config.py:49-66).  Unlike the reference it does not write
    `cswin_unet.pth` into the CWD at construction."""

    def __init__(self, config=None, img_size=224, num_classes=9, zero_head=False, vis=False, **kw):
        super().__init__()
        self.num_classes = num_classes
        self.zero_head = zero_head
        self.config = config
        if config is not None:
            c = config.MODEL.CSWIN
            hp = dict(img_size=config.DATA.IMG_SIZE, patch_size=c.PATCH_SIZE, in_chans=c.IN_CHANS, embed_dim=c.EMBED_DIM,
                      depth=c.DEPTH, split_size=c.SPLIT_SIZE, num_heads=c.NUM_HEADS, mlp_ratio=c.MLP_RATIO,
                      qkv_bias=c.QKV_BIAS, qk_scale=c.QK_SCALE, drop_rate=config.MODEL.DROP_RATE,
                      drop_path_rate=config.MODEL.DROP_PATH_RATE)
        else:
            hp = dict(img_size=img_size, patch_size=4, in_chans=3, embed_dim=64, depth=[1, 2, 9, 1],
                      split_size=[1, 2, 7, 7], num_heads=[2, 4, 8, 16], mlp_ratio=4., qkv_bias=True, qk_scale=None,
                      drop_rate=0., drop_path_rate=0.2)
        hp.update(kw)
        self.cswin_unet = CSWinTransformer(num_classes=self.num_classes, **hp)

    def forward(self, x: Tensor) -> Tensor:
        if x.size()[1] == 1:
            x = x.repeat(1, 3, 1, 1)
        return self.cswin_unet(x)


def cswin_tiny_224(num_classes: int = 9, img_size: int = 224, **kw) -> CSWinTransformer:
    hp = dict(img_size=img_size, patch_size=4, in_chans=3, num_classes=num_classes, embed_dim=64, depth=[1, 2, 9, 1],
              split_size=[1, 2, 7, 7], num_heads=[2, 4, 8, 16], mlp_ratio=4., qkv_bias=True, qk_scale=None,
              drop_rate=0., drop_path_rate=0.2)
    hp.update(kw)
    return CSWinTransformer(**hp)
