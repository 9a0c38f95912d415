"""Slice-inference engine: the call a user of the reference's `test_single_volume` loop makes, B200-style.

The reference evaluates one slice at a time (utils.py:61-90: host zoom -> `.cuda()` -> forward -> argmax -> `.cpu()`).
Slices are independent, so the engine batches them, keeps the forward in a CUDA graph, and pipelines
host->device copy / forward / device->host copy of consecutive batches; `inflight` (default CSWIN_INFLIGHT = 3) batches run
their forwards CONCURRENTLY, each on its own stream with its own captured graph and buffers: at batch 24 every kernel of the
forward is at most one or two waves and bound by its own latency chain, so a second / third independent forward fills the
SMs the first leaves idle (measured on B200: 19.2k -> 23.4k -> 25.1k slices/s with 1 / 2 / 3 forwards in flight,
profiles/r02_concurrent_forwards.log):

    eng = SliceEngine(model, batch=24)                 # model: CSWinTransformer on a CUDA device
    labels = eng.predict(host_batch)                   # (B,3|1,H,W) float32 CPU tensor -> (B,H,W) uint8 CPU tensor
    for labels in eng.predict_stream(iter_of_batches): ...   # pipelined; results come back in order

Only a uint8 label map leaves the GPU (the arg-max is taken inside the head kernel).  Multi-GPU: one engine per
process / GPU over a shard of the slices (`shard_slices`); there is no collective on this path.
"""
from __future__ import annotations

import os

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


def shard_slices(n_slices: int, world: int, rank: int) -> range:
    """Contiguous shard [lo, hi) of slice indices for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n_slices, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


class SliceEngine:
    def __init__(self, model, batch: int, img_size: Optional[int] = None, in_chans: int = 3,
                 compute_dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None,
                 inflight: Optional[int] = None):
        self.model = model.eval()
        self.batch = batch
        self.size = img_size or model.img_size
        self.in_chans = in_chans
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("SliceEngine needs a CUDA device: cswin_unet_b200 has no CPU path")
        model.compute_dtype = compute_dtype
        self.inflight = max(1, int(os.environ.get("CSWIN_INFLIGHT", "3")) if inflight is None else int(inflight))
        self.streams = {k: torch.cuda.Stream(self.device) for k in ("h2d", "compute", "d2h")}
        shape = (batch, in_chans, self.size, self.size)
        self.slots = []
        with torch.cuda.device(self.device), torch.no_grad():
            # slots = forwards in flight + one batch in its host->device copy + one being drained / collected by the host: with
            # only `inflight` slots the copy-in of batch k+ns starts after batch k has left the device, and one of the compute
            # streams idles for a copy time per step (batch 96: 25.2k -> see profiles/r02_bench_*.json)
            n_slots = 2 if self.inflight == 1 else self.inflight + 2
            compute = [self.streams["compute"]] + [torch.cuda.Stream(self.device) for _ in range(self.inflight - 1)]
            for i in range(n_slots):
                x = torch.zeros(shape, dtype=torch.float32, device=self.device)
                slot = {"x": x, "host_out": torch.empty((batch, self.size, self.size), dtype=torch.uint8).pin_memory(),
                        "ev_in": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                        "ev_free": torch.cuda.Event(),
                        "stream": compute[i % self.inflight]}        # exactly `inflight` forwards run concurrently
                self.slots.append(slot)
        self._n = 0
        self._resample_bufs = {}                                 # (H, W) -> device / pinned buffers of predict_volume(resample='gpu')
        self._capture()

    def _weights_signature(self):
        """(parameter epoch, sum of torch versions): changes whenever the model's weights were written — by the native SGD
        (raw pointers, modules.bump_param_epoch) or by torch (load_state_dict, optimizers, TPGM's .data.copy_ ...)."""
        from .modules import param_epoch
        return param_epoch(), sum(p._version for p in self.model.parameters())

    def _capture(self) -> None:
        """(Re-)capture the forward graphs.  The graphs bake pointers to weight tensors DERIVED from the parameters (bf16 casts,
        folded LayerNorm / head matrices): they must be rebuilt when the parameters change, which `refresh_if_stale` does."""
        torch.cuda.synchronize(self.device)
        # CTAs of concurrent launches must fit next to each other on an SM: cap the tcgen05 Linear's operand ring while capturing
        # (process-wide option read when a launch is enqueued, i.e. baked into the graphs); restored afterwards
        from . import _lib
        # (batch <= 32 only: there every launch is <= 2 waves; at batch 96 the cap costs 4 %, measured)
        cap_default = "100" if self.batch <= 32 else "0"
        _lib.set_option(_lib.OPT_GEMM_SMEM_CAP_KB, int(os.environ.get("CSWIN_INFLIGHT_SMEM_CAP_KB", cap_default)) if self.inflight > 1 else 0)
        try:
            self._capture_graphs()
        finally:
            _lib.set_option(_lib.OPT_GEMM_SMEM_CAP_KB, 0)
        self._sig = self._weights_signature()

    def _capture_graphs(self) -> None:
        with torch.cuda.device(self.device), torch.no_grad():
            for slot in self.slots:                              # one graph per slot (static input / output buffers), captured
                cs = slot["stream"]                              # on the stream it will be replayed on; concurrent replays must
                cs.wait_stream(torch.cuda.current_stream(self.device))   # not share a memory pool, so each graph owns its own
                with torch.cuda.stream(cs):
                    for _ in range(2):                           # warm-up: derive cached weights, load kernels
                        self.model.predict_labels(slot["x"])
                cs.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cs):
                    slot["y"] = self.model.predict_labels(slot["x"])
                slot["graph"] = g
                with torch.cuda.stream(cs):                      # first replay uploads the graph (several ms for ~160 nodes): pay it here,
                    g.replay()                                   # not inside the first timed / user-visible batch of every slot
                cs.synchronize()

    def refresh_if_stale(self) -> bool:
        """Re-capture when the model's weights changed since the graphs were built (train -> validate loops of the reference:
        universal_train.py:644/868).  Called by every entry point; costs one integer comparison when nothing changed."""
        if self._weights_signature() == self._sig:
            return False
        self.model.eval()
        self._capture()
        return True

    # ---- one batch ------------------------------------------------------------------------------------
    def _submit(self, host_x: Tensor) -> dict:
        self.refresh_if_stale()
        slot = self.slots[self._n % len(self.slots)]
        self._n += 1
        n = host_x.shape[0]
        if n > self.batch or tuple(host_x.shape[2:]) != (self.size, self.size):
            raise ValueError(f"expected at most {self.batch} slices of {self.size}x{self.size}, got {tuple(host_x.shape)}")
        h2d, cs, d2h = self.streams["h2d"], slot["stream"], self.streams["d2h"]
        with torch.cuda.stream(h2d):
            h2d.wait_event(slot["ev_done"])                        # previous forward on this slot has consumed x
            if host_x.shape[1] == 1 and self.in_chans == 3:
                # 1 -> 3 channel repeat (vision_transformer.py:40-41; test_single_volume feeds single-channel slices, utils.py:69-71):
                # ONE channel crosses PCIe, the repeat is a device-side broadcast copy (a third of the host->device bytes)
                if "x1" not in slot:
                    slot["x1"] = torch.empty((self.batch, 1, self.size, self.size), dtype=torch.float32, device=self.device)
                slot["x1"][:n].copy_(host_x, non_blocking=True)
                slot["x"][:n].copy_(slot["x1"][:n].expand(-1, 3, -1, -1))
            else:
                slot["x"][:n].copy_(host_x, non_blocking=True)
            slot["ev_in"].record(h2d)
        with torch.cuda.stream(cs):
            cs.wait_event(slot["ev_in"])
            cs.wait_event(slot["ev_out"])                          # previous D2H of this slot's y has finished
            slot["graph"].replay()
            slot["ev_done"].record(cs)
        with torch.cuda.stream(d2h):
            d2h.wait_event(slot["ev_done"])
            slot["host_out"].copy_(slot["y"], non_blocking=True)
            slot["ev_out"].record(d2h)
        slot["n"] = n
        return slot

    def _collect(self, slot: dict) -> Tensor:
        slot["ev_out"].synchronize()
        return slot["host_out"][:slot["n"]].clone()

    def predict(self, host_x: Tensor) -> Tensor:
        return self._collect(self._submit(host_x))

    def predict_stream(self, batches: Iterable[Tensor]) -> Iterator[Tensor]:
        """Pipelined: up to len(slots) batches are between their host->device copy and their device->host copy; their forwards
        run concurrently on the slots' streams."""
        pending: List[dict] = []
        for hx in batches:
            if len(pending) == len(self.slots):
                yield self._collect(pending.pop(0))
            pending.append(self._submit(hx))
        while pending:
            yield self._collect(pending.pop(0))

    def bytes_per_batch(self, host_chans: Optional[int] = None) -> Tuple[int, int]:
        """(host->device, device->host) bytes of one full batch; host_chans = channels of the host tensor (1: repeated on the device)."""
        return self.batch * (host_chans or self.in_chans) * self.size * self.size * 4, self.batch * self.size * self.size


STAGING_THREADS = max(1, int(os.environ.get("CSWIN_STAGING_THREADS", "4")))   # host threads copying a batch into its pinned staging buffer
_STAGING_POOL = None


def _staged_copy(dst: Tensor, src: Tensor) -> None:
    """dst[:] = src (host -> pinned host, both contiguous along dim 0) split over STAGING_THREADS threads: one thread moves ~10 GB/s,
    which made this memcpy the longest step of predict_volume(resample='gpu'); torch's copy_ releases the GIL."""
    global _STAGING_POOL
    n = dst.shape[0]
    if STAGING_THREADS == 1 or n < 2 * STAGING_THREADS:
        dst.copy_(src)
        return
    if _STAGING_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _STAGING_POOL = ThreadPoolExecutor(max_workers=STAGING_THREADS, thread_name_prefix="cswin-staging")
    cuts = [n * i // STAGING_THREADS for i in range(STAGING_THREADS + 1)]
    futs = [_STAGING_POOL.submit(dst[a:b].copy_, src[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    for f in futs:
        f.result()


REGISTER_VOLUME = os.environ.get("CSWIN_VOLUME_REGISTER", "0") == "1"     # measured: page-locking a 150 MB volume costs as much as the staging memcpy it saves


def _predict_volume_gpu(engine: "SliceEngine", image, rng: range):
    """predict_volume with both resampling steps on the GPU (cswin_zoom_cubic_fwd / cswin_zoom_nearest_u8): per batch the raw
    slices go host -> device, are zoomed straight into the engine's input buffer (all three channel planes), the captured
    forward graph runs, and the label map is zoomed back before it leaves the device.  Same arithmetic as scipy (float64
    spline prefilter and interpolation, scipy's edge rule), so the label maps equal the host-resampled ones."""
    import numpy as np
    from . import ops
    D, H, W = image.shape
    P, Bt = engine.size, engine.batch
    dev = engine.device
    key = (H, W)
    ns = len(engine.slots)
    buf = engine._resample_bufs.get(key)
    if buf is None:
        with torch.cuda.device(dev):
            buf = {"raw": [torch.empty((Bt, H, W), dtype=torch.float32, device=dev) for _ in range(ns)],
                   "work": [torch.empty(Bt * H * W, dtype=torch.float64, device=dev) for _ in range(ns)],
                   "lab": [torch.empty((Bt, H, W), dtype=torch.uint8, device=dev) for _ in range(ns)],
                   "host_in": [torch.empty((Bt, H, W), dtype=torch.float32).pin_memory() for _ in range(ns)],
                   "ev": [torch.cuda.Event() for _ in range(ns)]}
        engine._resample_bufs[key] = buf
    out = torch.empty((len(rng), H, W), dtype=torch.uint8).pin_memory()
    engine.refresh_if_stale()
    cur = torch.cuda.current_stream(dev)
    for slot in engine.slots:
        slot["stream"].wait_stream(cur)
    idx = list(rng)
    # page-lock the caller's volume for the duration of the call: the slices then go host -> device straight from it (no
    # staging memcpy); if registration is refused the pinned staging buffers are used
    vol_t = torch.from_numpy(image)
    rt = torch.cuda.cudart()
    registered = REGISTER_VOLUME and int(rt.cudaHostRegister(vol_t.data_ptr(), vol_t.numel() * 4, 0)) == 0
    try:
        with torch.no_grad():
            for bi, i0 in enumerate(range(0, len(idx), Bt)):
                sl = idx[i0:i0 + Bt]
                n = len(sl)
                k = bi % ns                                          # batches alternate over the slots; each slot's chain (copy in,
                slot = engine.slots[k]                               # zoom, forward, zoom back, copy out) runs on its own stream
                with torch.cuda.stream(slot["stream"]):
                    src = vol_t[sl[0]:sl[-1] + 1]
                    if not registered:
                        buf["ev"][k].synchronize()                   # the pinned staging buffer of this slot is free again
                        _staged_copy(buf["host_in"][k][:n], src)
                        src = buf["host_in"][k][:n]
                    buf["raw"][k][:n].copy_(src, non_blocking=True)
                    buf["ev"][k].record(slot["stream"])
                    ops.zoom_cubic(buf["raw"][k][:n], (P, P), out=slot["x"], work=buf["work"][k])
                    slot["graph"].replay()
                    ops.zoom_nearest_u8(slot["y"][:n].contiguous(), (H, W), out=buf["lab"][k])
                    out[i0:i0 + n].copy_(buf["lab"][k][:n], non_blocking=True)
        for slot in engine.slots:
            slot["stream"].synchronize()
    finally:
        if registered:
            for slot in engine.slots:
                slot["stream"].synchronize()
            rt.cudaHostUnregister(vol_t.data_ptr())
    return out.numpy()                                               # (a view of the pinned result buffer: no extra copy)


def predict_volume(engine: "SliceEngine", image, order_in: int = 3, shard: Optional[Tuple[int, int]] = None,
                   resample: str = "scipy"):
    """The slice loop of `test_single_volume` (utils.py:61-80) on the engine: every slice of the (D, H, W) float volume is
    resized to the network resolution with scipy `zoom(order=3)` on the host (as the reference does), segmented in
    batches, and the label map is resized back with `zoom(order=0)`.  `shard=(rank, world)` restricts the work to this
    rank's contiguous slice range; returns (labels uint8 (n_local, H, W), range).
    `resample="gpu"` does both zooms on the device instead (same arithmetic, ~15 ms of host time per 512^2 slice saved)."""
    import numpy as np
    image = np.ascontiguousarray(np.asarray(image, dtype=np.float32))
    D, H, W = image.shape
    rng = shard_slices(D, shard[1], shard[0]) if shard else range(D)
    P = engine.size
    if resample == "gpu" and order_in == 3 and len(rng) > 0 and (H, W) != (P, P):      # (same size: the reference does not zoom)
        return _predict_volume_gpu(engine, image, rng), rng
    if resample not in ("scipy", "gpu"):
        raise ValueError(f"resample must be 'scipy' or 'gpu', got {resample!r}")
    from scipy.ndimage import zoom
    resized = np.empty((len(rng), 1, P, P), np.float32)
    for j, d in enumerate(rng):
        sl = image[d]
        resized[j, 0] = zoom(sl, (P / H, P / W), order=order_in) if (H, W) != (P, P) else sl
    host = torch.from_numpy(resized)
    outs = list(engine.predict_stream(host[i:i + engine.batch] for i in range(0, len(rng), engine.batch)))
    lab = torch.cat(outs, 0).numpy() if outs else np.zeros((0, P, P), np.uint8)
    if (H, W) != (P, P):
        lab = np.stack([zoom(l, (H / P, W / P), order=0) for l in lab]) if len(lab) else np.zeros((0, H, W), np.uint8)
    return lab, rng
