#!/usr/bin/env python
"""bench.py — CSWin-UNet-tiny 224^2 slices/sec on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference|reference-cuda]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

native arm  : a "step" is one bf16 forward of cswin_tiny_224_lite over one batch of 96 synthetic 3x224x224 slices
              per GPU through the native kernels (libcswin_b200.so), replayed as a CUDA graph, 2 such forwards in flight.
              (Slice inference has no batch of its own — test_single_volume, utils.py:61-90, walks the ~150 slices of a
              volume — so the engine's batch is a free parameter: 96 x 2 in flight is where the B200 saturates; the batch-24
              figures, single and 3 in flight, stay in the line as `extra.forward_batch24`; the train step is batch 24.)
              `value`  = slices/s with the inputs resident in HBM (8 rotating batches = 462 MB > L2),
              `e2e`    = same metric through the public nn.Module call with HOST (pinned) inputs: H2D copy of the
                         batch, forward, argmax label map, D2H of the label map, every step,
              `roofline` = the dominant kernel family (tcgen05 Linear incl. the implicit-GEMM convs): algorithmic flops / CUDA-event
                         time vs the measured bf16 peak; `roofline_attention` = fused LePE attention family: algorithmic bytes / time
                         vs the measured HBM peak; `roofline_model` = the whole forward,
              `cpu_baseline` = the unmodified reference (baseline/_ref; the oracle port if it did not travel) on this box's host cores.
reference arm: the reference's own CPU path on all host cores, same metric / config; rank 0 only.  It is the UNMODIFIED
              reference (baseline/_ref/networks/cswin_unet.py, copied verbatim by build(); kind "reference") when that copy
              travelled with the snapshot, else the oracle port (kind "port").
reference-cuda: the unmodified reference in eager PyTorch on the same B200 (fp32 as the reference runs it, and bf16 autocast)
              — the comparator SURVEY 8d / BASELINE.md 3 ask for; also reported inside the native line as `reference_cuda`.
Slices are independent, so N GPUs = N replicas each running its own batches: weak scaling, no collective on the
data path; timing = max over ranks of the device time, bracketed by barrier + synchronize.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CSWin-UNet-tiny 224^2 slices/sec (bf16 fwd)"
UNIT = "slices/s"
BATCH = 24                         # BASELINE configs[2]: the train step's batch, and the latency point of the forward
CPU_CHUNK = 24                     # sub-batch the CPU reference arm processes a step in (its faster batching on the host cores)
FWD_BATCH = 96                     # slices per forward step of the headline (throughput point, see the module docstring)
GFLOP_PER_SLICE_FWD = 10.028       # BASELINE.md section 2 (FlopCounterMode on the unmodified reference)
# dram__bytes_read.sum + dram__bytes_write.sum of lepe_attn_fwd_tc_kernel, ncu --set full, batch 24, summed over the 26 launches
# of one forward (2 x 28.95 MB + 4 x 14.50 MB + 18 x 7.28 MB + 2 x 3.67 MB read, ~0 written inside the kernel: the output
# stays in L2): profiles/r01_ncu_full_summary_v2.txt.  Reads equal the algorithmic q,k,v bytes exactly (no re-reads).
ATTN_DRAM_TRAFFIC_PER_FORWARD = 254_194_000
# the same at batch 96 (profiles/r02_ncu_full_b96_summary.txt): 2 x (228.83 + 30.02) + 4 x (57.85 + 1.51) + 18 x 28.95 + 2 x 14.50 MB: the
# stage-1 output (154 MB per launch with q, k, v) no longer stays in L2; algorithmic bytes are 1,348.7 MB
ATTN_DRAM_TRAFFIC = {24: ATTN_DRAM_TRAFFIC_PER_FORWARD, 96: 1_305_240_000}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML during the timed region."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def shutdown_dist(timeout_s: float = 20.0) -> None:
    """destroy_process_group() with a watchdog: NCCL teardown after CUDA-graph-captured collectives has been seen to block;
    the result line is already printed and flushed, so a stuck teardown ends the process instead of hanging the launcher."""
    import gc
    gc.collect()
    torch.cuda.synchronize()
    done = threading.Event()

    def watch():
        if not done.wait(timeout_s):
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
    threading.Thread(target=watch, daemon=True).start()
    try:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    finally:
        done.set()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------
# CPU oracle legs
# --------------------------------------------------------------------------------------------------
def oracle_model():
    from oracle import cswin_oracle as O          # test infrastructure; used here only as the CPU baseline
    from cswin_unet_b200 import synth
    shapes = O.state_dict_shapes()
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}
    return O, sd


def cpu_forward_rate(budget_s: float, batch: int, min_iters: int = 2):
    from cswin_unet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = reference_model("cpu")
    if ref is not None:
        kind, fwd = "reference", ref
    else:
        O, sd = oracle_model()
        kind, fwd = "port", (lambda t: O.cswin_unet_forward(sd, t))
    x = torch.from_numpy(synth.synth_image_batch(batch, 3, 224, seed=0, kind="ct"))
    chunks = list(x.split(CPU_CHUNK))                          # sub-batches of 24: the reference's faster batching on the host
    with torch.no_grad():
        fwd(x[:2])                                             # warm-up
        t0 = time.perf_counter(); n = 0
        while n < min_iters or (time.perf_counter() - t0) < budget_s:
            for c in chunks:
                fwd(c)
            n += 1
        dt = time.perf_counter() - t0
    return n * batch / dt, cores, n, dt, kind


def cpu_extra_rates(budget_s: float = 8.0):
    """BASELINE.md 3 plan, items 2: the reference on the host cores at batch 1 (eval forward) and one train step at batch 24
    (trainer.py:42-63 loop).  Bounded: a few forwards, one warm-up + >= 1 timed train step.  None if baseline/_ref is absent."""
    from cswin_unet_b200 import synth
    ref = reference_model("cpu")
    if ref is None:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cores": cores}
    x1 = torch.from_numpy(synth.synth_image_batch(1, 3, 224, seed=0, kind="ct"))
    with torch.no_grad():
        ref(x1)
        t0 = time.perf_counter(); n = 0
        while n < 3 or (time.perf_counter() - t0) < 1.0:
            ref(x1); n += 1
        out["forward_batch1"] = {"slices_per_s": n / (time.perf_counter() - t0), "ms_per_forward": 1e3 * (time.perf_counter() - t0) / n}
    ref.train()
    x = torch.from_numpy(synth.synth_image_batch(BATCH, 3, 224, seed=0, kind="ct"))
    y = torch.from_numpy(synth.synth_labels(BATCH, 224, 9, seed=0)).long()
    step = _reference_train_step_fn(ref, x, y, 9, None)
    step()
    t0 = time.perf_counter(); n = 0
    while n < 1 or (time.perf_counter() - t0) < budget_s - 3.0:
        step(); n += 1
    dt = time.perf_counter() - t0
    out[f"train_step_batch{BATCH}"] = {"slices_per_s": n * BATCH / dt, "ms_per_step": 1e3 * dt / n, "steps": n}
    return out


def reference_model(device="cpu"):
    """The unmodified reference model with the bench's synthetic weights, or None if baseline/_ref did not travel."""
    from baseline import ref_loader
    from cswin_unet_b200 import synth
    if not ref_loader.available():
        return None
    m = ref_loader.build_reference_model().eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    return m.to(device)


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from cswin_unet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = reference_model("cpu")
    if ref is not None:
        kind, what = "reference", "UNMODIFIED reference networks/cswin_unet.py (baseline/_ref), torch CPU fp32"
        fwd = ref
    else:
        O, sd = oracle_model()
        kind, what = "port", "CPU port of the reference PyTorch path (oracle/cswin_oracle.py), fp32"
        fwd = lambda x: O.cswin_unet_forward(sd, x)       # noqa: E731
    # bounded sample per step so that K+W steps end within minutes on any host: a batch of 96 slices is ~1.5 s on 16 cores
    sample = args.batch
    x = torch.from_numpy(synth.synth_image_batch(sample, 3, 224, seed=0, kind="ct"))
    # the step's slices go through the reference in sub-batches of 24: its faster batching on the host (one batch-96 forward runs at
    # 42-52 slices/s on 16 cores, four batch-24 forwards at 64-72; its own evaluation loop feeds ONE slice at a time: 13 slices/s)
    chunks = list(x.split(CPU_CHUNK))
    with torch.no_grad():
        for _ in range(args.warmup):
            for c in chunks:
                fwd(c)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for c in chunks:
                fwd(c)
        dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cswin_tiny_224_lite eval forward, {sample} slices per step (as {len(chunks)} forwards of {CPU_CHUNK}: the CPU's faster batching), "
                                   "3x224x224 synthetic slices, " + what},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} steps x {sample} slices ({len(chunks)} forwards of {CPU_CHUNK} each), torch CPU fp32, {cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def _reference_train_step_fn(ref, x, y, n_classes=9, autocast=None):
    """One step of the reference's own training loop (trainer.py:42-63: SGD momentum .9 / wd 1e-4, 0.4 CE + 0.6 Dice (utils.py:9-45,
    restated by the oracle and pinned against the reference's DiceLoss), zero_grad / backward / step) on a reference model."""
    from oracle import cswin_oracle as O          # the loss restatement only (baseline arm, not the product path)
    opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)

    def step():
        if autocast is None:
            out = ref(x)
        else:
            with autocast:
                out = ref(x)
        out = out.float()
        loss = 0.4 * torch.nn.functional.cross_entropy(out, y) + 0.6 * O.dice_loss(out, y, n_classes)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    return step


def reference_cuda_rates(dev, batches=(BATCH, 1, FWD_BATCH, 192), iters=10, warm=3):
    """Eager-CUDA throughput of the unmodified reference on this GPU (stock code path: cuBLAS / cuDNN / ATen kernels, no repo
    module): fp32 (how the reference runs, TF32 off as train.py:73-78 leaves it) and bf16 autocast; device-resident inputs,
    CUDA events.  Returns None when baseline/_ref did not travel."""
    from cswin_unet_b200 import synth
    ref = reference_model(dev)
    if ref is None:
        return None
    out = {"what": "unmodified reference (baseline/_ref/networks/cswin_unet.py), eager PyTorch on this GPU, eval forward, "
                   "device-resident inputs, CUDA events", "torch": torch.__version__}
    for B in batches:
        x = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=0, kind="ct")).to(dev)
        for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with torch.no_grad():
                def fwd():
                    if ctx is None:
                        return ref(x)
                    with ctx:
                        return ref(x)
                it_b = iters if B <= BATCH else 3                # batch 192 is ~150 ms per eager forward
                for _ in range(warm if B <= BATCH else 1):
                    fwd()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(it_b):
                    fwd()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / it_b
            out[f"batch{B}_{name}"] = {"slices_per_s": B / (ms * 1e-3), "ms_per_forward": ms}
    # the reference's train step (trainer.py:42-63) in eager PyTorch on this GPU, batch 24: the baseline of `train_step`
    try:
        ref.train()
        x = torch.from_numpy(synth.synth_image_batch(BATCH, 3, 224, seed=0, kind="ct")).to(dev)
        y = torch.from_numpy(synth.synth_labels(BATCH, 224, 9, seed=0)).to(dev).long()
        for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            step = _reference_train_step_fn(ref, x, y, 9, ctx)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out[f"train_step_batch{BATCH}_{name}"] = {"slices_per_s": BATCH / (ms * 1e-3), "ms_per_step": ms}
    except Exception as e:                                         # noqa: BLE001 — reported, the forward numbers stand
        out["train_step_error"] = f"{type(e).__name__}: {e}"[:200]
    del ref
    torch.cuda.empty_cache()
    return out


def run_reference_cuda(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    if not torch.cuda.is_available():
        emit({"impl": "reference-cuda", "unavailable": "no CUDA device"})
        return 0
    torch.cuda.set_device(local)
    r = reference_cuda_rates(torch.device("cuda", local), batches=tuple(dict.fromkeys((args.batch, BATCH, 1, 192))), iters=max(args.steps, 5), warm=max(args.warmup, 3))
    if r is None:
        emit({"impl": "reference-cuda", "unavailable": "baseline/_ref/networks/cswin_unet.py did not travel with the snapshot"})
        return 0
    v = r[f"batch{args.batch}_fp32"]
    emit({"impl": "reference-cuda", "metric": METRIC, "value": v["slices_per_s"], "unit": UNIT, "n_gpus": 1, "steps": max(args.steps, 5),
          "warmup": max(args.warmup, 3), "ms_per_step": v["ms_per_forward"], "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"cswin_tiny_224_lite eval forward, batch {args.batch}, unmodified reference in eager PyTorch on the GPU"},
          "detail": r})
    return 0


# --------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------
def attention_bytes_per_image():
    """Ideal traffic of the fused LePE attention per image, bf16: read q,k,v once + write out once."""
    per_block = {1: 3136 * 64, 2: 784 * 128, 3: 196 * 256, 4: 49 * 512}
    blocks = {1: 2, 2: 4, 3: 18, 4: 2}
    return sum(4 * per_block[s] * 2 * blocks[s] for s in per_block)       # = 14,049,280 B


def run_native(args):
    import cswin_unet_b200 as cw
    from cswin_unet_b200 import ops, synth

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cw.lib()                                                       # fail loudly if the extension is missing

    model = cw.cswin_tiny_224(num_classes=9).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    model = model.to(dev)
    model.compute_dtype = torch.bfloat16

    B = args.batch
    n_rot = 16 if B <= 32 else 8
    host = torch.from_numpy(synth.synth_image_batch(B, 3, 224, seed=1000 + rank, kind="ct"))
    pool = [(host + 0.001 * i).to(dev) for i in range(n_rot)]      # 8 x 57.8 MB fp32 (batch 96) = 462 MB > 126 MB L2
    # K forwards IN FLIGHT: K captured graphs of the same model (own static input / output buffers, own stream each), steps go
    # round-robin over them.  Every step is still one batch-B forward; at batch 24 each kernel is <= 2 waves and latency-bound,
    # so independent forwards overlap on the SMs (the engine's predict_stream does the same, see cswin_unet_b200/engine.py).
    K = max(1, args.inflight)
    from cswin_unet_b200 import _lib as cwlib
    cap_kb = int(os.environ.get("CSWIN_INFLIGHT_SMEM_CAP_KB", "100" if B <= 32 else "0")) if K > 1 else 0     # as SliceEngine(inflight > 1) sets it
    streams = [torch.cuda.Stream() for _ in range(K + 1)]
    static_x = [pool[0].clone() for _ in range(K + 1)]
    graphs, static_y = [], []
    with torch.no_grad():
        n0 = cw.launch_count()
        model(static_x[0])
        launches_per_fwd = cw.launch_count() - n0
        for k in range(K + 1):                                     # graph K: the single-forward (latency) configuration, no smem cap
            cwlib.set_option(cwlib.OPT_GEMM_SMEM_CAP_KB, cap_kb if k < K else 0)
            s = streams[k]
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    model(static_x[k])
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                static_y.append(model(static_x[k]))
            graphs.append(g)
    cwlib.set_option(cwlib.OPT_GEMM_SMEM_CAP_KB, 0)                # everything captured / launched below is single-stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step(i):
        k = i % K
        with torch.cuda.stream(streams[k]):
            static_x[k].copy_(pool[i % n_rot], non_blocking=True)
            graphs[k].replay()

    def timed_steps(n, ks):
        """n steps round-robin over the graphs `ks`; CUDA events on the current stream bracket all of them."""
        cur = torch.cuda.current_stream()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(cur)
        for kk in ks:
            streams[kk].wait_event(a0)
        for i in range(n):
            kk = ks[i % len(ks)]
            with torch.cuda.stream(streams[kk]):
                static_x[kk].copy_(pool[i % n_rot], non_blocking=True)
                graphs[kk].replay()
        for kk in ks:
            cur.wait_stream(streams[kk])
        a1.record(cur)
        return a0, a1

    # ---- value: device-resident inputs ----
    for i in range(max(args.warmup, K)):
        step(i)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    prof = os.environ.get("CSWIN_BENCH_PROFILER") == "1"          # `ncu --profile-from-start off`: record exactly the timed region
    if prof:
        torch.cuda.profiler.start()
    e0, e1 = timed_steps(args.steps, list(range(K)))
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.steps * B / (ms * 1e-3)
    # the same K steps with ONE forward in flight: latency of a single batch-B forward (what round 1 reported as `value`)
    barrier()
    timed_steps(3, [K])
    l0, l1 = timed_steps(args.steps, [K])
    barrier()
    ms_single = l0.elapsed_time(l1) / args.steps

    # ---- e2e: HOST buffers through the public API (SliceEngine): every step copies its batch host->device from pinned
    #      memory, runs the forward, and reads the uint8 label map back; copies of adjacent steps overlap the forward ----
    engine = cw.SliceEngine(model, batch=B, compute_dtype=torch.bfloat16, inflight=K)
    # the host batch is what the reference's evaluation loop hands to the model: single-channel slices (utils.py:69-71), which
    # CSwinUnet.forward repeats to 3 channels (vision_transformer.py:40-41) — here on the device, after ONE channel crossed PCIe.
    # (`host` has three identical channels: synth 'ct' slices are that repeat.)
    pinned = [(host[:, :1] + 0.001 * i).contiguous().pin_memory() for i in range(4)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d_bytes, d2h_bytes = engine.bytes_per_batch(host_chans=1)

    def e2e_run(n):
        done = 0
        for lab in engine.predict_stream(pinned[i % 4] for i in range(n)):
            done += lab.shape[0]
        return done

    e2e_run(max(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    n_done = e2e_run(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(e0.elapsed_time(e1), wall_ms)                    # results are consumed on the host: wall clock bounds it
    assert n_done == args.steps * B
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * args.steps * B / (ms_e2e * 1e-3)

    # ---- roofline: the kernel family is re-launched from a CUDA graph that contains ONLY those launches (same
    #      arguments and buffers as one real forward, L2-warm like in the step), timed with CUDA events ----
    pk = peaks()

    def family_ms(fn_names, reps=20):
        """(ms per forward, calls) of the launches that the ops `fn_names` make in one forward; a call is (fn, args, kwargs)."""
        fn_names = [fn_names] if isinstance(fn_names, str) else list(fn_names)
        calls = []
        origs = {n: getattr(ops, n) for n in fn_names}

        def mk(n):
            def rec(*a, **k):
                calls.append((origs[n], a, k))
                return origs[n](*a, **k)
            return rec
        for n in fn_names:
            setattr(ops, n, mk(n))
        try:
            with torch.no_grad():
                model(pool[0])
        finally:
            for n in fn_names:
                setattr(ops, n, origs[n])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s2 = torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2), torch.no_grad():
            for f, a, k in calls:
                f(*a, **k)
        torch.cuda.current_stream().wait_stream(s2)
        with torch.no_grad(), torch.cuda.graph(g):
            for f, a, k in calls:
                f(*a, **k)
        g.replay()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            g.replay()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / reps, calls

    att_total_ms, att_calls = family_ms("lepe_attention_fwd")
    att_bytes = attention_bytes_per_image() * B
    att_gbs = att_bytes / (att_total_ms * 1e-3) / 1e9
    roofline_attention = {"kernel": "lepe_attn_fwd_tc_kernel (fused LePE stripe attention; 26 launches per forward, both branches per launch)",
                "bound": "hbm", "achieved": att_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": att_gbs / pk["hbm_gbs"],
                "traffic": ATTN_DRAM_TRAFFIC.get(B), "peak_source": pk["source"], "bytes_per_forward": att_bytes,
                "ms_per_forward": att_total_ms, "launches_per_forward": len(att_calls),
                "how": "CUDA graph of the 26 attention launches of one forward (real buffers, L2-warm as in the step), "
                       "CUDA events over 20 replays"}
    lin_total_ms, lin_calls = family_ms(["linear", "conv_tokens"])      # conv_tokens = the same kernel fetching its A operand from the token image
    # the same family with K copies of that graph in flight (own streams), as in the timed region of `value`
    lin_inflight_ms = None
    if K > 1:
        try:
            gs = []
            cwlib.set_option(cwlib.OPT_GEMM_SMEM_CAP_KB, cap_kb)       # as the forwards of `value` were captured
            for k in range(K):
                with torch.no_grad(), torch.cuda.stream(streams[k]):
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2, stream=streams[k]):
                        for f, a, kw in lin_calls:
                            f(*a, **kw)
                gs.append(g2)
            torch.cuda.synchronize()
            cur = torch.cuda.current_stream()
            best = None
            for _ in range(2):
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record(cur)
                for k in range(K):
                    streams[k].wait_event(b0)
                for rep in range(10):
                    for k in range(K):
                        with torch.cuda.stream(streams[k]):
                            gs[k].replay()
                for k in range(K):
                    cur.wait_stream(streams[k])
                b1.record(cur)
                torch.cuda.synchronize()
                best = b0.elapsed_time(b1) / (10 * K)
            lin_inflight_ms = best
            del gs
        except Exception as e:                                     # noqa: BLE001 — an extra, never the headline
            print(f"[bench] in-flight family timing skipped: {e}", file=sys.stderr)
        finally:
            cwlib.set_option(cwlib.OPT_GEMM_SMEM_CAP_KB, 0)
    lin_flops = 0.0
    for f, a, k in lin_calls:
        if f is ops.conv_tokens:                                   # (x, H, W, w, bias, KH, KW, stride, pad): M = B OH OW, N x K = w.shape
            x_, H_, W_, w_, _, KH_, KW_, st_, pd_ = a
            m = x_.shape[0] * ((H_ + 2 * pd_ - KH_) // st_ + 1) * ((W_ + 2 * pd_ - KW_) // st_ + 1)
            lin_flops += 2.0 * m * w_.shape[0] * w_.shape[1]
            continue
        kk = a[1].shape[1]
        m = a[0].numel() // a[0].shape[-1]
        lin_flops += 2.0 * m * kk * (k.get("n_out") or a[1].shape[0])
    # the DOMINANT kernel of the step (about half of the forward's device time): the tcgen05 Linear family
    roofline = {"kernel": "linear_tc_kernel (tcgen05 Linear + fused LayerNorm / bias / GELU / residual epilogue; every nn.Linear, 1x1 conv and "
                          "(as an implicit GEMM) 3x3 conv of the forward)", "bound": "tensor",
                "achieved": lin_flops / (lin_total_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": lin_flops / (lin_total_ms * 1e-3) / 1e12 / pk["bf16_tflops"], "traffic": None, "peak_source": pk["source"] + " (burst)",
                "launches_per_forward": len(lin_calls), "ms_per_forward": lin_total_ms, "flops_per_forward": lin_flops,
                "share_of_step": lin_total_ms / ms_single,
                "inflight": None if lin_inflight_ms is None else {
                    "forwards_in_flight": K, "ms_per_forward": lin_inflight_ms,
                    "achieved": lin_flops / (lin_inflight_ms * 1e-3) / 1e12, "frac": lin_flops / (lin_inflight_ms * 1e-3) / 1e12 / pk["bf16_tflops"],
                    "what": "the same launches with K copies of the graph replayed concurrently on K streams (the regime `value` is timed in)"},
                "how": "CUDA graph of the Linear launches of one forward (real buffers, L2-warm as in the step), CUDA events over 20 replays; "
                       "algorithmic flops = 2 M N K of every launch",
                "why_low": "K = 64..256 for 98 of the 114 launches: arithmetic intensity 40-120 flop/B, i.e. these GEMMs sit at or below the "
                           "HBM / L2 ridge (252 flop/B), and their time is the epilogue (bias / LayerNorm fold / GELU / residual / row statistics: "
                           "~14 instructions per output element-pair, issue-bound at 0.4-0.5 IPC per scheduler, profiles/r02_ncu_linear_fc1_b96.txt) "
                           "plus one TMA round trip per tile; at batch 24 every launch is additionally <= 2 waves (latency chain)"}
    model_tflops = value / world * GFLOP_PER_SLICE_FWD / 1e3
    roofline_model = {"bound": "tensor", "achieved": model_tflops, "peak": pk["bf16_tflops"],
                      "unit": "TFLOP/s", "frac": model_tflops / pk["bf16_tflops"],
                      "note": "whole forward, 10.028 GFLOP/slice algorithmic, per GPU, vs the burst bf16 peak"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cswin_tiny_224_lite (9 classes) eval forward, batch {B}/GPU, 3x224x224 synthetic CT-like "
                                   "slices, synthetic weights; BASELINE configs[0] / [2] network, forward pass (slice batches as the volume loop of "
                                   "configs[3] would feed them; batch-24 figures in extra.forward_batch24)",
                       "global_batch": B * world, "parallelism": f"slice-sharded replicas x{world}, no collective",
                       "l2": f"inputs rotate over {n_rot} batches = {n_rot * host.numel() * 4 / 1e6:.0f} MB > 126 MB L2",
                       "launch": f"CUDA graph replay of the native forward, {K} forwards in flight (one graph + stream each, steps round-robin)",
                       "inflight": K},
            "single_forward": {"ms_per_step": ms_single, "value": world * B / (ms_single * 1e-3), "unit": UNIT,
                               "what": "same steps with ONE forward in flight: latency of a batch-" + str(B) + " forward"},
            "clocks": clocks, "gpu_launches": int(launches_per_fwd * args.steps),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / args.steps,
                    "path": "SliceEngine.predict_stream: pinned host fp32 batch of single-channel slices (as test_single_volume feeds them) -> H2D -> 1->3 channel repeat on the device -> graph-replayed forward with in-kernel "
                            f"argmax -> D2H uint8 label map; {K} forwards in flight (own slot, graph and stream each), copies overlap the forwards"},
            "roofline": roofline, "roofline_attention": roofline_attention, "roofline_model": roofline_model}

    # ---- train step (BASELINE configs[2]): forward + native backward + gradient all-reduce (NCCL, N > 1) + SGD, bf16
    #      compute with fp32 master weights, batch 24 per GPU, drop_path 0.2 active.  Reported next to the headline. ----
    if not args.no_train:
        # The train leg must never cost the headline: a failure is reported inside `train_step`, and a stall (e.g. a wedged
        # collective on one rank) is cut by a watchdog that still prints the line — every rank arms its own.
        fired = threading.Event()

        def train_watchdog():
            if not fired.wait(args.train_timeout):
                line["train_step"] = {"error": f"train leg exceeded {args.train_timeout:.0f} s and was abandoned"}
                if rank == 0:
                    emit(line)
                sys.stderr.flush()
                os._exit(0)
        threading.Thread(target=train_watchdog, daemon=True).start()
        try:
            import copy
            tmodel = copy.deepcopy(model).train()
            step_fn = cw.TrainStep(tmodel, lr=0.05, compute_dtype=torch.bfloat16)
            TB = min(BATCH, B)                                          # BASELINE configs[2]: batch 24 per GPU
            tpool = [p_[:TB].contiguous() for p_ in pool]
            timg = tpool[0]
            tlab = torch.from_numpy(synth.synth_labels(TB, 224, 9, seed=rank)).to(dev)
            tw, tk = 5, max(5, min(args.steps, 20))                     # warm-up: 3 eager steps + graph capture + 1 replay
            for _ in range(tw):
                step_fn(timg, tlab)
            barrier()
            n_tr0 = cw.launch_count()
            e0.record()
            for i in range(tk):
                step_fn(tpool[i % n_rot], tlab)
            e1.record()
            barrier()
            ms_tr = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms_tr], device=dev)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms_tr = float(t.item())
            line["train_step"] = {"value": world * tk * TB / (ms_tr * 1e-3), "unit": UNIT, "ms_per_step": ms_tr / tk, "steps": tk,
                                  "warmup": tw, "batch_per_gpu": TB, "dtype": "bf16 compute, fp32 master weights + gradients",
                                  "gpu_launches": int(step_fn.native_launches_per_step * tk + (cw.launch_count() - n_tr0)),
                                  "what": "forward + native backward + native loss (0.4 CE + 0.6 Dice) + NCCL gradient all-reduce overlapped with the "
                                          "backward on the pooled gradient buffer (N>1) + native fused SGD(momentum .9, wd 1e-4); whole step "
                                          "(collectives included) replayed as one CUDA graph",
                                  "roofline_frac_tensor": world and (tk * TB / (ms_tr * 1e-3)) * 33.231 / 1e3 / (pk["bf16_tflops_sustained"] or pk["bf16_tflops"])}
            step_fn.close()
            del tmodel, step_fn, tpool
        except Exception as e:                                  # noqa: BLE001 — reported, not swallowed
            import traceback
            traceback.print_exc()
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            fired.set()

    if not args.no_extras:
        done_x = threading.Event()

        def extras_watchdog():                                     # a stalled extra leg must not cost the headline line
            if not done_x.wait(args.train_timeout):
                line["extra"] = {"error": f"extra workloads exceeded {args.train_timeout:.0f} s and were abandoned"}
                if rank == 0:
                    emit(line)
                sys.stderr.flush()
                os._exit(0)
        threading.Thread(target=extras_watchdog, daemon=True).start()
        try:
            line["extra"] = extra_workloads(cw, synth, model, dev, rank, world, barrier, pool[0][:BATCH].contiguous(), args)
        except Exception as e:                                     # noqa: BLE001 — reported, not swallowed
            import traceback
            traceback.print_exc()
            line["extra"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            done_x.set()
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        try:                                                       # the eager-PyTorch reference on this same GPU
            rc = reference_cuda_rates(dev, batches=tuple(dict.fromkeys((BATCH, 1, B, 192))))
            line["reference_cuda"] = rc if rc is not None else {"unavailable": "baseline/_ref did not travel with the snapshot"}
            if rc is not None:
                line["speedup_vs_reference_cuda"] = {
                    "bf16_vs_ref_fp32": value / rc[f"batch{B}_fp32"]["slices_per_s"],
                    "bf16_vs_ref_bf16_autocast": value / rc[f"batch{B}_bf16_autocast"]["slices_per_s"],
                    "what": f"`value` over the unmodified reference's eager forward at the same batch ({B}) on the same GPU"}
                ts, rt = line.get("train_step", {}), rc.get(f"train_step_batch{BATCH}_bf16_autocast")
                if ts.get("value") and rt:
                    line["speedup_vs_reference_cuda"]["train_step_vs_ref_bf16_autocast"] = ts["value"] / rt["slices_per_s"]
                    line["speedup_vs_reference_cuda"]["train_step_vs_ref_fp32"] = ts["value"] / rc[f"train_step_batch{BATCH}_fp32"]["slices_per_s"]
        except Exception as e:                                     # noqa: BLE001
            line["reference_cuda"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, n, dt, kind = cpu_forward_rate(args.cpu_budget, B)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{n} x {B} slices (as forwards of batch {min(B, CPU_CHUNK)}) in {dt:.1f} s, "
                                          f"{'unmodified reference' if kind == 'reference' else 'oracle port'} (torch CPU fp32), {cores} threads"}
        try:                                                       # batch-1 forward and one train step on the host cores
            ex = cpu_extra_rates()
            if ex is not None:
                line["cpu_baseline"]["more"] = ex
        except Exception as e:                                     # noqa: BLE001
            line["cpu_baseline"]["more"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        emit(line)
    if world > 1:
        shutdown_dist()
    return 0


def extra_workloads(cw, synth, model, dev, rank, world, barrier, x24, args):
    """The other BASELINE configs as extra keys of the same line (all ranks take part; values are whole-job, max-over-ranks time):
    fp32 forward (the <= 1e-4 parity path), configs[3] = 150 x 512^2 volume through the test_single_volume loop slice-sharded over
    the ranks, configs[4] = 512^2 / 3 classes / split [1,2,8,8] forward and train step."""
    import torch.distributed as dist

    def max_ms(ms_local):
        if world == 1:
            return ms_local
        t = torch.tensor([ms_local], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def graph_rate(m, x, reps):
        with torch.no_grad():
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    m(x)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m(x)
        for _ in range(3):
            g.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        barrier()
        ms = max_ms(e0.elapsed_time(e1)) / reps
        del g
        return ms

    out = {}
    # ---- the same bf16 forward at batch 24 (BASELINE configs[2]'s batch; the headline of the earlier rounds): one forward at a
    #      time = the latency of a batch-24 forward, and 3 in flight ----
    model.compute_dtype = torch.bfloat16
    ms = graph_rate(model, x24, 20)
    fb24 = {"value": world * x24.shape[0] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch_per_gpu": int(x24.shape[0]),
            "what": "same network and kernels, ONE batch-24 forward at a time: every launch is <= 2 waves, the step is a latency chain of ~160 launches"}
    try:
        K3 = 3
        ss = [torch.cuda.Stream() for _ in range(K3)]
        xs = [x24 + 0.001 * i for i in range(K3)]
        gs = []
        with torch.no_grad():
            for s_, x_ in zip(ss, xs):
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    model(x_)
                s_.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=s_):
                    model(x_)
                gs.append(g_)
        barrier()
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        e0.record(cur)
        for s_ in ss:
            s_.wait_event(e0)
        for i in range(reps):
            with torch.cuda.stream(ss[i % K3]):
                gs[i % K3].replay()
        for s_ in ss:
            cur.wait_stream(s_)
        e1.record(cur)
        barrier()
        ms3 = max_ms(e0.elapsed_time(e1)) / reps
        fb24["inflight3"] = {"value": world * x24.shape[0] / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3,
                             "what": "3 batch-24 forwards in flight (own graph + stream each): the earlier headline configuration"}
        del gs
    except Exception as e:                                         # noqa: BLE001 — an extra
        fb24["inflight3"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    out["forward_batch24"] = fb24
    # ---- fp32 forward: the exact SIMT path every <= 1e-4 parity claim is made on ----
    model.compute_dtype = torch.float32
    ms = graph_rate(model, x24, 5)
    model.compute_dtype = torch.bfloat16
    out["fp32_forward"] = {"value": world * x24.shape[0] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch_per_gpu": int(x24.shape[0]),
                           "what": "same workload on the fp32 SIMT kernels (parity path, logits within 1e-4 of the fp32 reference)"}
    # ---- configs[3]: Synapse-shaped volume, 150 slices of 512^2 -> 224^2 -> labels -> 512^2, slice-sharded, no collective ----
    D, S = 150, 512
    v10, _ = synth.synth_seg_volume(10, S, 9, seed=5)
    vol = np.ascontiguousarray(np.tile(v10, (D // 10, 1, 1)))
    eng = cw.SliceEngine(model, batch=BATCH, compute_dtype=torch.bfloat16)
    res = {}
    for mode in ("gpu", "scipy"):
        if mode == "scipy" and world > 1:
            continue                                              # host-side scipy zoom: single-rank reference point only
        cw.predict_volume(eng, vol[:BATCH], resample=mode)        # warm-up (buffers, kernels)
        barrier()
        t0 = time.perf_counter()
        lab, rng = cw.predict_volume(eng, vol, shard=(rank, world), resample=mode)
        torch.cuda.synchronize()
        ms = max_ms((time.perf_counter() - t0) * 1e3)
        res[mode] = ms
    out["volume_150x512"] = {"value": D / (res["gpu"] * 1e-3), "unit": UNIT, "ms_per_volume": res["gpu"], "slices": D, "ranks": world,
                             "host_scipy_resampling_ms": res.get("scipy"),
                             "what": "BASELINE configs[3]: test_single_volume loop (utils.py:61-90) end to end from a host float32 volume to host uint8 "
                                     "labels: zoom order 3 to 224^2, bf16 forward with in-kernel argmax, zoom order 0 back to 512^2, both zooms on the "
                                     "GPU (bit-compatible with scipy); contiguous slice shards per rank, no collective, wall clock, max over ranks"}
    del eng
    # ---- configs[4]: 512^2 input, 3 classes, split [1,2,8,8] (windows of 128 / 256 tokens), batch 4 per GPU ----
    B5 = 4
    m5 = cw.cswin_tiny_224(num_classes=3, img_size=512, split_size=[1, 2, 8, 8]).eval()
    shapes = {k: tuple(v.shape) for k, v in m5.state_dict().items()}
    m5.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed=1234).items()}, strict=True)
    m5 = m5.to(dev)
    m5.compute_dtype = torch.bfloat16
    x5 = torch.from_numpy(synth.synth_image_batch(B5, 3, 512, seed=rank, kind="ct")).to(dev)
    ms5 = graph_rate(m5, x5, 10)
    c5 = {"forward": {"value": world * B5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5},
          "batch_per_gpu": B5, "gflop_per_slice_fwd": 56.60,
          "what": "BASELINE configs[4]: cswin_tiny at 512^2, 3 classes, split [1,2,8,8], bf16"}
    if not args.no_train:
        step5 = cw.TrainStep(m5.train(), lr=0.05, compute_dtype=torch.bfloat16)
        y5 = torch.from_numpy(synth.synth_labels(B5, 512, 3, seed=rank)).to(dev)
        for _ in range(5):
            step5(x5, y5)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step5(x5, y5)
        e1.record()
        barrier()
        ms = max_ms(e0.elapsed_time(e1)) / 5
        c5["train_step"] = {"value": world * B5 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
        step5.close()
        del step5
    out["config_512"] = c5
    del m5
    torch.cuda.empty_cache()
    return out


class StdoutGuard:
    """Everything that libraries write to fd 1 while the benchmark runs (NCCL's version banner, ...) goes to stderr; only the
    result line reaches the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str) -> None:
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


GUARD = None


def emit(line: dict) -> None:
    text = json.dumps(line)
    if GUARD is not None:
        GUARD.emit(text)
    else:
        print(text, flush=True)


def main():
    global GUARD
    GUARD = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-cuda"])
    ap.add_argument("--batch", type=int, default=FWD_BATCH, help="slices per forward step (headline: 96; 24 = the latency point)")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("CSWIN_INFLIGHT", "2")),
                    help="batch-B forwards in flight (one CUDA graph + stream each); 1 = strictly one after the other")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-oracle work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip timing the eager reference on the GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra workloads (fp32 forward, volume, 512^2 config)")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step leg")
    ap.add_argument("--train-timeout", type=float, default=150.0, help="seconds before a stalled train-step leg is abandoned")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
