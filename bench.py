#!/usr/bin/env python
"""bench.py -- frames/sec of the detection post-processing hot path (fused decode + box_nms).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config coco608_b64] [--regime R|T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference path (oracle/)

A "step" is one pass of the hot path (YOLOv3 head maps -> (ids, scores, bboxes), yolo3.py:496,523-534)
over one batch of synthetic head maps.  Frames are independent, so N GPUs each take their own batch
(weak scaling, no collective on the data path); the only collective is the MAX of the per-rank
elapsed times.  One JSON line is printed by rank 0.

  value         frames/sec with the head maps already resident in HBM (CUDA events, max over ranks)
  e2e           frames/sec through videoyolo_b200.HostDetector: pinned host head maps -> device ->
                kernels -> host results, copies inside the timed region
  roofline      the dominant kernel (candidate selection, the only one that streams the head maps):
                algorithmic bytes per launch / its average launch duration (CUDA events on the launch
                stream, recorded by the library: vy_prof_*) against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle/ (C restatement of decode + box_nms) on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (BASELINE.json config, classes, input size, frames per GPU)
CONFIGS = {
    "voc416_b1": ("configs[0]: YOLOv3 VOC (20 cls) 416x416 batch 1", 20, 416, 1),
    "coco608_b64": ("configs[1]: YOLOv3 COCO (80 cls) 608x608 batch 64 decode + box_nms (22743 boxes/frame)", 80, 608, 64),
    "vid416_b32": ("configs[2]: ImageNet-VID (30 cls) 416x416 decode + NMS, batch 32", 30, 416, 32),
    "stress416_b128": ("configs[3] shape: 80 cls 416x416 (10647 boxes) batch 128, reference NMS arguments", 80, 416, 128),
    "vid320_b256": ("configs[4]: frame-sharded stream 320x320 batch 256/GPU, VID 30 cls", 30, 320, 256),
}
NMS = dict(nms_thresh=0.45, valid_thresh=0.01, topk=400, post_nms=100)      # detect_yolo3.py:200, yolo3.py:527
FALLBACK_HBM_GBS = 6650.0                                                     # B200_PROFILING.md fallback


def grid_sizes(size):
    return [size // 32, size // 16, size // 8]


def frame_bytes(C, size, post_nms=100):
    """Algorithmic bytes per frame: compulsory fp32 read of the three head maps + the
    write of the (post_nms, 6) triple and its kept-row indices."""
    n_box = sum(g * g * 3 for g in grid_sizes(size))
    return 4 * n_box * (5 + C), 24 * post_nms + 4 * post_nms


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_bf16_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json bf16_tflops)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md)"


def fusion_conv_leg(dev, B=8, T=3, iters=10):
    """Temporal fusion conv (layers.py:73-79 as used at yolo3.py:250-251): the three K=3 tip convs of
    BASELINE configs[2] (ImageNet-VID 416^2), 3x3x3, Cin -> 2*Cin, bf16 operands / fp32 accumulation in
    TMEM.  Kernel timed alone (CUDA events, L2 flushed between launches) against the measured cuBLAS bf16
    burst peak.  'formula' FLOPs = SURVEY 8d: 2*B*T*H*W*Cout*Cin*27 (counts the zero-padded temporal taps
    the kernel skips and not the border pixels it computes); 'executed' = what the tensor pipe really did."""
    import torch
    from videoyolo_b200 import ops
    peak, src = load_bf16_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for g, cin in ((13, 512), (26, 256), (52, 128)):
        cout = 2 * cin
        x = ops.PTensor(torch.zeros((T, B, g + 2, g + 2, cin), dtype=torch.bfloat16, device=dev), B, T, g, g, cin)
        x.data[:, :, 1:-1, 1:-1] = torch.randn((T, B, g, g, cin), device=dev).to(torch.bfloat16)
        w = (torch.rand((cout, 3, 3, 3, cin), device=dev) * 0.14 - 0.07).to(torch.bfloat16)      # MXNet Uniform(0.07)
        sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        for _ in range(3):
            ops.fusion_conv(x, w, sc, sh)
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.fusion_conv(x, w, sc, sh); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        f_formula = 2.0 * B * T * g * g * cout * cin * 27
        rows = -(-(B * (g + 2) * (g + 2)) // 128) * 128
        taps = sum(sum(1 for dt in (-1, 0, 1) if 0 <= t + dt < T) for t in range(T)) * 9
        f_exec = 2.0 * rows * cout * cin * taps
        out.append({"shape": "B%d T%d %dx%d %d->%d k3x3x3" % (B, T, g, g, cin, cout), "ms": round(ms, 4),
                    "tflops_formula": round(f_formula / ms / 1e9, 1), "tflops_executed": round(f_exec / ms / 1e9, 1),
                    "frac_formula": round(f_formula / ms / 1e9 / peak, 4), "frac_executed": round(f_exec / ms / 1e9 / peak, 4)})
    tot_ms = sum(o["ms"] for o in out)
    return {"workload": "configs[2] fusion conv: K=3 tip convs of YOLODetectionBlockV3 at 416^2, batch %d windows" % B,
            "bound": "tensor", "unit": "TFLOP/s", "peak": peak, "peak_source": src, "dtype": "bf16 x bf16 -> f32 (TMEM)",
            "windows_per_s": round(B / (tot_ms * 1e-3), 1), "shapes": out,
            "l2": "L2 flushed between launches (256 MiB write)"}


def graphed_ms(net, xs, flush, iters):
    """The same forward as ONE CUDA-graph replay per call (pipeline.GraphedModule; the inputs sit in the graph's static
    buffers, where a backbone in front of it would write them; L2 flushed between calls): the host is out of the
    launch chain."""
    import torch
    from videoyolo_b200.pipeline import GraphedModule
    g = GraphedModule(net, xs)
    for _ in range(3):
        g()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def temporal_tail_leg(dev, B=32, K=3, C=30, size=416, iters=10):
    """BASELINE configs[2] end to end on the device: ImageNet-VID (30 cls) 416^2, temporal window K=3:
    (B, K, channel, g, g) fp32 block outputs -> P-layout pack -> 3x3x3 tip conv (tcgen05) -> late 'max' join ->
    1x1 prediction conv (same tcgen05 kernel, identity activation, fp32 out) -> fused decode + box_nms -> (ids, scores, bboxes).
    One window = one detection; CUDA events around the whole call, L2 flushed between calls."""
    import torch
    import videoyolo_b200 as vy
    from videoyolo_b200 import _lib
    torch.manual_seed(7)
    net = vy.YOLOV3T(["c%d" % i for i in range(C)], k=K, k_join_type="max", block_conv_type="3").to(dev).eval()
    xs = [torch.randn((B, K, c, g, g), device=dev) for c, g in zip((512, 256, 128), grid_sizes(size))]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net(*xs)
        _lib.prof_enable(True); _lib.prof_read()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net(*xs); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        prof = _lib.prof_read(); _lib.prof_enable(False)
    ms = sorted(ts)[len(ts) // 2]
    gms = graphed_ms(net, xs, flush, iters)
    return {"workload": "configs[2]: ImageNet-VID (30 cls) 416x416, K=3: tip fusion conv + max join + prediction conv + decode + NMS, batch %d windows" % B,
            "windows_per_s": round(B / (ms * 1e-3), 1), "ms_per_call": round(ms, 4),
            "cuda_graph": {"windows_per_s": round(B / (gms * 1e-3), 1), "ms_per_call": round(gms, 4)},
            "library_kernel_ms_per_call": {k: round(v[0] / iters, 4) for k, v in prof.items()},
            "note": "device-resident fp32 inputs; every kernel on the path is the library's own (no cuDNN/cuBLAS); L2 flushed between calls"}


def temporal_neck_leg(dev, B=8, K=3, C=30, size=416, iters=10, calibrated_bn=True):
    """The whole post-backbone part of the temporal detector on the device: Darknet-53 stage
    outputs (B, K, 1024/512/256, g, g) fp32 -> detection blocks (five cells + tip, 3-D convs), transitions, upsample +
    concat, late 'max' join, prediction conv, fused decode + box_nms.  FLOPs by the SURVEY formula over every conv cell."""
    import torch
    import videoyolo_b200 as vy
    from videoyolo_b200 import _lib
    torch.manual_seed(9)
    net = vy.YOLOV3TNeck(["c%d" % i for i in range(C)], k=K, k_join_type="max", block_conv_type="3").to(dev).eval()
    # BatchNorm statistics that match the synthetic data, as a trained network's do: with the constructor's var = 1 the
    # seven Uniform(0.07) layers of a block amplify N(0,1) inputs ~200x, every sigmoid of the decode saturates and a
    # quarter of all scores tie at exactly 1.0 (selection worst case, 0.10 ms instead of 0.03 ms of decode + NMS per call).
    # Per cell: var = fan_in * Var(w) * E[x^2], E[x^2] = 1 for the N(0,1) stage outputs, 0.505 after a LeakyReLU(0.1).
    if calibrated_bn:
        from videoyolo_b200.layers import _Cell
        first = set()
        for blk in net.blocks:
            first.add(id(next(m for m in blk.modules() if isinstance(m, _Cell))))
        for m in net.modules():
            if isinstance(m, _Cell):
                fan_in = m.weight[0].numel()
                m.running_var.fill_(fan_in * (0.07 ** 2 / 3.0) * (1.0 if id(m) in first else 0.505))
    rs = [torch.randn((B, K, c, g, g), device=dev) for c, g in zip((1024, 512, 256), grid_sizes(size))]
    flops = 0.0
    for i, (blk, g) in enumerate(zip(net.blocks, grid_sizes(size))):
        convs = list(blk.body) + [net.head.tips[i]] + ([net.transitions[i].model] if i < len(net.transitions) else [])
        for conv in convs:
            for cell in conv.cells:
                co, ci = cell.weight.shape[0], cell.weight.shape[1]
                flops += 2.0 * B * K * g * g * co * ci * cell.k3[0] * cell.k3[1] * cell.k3[2]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net(*rs)
        _lib.prof_enable(True); _lib.prof_read()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net(*rs); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        prof = _lib.prof_read(); _lib.prof_enable(False)
    ms = sorted(ts)[len(ts) // 2]
    gms = graphed_ms(net, rs, flush, iters)
    return {"workload": "YOLOV3T after the backbone (yolo3.py:1126-1206), VID 30 cls 416x416, K=3, 3-D convs, batch %d windows" % B,
            "windows_per_s": round(B / (ms * 1e-3), 1), "ms_per_call": round(ms, 4),
            "gflop_per_window": round(flops / B / 1e9, 1), "tflops_formula": round(flops / (ms * 1e-3) / 1e12, 1),
            "cuda_graph": {"windows_per_s": round(B / (gms * 1e-3), 1), "ms_per_call": round(gms, 4),
                           "tflops_formula": round(flops / (gms * 1e-3) / 1e12, 1)},
            "batchnorm": ("running_var = fan_in * Var(w) * E[x^2] per cell (unit-scale activations, as trained statistics give)"
                          if calibrated_bn else "constructor defaults (var 1): saturated logits, a quarter of the scores tie at 1.0"),
            "library_kernel_ms_per_call": {k: round(v[0] / iters, 4) for k, v in prof.items()}}


def voc_latency_leg(dev):
    """BASELINE configs[0]: one VOC frame (20 cls, 416^2, batch 1) -- the call's latency, not a roofline: eager library
    call and CUDA-graph replay, back to back and isolated (launch -> synchronize wall time, median)."""
    import torch
    import videoyolo_b200 as vy
    from videoyolo_b200.pipeline import GraphedDetector
    from videoyolo_b200.synth import random_heads_cuda
    AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
    label, C, size, B = CONFIGS["voc416_b1"]
    heads = random_heads_cuda(B, C, size, 3, dev)
    det = GraphedDetector(C, AN, ST, [tuple(h.shape) for h in heads], dev)
    det(heads)

    def wall(fn, n=300):
        for _ in range(30):
            fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6

    def one(fn, n=200):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e6)
        return sorted(ts)[n // 2]

    e = lambda: vy.yolo3_decode_nms(heads, C, AN, ST, out=det.out, kept=det.kept, workspace_buf=det.workspace, **NMS)
    g = lambda: det()
    r = {"workload": label, "unit": "us per frame",
         "back_to_back_eager_us": round(wall(e), 2), "back_to_back_graph_us": round(wall(g), 2),
         "isolated_eager_us": round(one(e), 2), "isolated_graph_us": round(one(g), 2)}
    r["frames_per_s_graph"] = round(1e6 / r["back_to_back_graph_us"], 1)
    return r


def stress_nms_leg(dev, peak, B=128):
    """BASELINE configs[3] with ITS arguments: 80 cls, 10 647 boxes (R = 851 760 rows per frame), batch 128,
    valid_thresh = 0.001, topk = -1 (every valid row takes part), force_suppress off and on; the operator call of
    yolo3.py:526-528 on the materialised (B, R, 6) tensor, full-size output (post_nms = -1).  Bytes per frame
    (SURVEY.md 8d): the score/box read 24*R + the mandatory full-size output 24*R."""
    import torch
    import videoyolo_b200 as vy
    from videoyolo_b200.synth import random_heads_cuda
    AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
    heads = random_heads_cuda(B, 80, 416, 5, dev)
    dets = vy.yolo3_decode(heads, 80, AN, ST)
    del heads
    R = dets.shape[1]
    out = {"workload": CONFIGS["stress416_b128"][0].replace("reference NMS arguments", "valid_thresh 0.001, topk -1, full output"),
           "rows_per_frame": R, "batch": B, "bytes_per_frame": 48 * R, "regime": "R: logits ~ N(0,1)"}
    for force in (False, True):
        fn = lambda: vy.box_nms(dets, 0.45, 0.001, -1, id_index=0, force_suppress=force)
        o = fn(); torch.cuda.synchronize()
        surv = float((o[..., 0] >= 0).sum()) / B
        del o
        ts = []
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); o = fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
            del o
        ms = min(ts)
        out["force_suppress_%s" % ("on" if force else "off")] = {
            "ms_per_call": round(ms, 2), "frames_per_s": round(B / ms * 1e3, 1), "survivors_per_frame": round(surv, 1),
            "hbm_frac": round(48.0 * R * B / (ms * 1e-3) / 1e9 / peak, 5)}
    return out


def load_traffic(config_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(config_name, {}).get(kernel)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while `active` is set."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = threading.Event()
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                time.sleep(0.002)
            else:
                time.sleep(0.0005)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------- CPU legs
def cpu_port_fps(C, size, frames, seed, threads, steps=1, warmup=0):
    """Oracle (C restatement of decode yolo3.py:151-199 + box_nms yolo3.py:523-534) on `frames` frames
    per step with `threads` host threads.  Returns (frames/sec, seconds per step)."""
    import numpy as np
    import oracle
    oracle.set_threads(threads)
    rng = np.random.RandomState(seed)
    heads = [rng.standard_normal(size=(frames, 3 * (5 + C), g, g)).astype(np.float32) for g in grid_sizes(size)]
    for _ in range(warmup):
        oracle.yolov3_postprocess(heads, C, NMS["nms_thresh"], NMS["topk"], NMS["post_nms"], NMS["valid_thresh"])
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.yolov3_postprocess(heads, C, NMS["nms_thresh"], NMS["topk"], NMS["post_nms"], NMS["valid_thresh"])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return frames / dt, dt


def make_config(args, world, n_streams=None):
    """The `config` object of the JSON line: the SAME dict for the native and the reference arm (the driver compares them)."""
    label, C, size, B = CONFIGS[args.config]
    in_bytes_frame, _ = frame_bytes(C, size, NMS["post_nms"])
    in_bytes = in_bytes_frame * B
    return {"workload": label, "classes": C, "input": size, "frames_per_gpu": B,
            "global_frames_per_step": B * world, "boxes_per_frame": in_bytes_frame // (4 * (5 + C)),
            "regime": args.regime + (": logits ~ N(0,1) (random-init weights)" if args.regime == "R" else ": trained-like"),
            "nms": NMS, "parallelism": "frames sharded over %d GPU(s), no data-path collective" % world,
            "l2": ("inputs larger than L2 (%.0f MB per step)" % (in_bytes / 1e6)) if in_bytes >= (160 << 20)
                  else "L2 flushed between steps (256 MiB write)",
            "streams": args.streams if in_bytes >= (160 << 20) else 1}


def run_reference(args):
    """--impl reference: the reference's CPU path.  MXNet/GluonCV cannot be installed here (no wheel,
    no network), so this is the oracle port with all host threads; a step is the whole batch of the
    config whenever K + W steps of that fit ~2.5 minutes (else a bounded sample of it, stated in cpu_baseline.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    label, C, size, B = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    frames = B
    fps1, dt1 = cpu_port_fps(C, size, min(B, max(cores, 8)), 1236, cores)
    budget = 150.0
    while frames > 1 and (frames / fps1) * (args.steps + args.warmup) > budget:
        frames = max(1, frames // 2)
    fps, dt = cpu_port_fps(C, size, frames, 1236, cores, steps=args.steps, warmup=args.warmup)
    sample = "%d of %d frames per step, %d threads over frames (oracle/vy_oracle.c)" % (frames, B, cores)
    line = {"impl": "reference", "metric": "frames/sec decode+NMS", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, max(1, args.gpus)),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs nearest to its GPU (NVML's ideal affinity) BEFORE any pinned host buffer is
    allocated: first-touch then places the buffers on that NUMA node, and eight ranks stop sharing node 0's memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="coco608_b64", choices=sorted(CONFIGS))
    ap.add_argument("--regime", default="R", choices=["R", "T"], help="R: N(0,1) logits (random-init); T: trained-like")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-conv", action="store_true", help="skip the fusion-conv (tensor-pipe) leg")
    ap.add_argument("--no-other", action="store_true", help="skip the short legs over the other BASELINE workloads")
    ap.add_argument("--no-graph", action="store_true", help="headline steps as eager library calls instead of CUDA-graph replays")
    ap.add_argument("--streams", type=int, default=4,
                    help="CUDA streams the timed steps alternate over (1 = every step waits for the one before)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpus = bind_to_gpu_numa(local)

    import torch
    import torch.distributed as dist
    import videoyolo_b200 as vy
    from videoyolo_b200 import _lib
    from videoyolo_b200.pipeline import GraphedDetector, HostDetector
    from videoyolo_b200.synth import random_heads_cuda

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up: send stdout to stderr until
        # then, so that the JSON line is the only thing rank 0 writes to stdout
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.lib()
    AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
    peak, peak_src = load_peaks()
    side = [torch.cuda.Stream(dev) for _ in range(max(1, args.streams))]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Workload:
        """One benchmarked config on this rank: ONE INPUT SET PER STREAM (own head maps, outputs and workspace), each
        captured once in a CUDA graph (pipeline.GraphedDetector: the same library calls, replayed).  Steps are
        independent batches: step i runs on stream i % n, so the latency-bound head (sample) and tail (finalize) of
        one step hide behind the bandwidth-bound streaming pass of its neighbours.  Inputs smaller than L2 run on one
        stream with a 256 MiB flush write between steps."""

        def __init__(self, name, seed, regime, graph=True):
            self.label, self.C, self.size, self.B = CONFIGS[name]
            self.in_frame, self.out_frame = frame_bytes(self.C, self.size, NMS["post_nms"])
            self.in_bytes = self.in_frame * self.B
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if self.in_bytes < (160 << 20) else None
            self.n = max(1, args.streams) if self.flush is None else 1
            self.graph = graph
            self.sets = []
            for i in range(self.n):
                heads = random_heads_cuda(self.B, self.C, self.size, seed + 1000 * i + rank, dev, regime=regime)
                det = GraphedDetector(self.C, AN, ST, [tuple(h.shape) for h in heads], dev, **{
                    "nms_thresh": NMS["nms_thresh"], "valid_thresh": NMS["valid_thresh"], "nms_topk": NMS["topk"],
                    "post_nms": NMS["post_nms"]})
                for dst, src in zip(det.heads, heads):
                    dst.copy_(src)
                del heads
                self.sets.append(det)
            torch.cuda.synchronize()

        def step(self, i=0, n=1):
            det = self.sets[i % len(self.sets)]
            if n == 1:
                det.graph.replay()
            else:
                with torch.cuda.stream(side[i % n]):
                    det.graph.replay()

        def step_eager(self, i=0, n=1):
            det = self.sets[i % len(self.sets)]
            if n == 1:
                vy.yolo3_decode_nms(det.heads, self.C, AN, ST, out=det.out, kept=det.kept, workspace_buf=det.workspace, **NMS)
            else:
                with torch.cuda.stream(side[i % n]):
                    vy.yolo3_decode_nms(det.heads, self.C, AN, ST, out=det.out, kept=det.kept, workspace_buf=det.workspace, **NMS)

        def timed(self, fn, steps, warmup, sampler=None, n=1):
            """max-over-ranks milliseconds for `steps` calls of fn, device-timed."""
            for i in range(warmup):
                fn(i, n)
            barrier()
            if sampler:
                sampler.active.set()
            if self.flush is None:
                cur = torch.cuda.current_stream(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(cur)
                if n > 1:
                    for s_ in side[:n]:
                        s_.wait_event(a)
                for i in range(steps):
                    fn(i, n)
                if n > 1:
                    for s_ in side[:n]:
                        cur.wait_event(s_.record_event())
                b.record(cur)
                torch.cuda.synchronize()
                total = a.elapsed_time(b)
            else:
                evs = []
                for i in range(steps):
                    self.flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(i, 1); b.record()
                    evs.append((a, b))
                torch.cuda.synchronize()
                total = sum(a.elapsed_time(b) for a, b in evs)
            if sampler:
                sampler.active.clear()
            barrier()
            return max_over_ranks(total)

        def frac(self, ms_per_step):
            return (self.in_bytes + self.out_frame * self.B) / (ms_per_step * 1e-3) / 1e9 / peak

    label, C, size, B = CONFIGS[args.config]
    wl = Workload(args.config, 1236, args.regime)
    n_streams = wl.n
    step = wl.step_eager if args.no_graph else wl.step
    in_bytes = wl.in_bytes

    sampler = ClockSampler(local)
    sampler.start()

    # a fresh process on an idle GPU: a small config (one frame: inputs of a megabyte, the whole timed region a few
    # milliseconds) is over before the clocks have come up, so keep the device busy for a moment first.  The large
    # configs have just generated gigabytes of inputs on the device and go straight to their W warm-up steps.
    if wl.flush is not None:
        t_spin = time.perf_counter()
        while time.perf_counter() - t_spin < 0.5:
            for i in range(8):
                step(i, n_streams)
            torch.cuda.synchronize()

    # ---- headline: device-resident
    ms_total = wl.timed(step, args.steps, args.warmup, sampler, n_streams)
    # the same steps strictly one after the other (reported beside the headline), and both as eager library calls
    ms_serial = wl.timed(step, args.steps, 3, None, 1) if n_streams > 1 else ms_total
    c0 = _lib.launch_counts()
    ms_eager = wl.timed(wl.step_eager, args.steps, 3, None, n_streams)
    c1 = _lib.launch_counts()
    ms_eager_serial = wl.timed(wl.step_eager, args.steps, 3, None, 1) if n_streams > 1 else ms_eager
    # kernels per step: counted by the library on the eager launches (a graph replay launches the same kernels)
    per_step = {k: (c1[k] - c0[k]) // (args.steps + 3) for k in c1 if c1[k] - c0[k]}
    gpu_launches = sum(per_step.values()) * args.steps
    ms_step = ms_total / args.steps
    fps = world * B * args.steps / (ms_total * 1e-3)

    # ---- roofline pass: same steps (eager, one stream) with the library's per-kernel events switched on
    _lib.prof_enable(True)
    _lib.prof_read()
    ms_prof_total = wl.timed(wl.step_eager, args.steps, 3, None, 1)
    prof = _lib.prof_read()
    _lib.prof_enable(False)
    # the dominant kernel = the one with the largest share of the step (the warm-up launches of the
    # profiled pass are in the record too: average over all of them)
    kernel_ms = {k: v[0] / max(v[1], 1) * (v[1] / max(1, min(x[1] for x in prof.values()))) for k, v in prof.items()}
    top = max(kernel_ms, key=kernel_ms.get) if kernel_ms else "none"
    top_ms = prof[top][0] / max(prof[top][1], 1) if kernel_ms else 0.0
    achieved = in_bytes / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": load_traffic(args.config, top),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": in_bytes,
                "kernel_ms_per_step": {k: round(v, 5) for k, v in kernel_ms.items()},
                "kernel_share_of_step": {k: round(v / sum(kernel_ms.values()), 4) for k, v in kernel_ms.items()},
                "step_frac": round(wl.frac(ms_step), 4),
                "step_frac_single_stream": round(wl.frac(ms_serial / args.steps), 4),
                "step_frac_eager": round(wl.frac(ms_eager / args.steps), 4),
                "step_frac_single_stream_eager": round(wl.frac(ms_eager_serial / args.steps), 4),
                "profiled_ms_per_step": round(ms_prof_total / args.steps, 5)}

    # ---- e2e: host buffers through the public host-facing call
    e2e = None
    if not args.no_e2e:
        d0 = wl.sets[0]
        h_heads = [torch.empty(h.shape, dtype=torch.float32).pin_memory() for h in d0.heads]
        for hh, h in zip(h_heads, d0.heads):
            hh.copy_(h)
        torch.cuda.synchronize()
        det = HostDetector(C, AN, ST, dev, chunk=max(1, min(8, B)), **{"nms_thresh": NMS["nms_thresh"],
                           "valid_thresh": NMS["valid_thresh"], "nms_topk": NMS["topk"], "post_nms": NMS["post_nms"]})
        e2e_steps = args.steps if in_bytes * args.steps < (60 << 30) else max(3, (60 << 30) // in_bytes)
        last = {}

        def e2e_step():
            last["r"] = det(h_heads, copy=False)

        for _ in range(3):
            e2e_step()
        barrier()
        sampler.active.set()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        sampler.active.clear()
        e2e_ms = max_over_ranks(max(a.elapsed_time(b), wall_ms))
        barrier()
        # the host results of the last e2e step must be the device-resident results of the same input set
        d0.graph.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(last["r"][0][..., 0], d0.out.cpu()[..., 0]))
        e2e = {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": det.h2d_bytes, "d2h_bytes_per_step": det.d2h_bytes, "steps": e2e_steps,
               "ms_per_step": e2e_ms / e2e_steps, "api": "videoyolo_b200.pipeline.HostDetector.__call__",
               "h2d_gbs": round(det.h2d_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9, 1),
               "bound": "PCIe host->device copy of the fp32 head maps (the kernels take ~1% of the step)",
               "host_placement": ("process bound to the GPU's NUMA-local CPUs (%d cpus) before the pinned buffers were allocated" % len(cpus))
                                 if cpus else "default placement (NVML affinity unavailable)",
               "matches_device_path": same}
        del h_heads, det
    sampler.stop()

    # ---- temporal fusion conv: the tensor-core part of the path (rank 0 only; not part of `value`)
    conv = None
    if rank == 0 and not args.no_conv:
        try:
            conv = fusion_conv_leg(dev)
            conv["temporal_tail"] = temporal_tail_leg(dev)
            conv["temporal_neck"] = temporal_neck_leg(dev)
            raw = temporal_neck_leg(dev, calibrated_bn=False)
            conv["temporal_neck"]["constructor_batchnorm"] = {k: raw[k] for k in ("ms_per_call", "tflops_formula", "cuda_graph", "batchnorm")}
        except Exception as e:                     # the headline line must still be printed
            conv = {"error": str(e)[:200]}

    # ---- the other device-resident workloads of BASELINE.json, same method, short, on EVERY rank (so the scaling run
    # carries them too): the metric's own 416^2 / 10 647-box shape at batch 128 (configs[3]'s shape with the reference's
    # NMS arguments), configs[4] (the config BASELINE names for scaling) and configs[2]'s decode + NMS part
    other = None
    if args.config == "coco608_b64" and wl.flush is None and not args.no_other:
        other = {}
        del wl.sets[1:]
        for name in ("stress416_b128", "vid320_b256", "vid416_b32"):
            w2 = Workload(name, 4321, args.regime)
            k2 = min(args.steps, 50)
            ms_a = w2.timed(w2.step, k2, 5, None, w2.n)
            ms_b = w2.timed(w2.step, k2, 3, None, 1) if w2.n > 1 else ms_a
            other[name] = {"workload": w2.label, "steps": k2, "streams": w2.n, "n_gpus": world,
                           "value": round(world * w2.B * k2 / (ms_a * 1e-3), 1),
                           "value_single_stream": round(world * w2.B * k2 / (ms_b * 1e-3), 1),
                           "unit": "frames/s", "step_frac": round(w2.frac(ms_a / k2), 4),
                           "step_frac_single_stream": round(w2.frac(ms_b / k2), 4),
                           "l2": "inputs larger than L2" if w2.flush is None else "L2 flushed between steps"}
            del w2
            torch.cuda.empty_cache()

    # ---- configs[0]: one VOC frame, latency of the call (rank 0)
    latency = None
    if rank == 0 and not args.no_other:
        try:
            latency = voc_latency_leg(dev)
        except Exception as e:
            latency = {"error": str(e)[:200]}
    # ---- configs[3] with ITS arguments: the box_nms operator at valid_thresh 0.001, topk -1, force_suppress off / on
    stress = None
    if rank == 0 and not args.no_other:
        try:
            stress = stress_nms_leg(dev, peak)
        except Exception as e:
            stress = {"error": str(e)[:200]}

    # ---- cpu baseline (rank 0, N=1 only): oracle port on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        frames = min(B, max(cores, 8))
        v, dt = cpu_port_fps(C, size, frames, 1236, cores)
        reps = 1
        while dt * reps < 8.0 and reps < 8:
            reps += 1
        if reps > 1:
            v, dt = cpu_port_fps(C, size, frames, 1236, cores, steps=reps)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d of %d frames x %d passes, %d threads over frames (oracle/vy_oracle.c: decode + box_nms)"
                         % (frames, B, reps, cores)}
        # the same path as the reference's GRAPH runs it on a CPU: one multi-threaded library op per
        # MXNet op with every intermediate materialised (oracle.decode_torch_graph), then the box_nms operator
        try:
            import numpy as np
            import torch as _t
            _t.set_num_threads(cores)
            oracle.set_threads(cores)
            gf = min(frames, 4)
            rng = np.random.RandomState(1237)
            hs = [rng.standard_normal(size=(gf, 3 * (5 + C), g, g)).astype(np.float32) for g in grid_sizes(size)]
            def graph_pass():
                dets = oracle.decode_torch_graph(hs, C)
                oracle.yolov3_tail(dets, NMS["nms_thresh"], NMS["topk"], NMS["post_nms"], valid_thresh=NMS["valid_thresh"])
            t0 = time.perf_counter()
            graph_pass()
            gdt = time.perf_counter() - t0
            cpu["graph_faithful"] = {"value": gf / gdt, "unit": "frames/s", "threads": cores,
                                     "sample": "%d frames, torch-CPU op-by-op decode graph (B,R,6 materialised) + oracle box_nms" % gf}
        except Exception as e:                                   # reported, never fatal for the bench line
            cpu["graph_faithful"] = {"error": str(e)[:200]}

    if rank == 0:
        cfg = make_config(args, world)
        line = {"metric": "frames/sec decode+NMS", "value": fps, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "launch_mode": ("eager library calls" if args.no_graph else
                                "one CUDA graph per stream (the library's calls captured once: pipeline.GraphedDetector), "
                                "one input set per stream"),
                "value_single_stream": world * B * args.steps / (ms_serial * 1e-3),
                "value_eager": world * B * args.steps / (ms_eager * 1e-3),
                "value_single_stream_eager": world * B * args.steps / (ms_eager_serial * 1e-3),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "fusion_conv": conv, "other_workloads": other,
                "voc416_b1_latency": latency, "stress_nms": stress, "gpu_launches": gpu_launches,
                "launches_per_step": per_step, "clocks": sampler.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
