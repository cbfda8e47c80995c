"""Per-kernel times of the fused call (library events, vy_prof_*) over the benchmarked workloads, plus the whole call
back to back on one stream.  A/B knobs are environment variables read by the library (VY_STREAM_V1, VY_S2_SMEM_KB, ...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
cfgs = [("coco608_b64", 64, 80, 608), ("stress416_b128", 128, 80, 416), ("vid320_b256", 256, 30, 320), ("vid416_b32", 32, 30, 416),
        ("vid416_b8", 8, 30, 416), ("voc416_b1", 1, 20, 416)]
only = sys.argv[1:] or None
for name, B, C, size in cfgs:
    if only and name not in only: continue
    for kind in ("R", "T", "S", "nohits"):
        heads = random_heads_cuda(B, C, size, 1234, dev, regime=kind if kind in ("T", "S") else "R")
        if kind == "nohits":
            for h in heads: h.fill_(-30.0)
        for _ in range(5): vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): vy.yolo3_decode_nms(heads, C, AN, ST)
        b.record(); torch.cuda.synchronize()
        call_us = a.elapsed_time(b) / 50 * 1e3
        _lib.prof_enable(True); _lib.prof_read()
        for _ in range(20): vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        r = _lib.prof_read(); _lib.prof_enable(False)
        nbytes = sum(h.numel() * 4 for h in heads)
        ks = {k.replace("vy_", "").replace("_kernel", ""): round(v[0] / v[1] * 1e3, 1) for k, v in r.items()}
        st = ks.get("decode_stream", 0.0)
        print("%-15s %-6s call %.1f us (%.0f GB/s) | stream %.1f us -> %.0f GB/s | %s" % (
            name, kind, call_us, nbytes / call_us / 1e3, st, nbytes / st / 1e3 if st else 0, ks), flush=True)
        del heads
