"""How much of the streaming kernel's time is the (rare) hit path?  Same shapes, logits that can never pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
for name, B, C, size in [("coco608_b64", 64, 80, 608), ("vid320_b256", 256, 30, 320)]:
    for kind in ("R", "nohits", "T"):
        heads = random_heads_cuda(B, C, size, 1234, dev, regime="T" if kind == "T" else "R")
        if kind == "nohits":
            for h in heads: h.fill_(-30.0)
        for _ in range(3): vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        _lib.prof_enable(True); _lib.prof_read()
        for _ in range(20): vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        r = _lib.prof_read(); _lib.prof_enable(False)
        nbytes = sum(h.numel() * 4 for h in heads)
        ms = r["vy_decode_stream_kernel"][0] / r["vy_decode_stream_kernel"][1]
        print("%s %-7s stream kernel %.2f us -> %.0f GB/s algorithmic | %s" % (
            name, kind, ms * 1e3, nbytes / ms / 1e6, {k: round(v[0] / v[1] * 1e3, 1) for k, v in r.items()}), flush=True)
        del heads
