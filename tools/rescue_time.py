"""What does the in-finalize rescue cost?  (a) the worst case: all-equal scores, EVERY image rescued (the streamed lists
overflow, radix passes over row bits); (b) VY_FORCE_RESCUE=1: every image rescued on ordinary inputs -- trained-like
logits end in the first pass (all valid keys), random-init ones take the radix passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
forced = os.environ.get("VY_FORCE_RESCUE") is not None
for name, B, C, size in (("coco608_b64", 64, 80, 608), ("stress416_b128", 128, 80, 416), ("vid320_b256", 256, 30, 320), ("voc416_b1", 1, 20, 416)):
    for kind in (("R", "T") if forced else ("zeros",)):
        heads = random_heads_cuda(B, C, size, 1, dev, regime="T" if kind == "T" else "R")
        if kind == "zeros":
            for h in heads: h.zero_()
        for _ in range(2): vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): vy.yolo3_decode_nms(heads, C, AN, ST)
        b.record(); torch.cuda.synchronize()
        print("%-14s %-5s all images rescued: %.1f us per call" % (name, kind, a.elapsed_time(b) / 5 * 1e3), flush=True)
        del heads
