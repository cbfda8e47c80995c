"""How long does a call take when EVERY image needs the in-finalize rescue?  (all-equal scores: every key passes the
sampled bound, the streamed lists overflow.)  The worst case of fin_rescue_heads; never met with real logits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
for name, B, C, size in (("coco608_b64", 64, 80, 608), ("vid320_b256", 256, 30, 320), ("voc416_b1", 1, 20, 416)):
    heads = random_heads_cuda(B, C, size, 1, dev)
    for h in heads: h.zero_()
    for _ in range(2): vy.yolo3_decode_nms(heads, C, AN, ST)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): vy.yolo3_decode_nms(heads, C, AN, ST)
    b.record(); torch.cuda.synchronize()
    print("%-12s all images rescued: %.1f us per call" % (name, a.elapsed_time(b) / 5 * 1e3), flush=True)
