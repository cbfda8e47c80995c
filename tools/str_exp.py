"""Per-kernel times of the fused path (library events), for kernel experiments.
usage: [VY_STR_MODE=0|1] str_exp.py [valid_thresh] [config]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.synth import random_heads_cuda
valid = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
cfg = sys.argv[2] if len(sys.argv) > 2 else "coco"
B, C, size = {"coco": (64, 80, 608), "vid": (256, 30, 320), "stress": (128, 80, 416)}[cfg]
dev = torch.device("cuda:0")
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
heads = random_heads_cuda(B, C, size, 1236, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    vy.yolo3_decode_nms(heads, C, AN, ST, valid_thresh=valid)
_lib.prof_enable(True); _lib.prof_read()
n = 30
tot = 0.0
for _ in range(n):
    if os.environ.get("VY_FLUSH"): flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); vy.yolo3_decode_nms(heads, C, AN, ST, valid_thresh=valid); b.record(); torch.cuda.synchronize()
    tot += a.elapsed_time(b)
prof = _lib.prof_read(); _lib.prof_enable(False)
print("mode=%s valid=%g %s: step %.1f us | " % (os.environ.get("VY_STR_MODE", "0"), valid, cfg, tot / n * 1e3) +
      ", ".join("%s %.1f us" % (k.replace("vy_", "").replace("_kernel", ""), v[0] / max(v[1], 1) * 1e3) for k, v in prof.items()), flush=True)
