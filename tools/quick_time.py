"""Scratch timing of the hot-path kernels (CUDA events, L2-flushed). Not the bench contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200.synth import random_heads_cuda

AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]

only = sys.argv[1] if len(sys.argv) > 1 else None
for name, B, C, size, regime in [("coco608_b64_R", 64, 80, 608, "R"), ("coco608_b64_T", 64, 80, 608, "T"),
                                 ("voc416_b1_R", 1, 20, 416, "R"), ("stress416_b128_R", 128, 80, 416, "R"),
                                 ("vid320_b256_R", 256, 30, 320, "R"), ("vid320_b256_T", 256, 30, 320, "T")]:
    if only and name != only: continue
    heads = random_heads_cuda(B, C, size, 1234, dev, regime=regime)
    nbytes = sum(h.numel() * 4 for h in heads)
    med, best = timeit(lambda: vy.yolo3_decode_nms(heads, C, AN, ST))
    print("%-18s fused decode+nms: median %.3f ms best %.3f ms | %.1f MB in -> %.0f GB/s (best), %.0f frames/s"
          % (name, med, best, nbytes / 1e6, nbytes / best / 1e6, B / med * 1e3), flush=True)
    del heads
if only: sys.exit(0)
heads = random_heads_cuda(8, 80, 608, 1, dev)
med, best = timeit(lambda: vy.yolo3_decode(heads, 80, AN, ST))
out_b = 8 * 1819440 * 24
print("decode full dets B=8 coco608: median %.3f ms, out %.0f MB -> %.0f GB/s" % (med, out_b / 1e6, out_b / med / 1e6))
dets = vy.yolo3_decode(heads, 80, AN, ST)
med, best = timeit(lambda: vy.box_nms(dets, 0.45, 0.01, 400, id_index=0, out_rows=100))
print("box_nms rows B=8 coco608 (topk 400, out_rows 100): median %.3f ms -> %.0f GB/s read" % (med, out_b / med / 1e6))
# BASELINE config 4 arguments: 80 cls, valid_thresh 0.001, topk -1, force_suppress on/off, 10647 boxes
for Bs in (4, 32):
    heads = random_heads_cuda(Bs, 80, 416, 5, dev)
    dets = vy.yolo3_decode(heads, 80, AN, ST)
    for force in (False, True):
        fn = lambda: vy.box_nms(dets, 0.45, 0.001, -1, id_index=0, force_suppress=force)
        med, best = timeit(fn, n=3, warm=1)
        out = fn()
        print("stress box_nms B=%d R=%d topk=-1 force=%s: median %.2f ms -> %.1f frames/s, survivors/frame %.0f"
              % (Bs, dets.shape[1], force, med, Bs / med * 1e3, float((out[..., 0] >= 0).sum()) / Bs), flush=True)
