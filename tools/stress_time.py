"""BASELINE config 4 at its named batch: 80 cls, valid_thresh 0.001, topk -1, force_suppress off/on, 10647 boxes x batch 128."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
EXH = len(sys.argv) > 2
heads = random_heads_cuda(B, 80, 416, 5, dev)
dets = vy.yolo3_decode(heads, 80, AN, ST)
for force in (False, True):
    fn = lambda: vy.box_nms(dets, 0.45, 0.001, -1, id_index=0, force_suppress=force, _exhaustive=EXH)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("stress box_nms B=%d R=%d valid_thresh=0.001 topk=-1 force_suppress=%s: %.1f ms -> %.1f frames/s, survivors/frame %.0f"
          % (B, dets.shape[1], force, min(ts), B / min(ts) * 1e3, float((out[..., 0] >= 0).sum()) / B), flush=True)
    del out
