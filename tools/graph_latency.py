"""Latency of one fused decode+NMS call at BASELINE configs[0] (VOC 20 cls, 416^2, batch 1): eager vs CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200.pipeline import GraphedDetector
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
for B, C, size in [(1, 20, 416), (8, 20, 416)]:
    heads = random_heads_cuda(B, C, size, 3, dev)
    det = GraphedDetector(C, AN, ST, [tuple(h.shape) for h in heads], dev)
    det(heads)
    def wall(fn, n=200):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6
    def one(fn, n=100):     # launch-to-result latency of a single isolated call
        ts = []
        for _ in range(n):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e6)
        return sorted(ts)[n // 2]
    e = lambda: vy.yolo3_decode_nms(heads, C, AN, ST, out=det.out, kept=det.kept)
    g = lambda: det()
    print("B=%d VOC416: back-to-back eager %.1f us/call, graph %.1f us/call | isolated call eager %.1f us, graph %.1f us"
          % (B, wall(e), wall(g), one(e), one(g)), flush=True)
