"""Scratch timing of the P-layout pack (fp32 (B,K,C,H,W) -> bf16 [T][B][H+2][W+2][C]) and the temporal pool at the
shapes of the temporal tail (batch 32 windows, K = 3).  CUDA events, L2 flushed.  usage: pack_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videoyolo_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(f, n=10):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[n // 2]


for c, g in ((512, 13), (256, 26), (128, 52)):
    x = torch.randn((B, 3, c, g, g), device=dev)
    ms = timed(lambda: ops.pack_p(x, "NTCHW"))
    by = x.numel() * 4 + 3 * B * (g + 2) * (g + 2) * c * 2
    print("pack  B=%d %dx%d C=%d: %.1f us  (%.0f MB -> %.0f GB/s)" % (B, g, g, c, ms * 1e3, by / 1e6, by / ms / 1e6), flush=True)
    p = ops.pack_p(torch.randn((B, 3, 2 * c, g, g), device=dev), "NTCHW")
    ms = timed(lambda: ops.temporal_pool(p, "max"))
    by = p.data.numel() * 2 * 4 // 3
    print("pool  B=%d %dx%d C=%d: %.1f us  (%.0f MB -> %.0f GB/s)" % (B, g, g, 2 * c, ms * 1e3, by / 1e6, by / ms / 1e6), flush=True)
