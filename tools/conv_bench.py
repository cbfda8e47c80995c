"""Scratch timing of the temporal fusion conv (tcgen05 implicit GEMM) at the canonical K=3 tip-conv
shapes of BASELINE configs[2].  CUDA events, L2 flushed.  Not the bench contract.
usage: conv_bench.py [B] [one]      (one: a single launch of each of the three tip shapes, for ncu)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
one = len(sys.argv) > 2
peak = 1638.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(13, 512, 1024, (3, 3, 3)), (26, 256, 512, (3, 3, 3)), (52, 128, 256, (3, 3, 3)),
          (26, 256, 512, (1, 3, 3)), (26, 512, 512, (3, 1, 1)), (26, 768, 256, (1, 1, 1))]
if one:
    shapes = shapes[:3]
T = 3
for g, Cin, Cout, k3 in shapes:
    x = ops.PTensor(torch.zeros((T, B, g + 2, g + 2, Cin), dtype=torch.bfloat16, device=dev), B, T, g, g, Cin)
    x.data[:, :, 1:-1, 1:-1] = torch.randn((T, B, g, g, Cin), device=dev).to(torch.bfloat16)
    w = (torch.rand((Cout,) + k3 + (Cin,), device=dev) * 0.14 - 0.07).to(torch.bfloat16)
    sc, sh = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
    fl = 2.0 * B * T * g * g * Cout * Cin * k3[0] * k3[1] * k3[2]
    if k3[0] == 3:   # temporal 'same' padding: edge frames see 2 of 3 taps -- count real MACs only? no: SURVEY 8d formula
        pass
    n = 1 if one else 10
    for _ in range(0 if one else 3):
        ops.fusion_conv(x, w, sc, sh)
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fusion_conv(x, w, sc, sh); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    med = ts[len(ts) // 2]
    print("conv B=%d g=%d Cin=%d Cout=%d k=%s: median %.3f ms best %.3f ms -> %.1f TFLOP/s (formula), %.1f%% of %.0f"
          % (B, g, Cin, Cout, k3, med, ts[0], fl / med / 1e9, 100 * fl / med / 1e9 / peak, peak), flush=True)
