"""Statistics of the head maps the temporal-neck leg of bench.py produces (calibrated BatchNorm), and the fused
decode + NMS call timed on exactly those maps (warm and flushed L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.layers import _Cell
import bench

dev = torch.device("cuda:0")
torch.manual_seed(9)
B, K, C = 8, 3, 30
net = vy.YOLOV3TNeck(["c%d" % i for i in range(C)], k=K, k_join_type="max", block_conv_type="3").to(dev).eval()
first = set(id(next(m for m in blk.modules() if isinstance(m, _Cell))) for blk in net.blocks)
for m in net.modules():
    if isinstance(m, _Cell):
        m.running_var.fill_(m.weight[0].numel() * (0.07 ** 2 / 3.0) * (1.0 if id(m) in first else 0.505))
rs = [torch.randn((B, K, c, g, g), device=dev) for c, g in zip((1024, 512, 256), bench.grid_sizes(416))]
with torch.no_grad():
    heads = net.head.head_maps(*net.routes(*rs))
for h in heads:
    v = h.view(B, 3, 5 + C, -1)
    print(tuple(h.shape), "obj std %.3f mean %.3f | cls std %.3f mean %.3f | max %.2f" % (
        v[:, :, 4].std().item(), v[:, :, 4].mean().item(), v[:, :, 5:].std().item(), v[:, :, 5:].mean().item(), h.abs().max().item()))
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("warm", "flushed"):
    for _ in range(3):
        vy.yolo3_decode_nms(heads, C, AN, ST)
    _lib.prof_enable(True); _lib.prof_read()
    for _ in range(10):
        if mode == "flushed":
            flush.zero_()
        vy.yolo3_decode_nms(heads, C, AN, ST)
    torch.cuda.synchronize()
    r = _lib.prof_read(); _lib.prof_enable(False)
    print(mode, {k: round(v[0] / v[1] * 1e3, 1) for k, v in r.items()})
