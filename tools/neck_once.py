"""One forward of the temporal neck (bench.py temporal_neck_leg workload) for an ncu launch list.  usage: neck_once.py [B] [calls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoyolo_b200 as vy
import bench

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(7)
net = vy.YOLOV3TNeck(["c%d" % i for i in range(30)], k=3, k_join_type="max", block_conv_type="3").to(dev).eval()
xs = [torch.randn((B, 3, c, g, g), device=dev) for c, g in zip((1024, 512, 256), bench.grid_sizes(416))]
with torch.no_grad():
    for _ in range(calls):
        net(*xs)
torch.cuda.synchronize()
