"""The fusion-conv legs of bench.py alone (tip convs at batch 8 and 32, temporal tail, temporal neck).
usage: conv_legs.py      (VY_CONV_CTA2=0: one-CTA kernel only; 128 / 256: that CTA-pair width wherever it divides Cout)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for B in (8, 32):
    leg = bench.fusion_conv_leg(dev, B=B)
    print("tips B=%d" % B, json.dumps(leg), flush=True)
print("tail", json.dumps(bench.temporal_tail_leg(dev)), flush=True)
print("neck", json.dumps(bench.temporal_neck_leg(dev)), flush=True)
print("neck, constructor BN", json.dumps(bench.temporal_neck_leg(dev, calibrated_bn=False)), flush=True)
