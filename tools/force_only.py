"""One class-agnostic box_nms call at BASELINE configs[3] arguments (for ncu captures of lg_nms_kernel): force_only.py [B]"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import videoyolo_b200 as vy
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
heads = random_heads_cuda(B, 80, 416, 5, dev)
dets = vy.yolo3_decode(heads, 80, AN, ST)
out = vy.box_nms(dets, 0.45, 0.001, -1, id_index=0, force_suppress=True)
torch.cuda.synchronize()
print("ok", float((out[..., 0] >= 0).sum()) / B)
