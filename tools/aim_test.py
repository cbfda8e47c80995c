"""Fill of the streamed candidate lists and rescued images per call (VY_DEBUG_LISTS) over seeds and both logit regimes;
run with VY_SAMP_AIM=<rank aimed at, in units of K> and development builds -DVY_SAMP_RUN=<items per sampled run>.
usage: aim_test.py [n_seeds] [config ...]"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import videoyolo_b200 as vy
from videoyolo_b200.synth import random_heads_cuda
AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
os.environ["VY_DEBUG_LISTS"] = "1"
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
only = sys.argv[2:]
for name, B, C, size in (("vid320_b256", 256, 30, 320), ("coco608_b64", 64, 80, 608), ("stress416_b128", 128, 80, 416),
                          ("coco416_b320", 320, 80, 416)):          # (one sample job per image)
    if only and name not in only: continue
    for kind in (os.environ.get("AIM_REGIMES", "R,T").split(",")):
        for seed in range(1, n_seeds + 1):
            heads = random_heads_cuda(B, C, size, seed, dev, regime=kind)
            print(name, kind, seed, end=" ", flush=True)
            vy.yolo3_decode_nms(heads, C, AN, ST); torch.cuda.synchronize()
            del heads
