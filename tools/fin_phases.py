"""Per-phase cycle counts of vy_nms_finalize_kernel (CTA 0), from a -DVY_FIN_TIMING development build.
Build here (CPU container):  python tools/fin_phases.py --build     then under gpurun:  python tools/fin_phases.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--build" in sys.argv:
    from videoyolo_b200 import build
    print(build.build(extra_flags=["-DVY_FIN_TIMING"], variant="fintiming"))
    sys.exit(0)
os.environ["VYOLO_LIB_VARIANT"] = "fintiming"
import torch
import videoyolo_b200 as vy
from videoyolo_b200 import _lib
from videoyolo_b200.synth import random_heads_cuda

AN, ST = vy.ANCHORS[::-1], vy.STRIDES[::-1]
dev = torch.device("cuda:0")
L = _lib.lib()
NAMES = ["front (sort)", "general front", "-", "prefetch", "class sort", "box decode", "seg_end", "bitmask",
         "greedy scan", "prefix", "output"]
for name, B, C, size, regime in [("coco608_b64_R", 64, 80, 608, "R"), ("coco608_b64_T", 64, 80, 608, "T"),
                                 ("voc416_b1_R", 1, 20, 416, "R"), ("vid320_b256_R", 256, 30, 320, "R")]:
    heads = random_heads_cuda(B, C, size, 1234, dev, regime=regime)
    acc = [0.0] * 11
    n = 5
    for it in range(n + 2):
        vy.yolo3_decode_nms(heads, C, AN, ST)
        torch.cuda.synchronize()
        clk = (ctypes.c_longlong * 16)()
        assert L.vy_debug_fin_clocks(clk) == 0
        if it >= 2:
            for k in range(11):
                acc[k] += (clk[k + 1] - clk[k]) / n
    print("   last merge sort (rank sort when classes are counted): start->warp-sorted %d, rounds %s" % (clk[12] - clk[3], [clk[k + 1] - clk[k] for k in range(12, 15)]))
    fc = (ctypes.c_longlong * 16)()
    if L.vy_debug_fin_front_clocks(fc) == 0:
        print("   bucket front: load+zero %d | min/max %d | histogram %d | scan+K-th bin %d | scatter %d | in-bin rank %d" % tuple(
            fc[k + 1] - fc[k] for k in range(6)))
    nb = min(B, 1024)
    ct = (ctypes.c_longlong * (4 * nb))()
    assert L.vy_debug_fin_ctas(ct, nb) == 0
    rows = sorted([tuple(ct[4 * i:4 * i + 4]) for i in range(nb)])
    sms = [r[1] for r in rows]
    print("   per CTA cycles: min %d median %d max %d | distinct SMs %d of %d CTAs | slowest: %s" % (
        rows[0][0], rows[nb // 2][0], rows[-1][0], len(set(sms)), nb, rows[-3:]))
    tot = sum(acc)
    print("%s: finalize CTA 0 = %.0f cycles" % (name, tot))
    for k in range(11):
        print("   %-12s %8.0f cycles  %5.1f%%" % (NAMES[k], acc[k], 100 * acc[k] / tot))
