"""Fraction of kept rows the fused GPU path shares with the CPU oracle fixtures (tests/golden/postproc_regress_*.npz)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import videoyolo_b200 as vy
for name in ["voc416_random", "vid320_trained", "coco_small_trained"]:
    z = np.load(os.path.join("tests/golden", "postproc_regress_%s.npz" % name))
    C = int(z["C"])
    net = vy.get_yolov3_postprocess(["c"] * C)
    ids, scores, bboxes = net(*[torch.from_numpy(z[k]).cuda() for k in ("h0", "h1", "h2")])
    kept = net.last_kept_rows.cpu().numpy()
    print(name, "identical kept rows: %d of %d = %.6f" % ((kept == z["kept_rows"]).sum(), kept.size, (kept == z["kept_rows"]).mean()))
