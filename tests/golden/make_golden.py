"""Generates the fixtures under tests/golden/.  Run in the BUILD container only:

    python tests/golden/make_golden.py

Three kinds of fixture, provenance stated per file:
  box_nms_mxnet_doc.json   HAND-TRANSCRIBED known answers of MXNet's public ``box_nms``
                           operator documentation example and ``test_box_nms_op`` unit-test
                           cases (apache/incubator-mxnet, tests/python/unittest/
                           test_contrib_operator.py).  MXNet is the un-vendored dependency that
                           executes yolo3.py:526-528; it is not installable here, so these are the
                           only reference-side answers that exist for the NMS step.
  bbox_iou_ref.npz         outputs of the REFERENCE ITSELF: /root/reference/utils/bbox.py:bbox_iou
                           imported here and run on seeded inputs (incl. degenerate boxes).
  decode_ref_*.npz         outputs of the REFERENCE ITSELF: models/definitions/yolo/yolo3.py imported
                           UNMODIFIED from /root/reference with stub mxnet/gluoncv modules
                           (tests/golden/mx_shim.py: a numpy-fp32 `F` / NDArray executing the ~15
                           array ops the file calls with MXNet's semantics).  `dets` is what
                           YOLOOutputV3.hybrid_forward (:130-199, agnostic :184-188) returns per
                           scale, concatenated by YOLOV3.hybrid_forward (:523); `ids/scores/bboxes`
                           is what YOLOV3.hybrid_forward (:448-534) returns when F.contrib.box_nms
                           is served by oracle.box_nms_c (MXNet's operator source is not in
                           /root/reference -- that one step is NOT reference-executed).  Full-size
                           tensors are 5-44 MB per frame, so those files hold every `row_step`-th
                           row + float64 column sums; the head maps are regenerated from the
                           stored seed (legacy numpy RandomState stream) and checked by checksum.
  postproc_regress_*.npz   outputs of OUR ORACLE (oracle/) on small seeded head maps: regression
                           pins so that neither the oracle nor the CUDA path can drift silently.
                           (Not reference outputs.)
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def corner_to_center(a):
    a = np.array(a, dtype=np.float64)
    out = a.copy()
    out[..., 2] = (a[..., 2] + a[..., 4]) / 2
    out[..., 3] = (a[..., 3] + a[..., 5]) / 2
    out[..., 4] = a[..., 4] - a[..., 2]
    out[..., 5] = a[..., 5] - a[..., 3]
    neg = a[..., 0] < 0            # padding rows stay -1
    out[neg] = -1
    return out.tolist()


def mxnet_doc_cases():
    boxes = [[0, 0.5, 0.1, 0.1, 0.2, 0.2], [1, 0.4, 0.1, 0.1, 0.2, 0.2],
             [0, 0.3, 0.1, 0.1, 0.14, 0.14], [2, 0.6, 0.5, 0.5, 0.7, 0.8]]
    pad = [-1] * 6
    e_force05 = [[2, 0.6, 0.5, 0.5, 0.7, 0.8], [0, 0.5, 0.1, 0.1, 0.2, 0.2],
                 [0, 0.3, 0.1, 0.1, 0.14, 0.14], pad]
    e_force01 = [[2, 0.6, 0.5, 0.5, 0.7, 0.8], [0, 0.5, 0.1, 0.1, 0.2, 0.2], pad, pad]
    e_noforce01 = [[2, 0.6, 0.5, 0.5, 0.7, 0.8], [0, 0.5, 0.1, 0.1, 0.2, 0.2],
                   [1, 0.4, 0.1, 0.1, 0.2, 0.2], pad]
    base = dict(coord_start=2, score_index=1, id_index=0)
    cases = [
        dict(name="doc_example_force_thresh0.1", data=boxes, expected=e_force01, kept=[3, 0, -1, -1],
             args=dict(overlap_thresh=0.1, force_suppress=True, **base)),
        dict(name="ut_case1_force_thresh0.5", data=boxes, expected=e_force05, kept=[3, 0, 2, -1],
             args=dict(overlap_thresh=0.5, force_suppress=True, **base)),
        dict(name="ut_case2_multibatch", data=[boxes] * 3, expected=[e_force05] * 3,
             kept=[[3, 0, 2, -1]] * 3, args=dict(overlap_thresh=0.5, force_suppress=True, **base)),
        dict(name="ut_case2_two_leading_dims", data=[[boxes] * 3] * 2, expected=[[e_force05] * 3] * 2,
             kept=[[[3, 0, 2, -1]] * 3] * 2, args=dict(overlap_thresh=0.5, force_suppress=True, **base)),
        dict(name="ut_case4_noforce_thresh0.1", data=boxes, expected=e_noforce01, kept=[3, 0, 1, -1],
             args=dict(overlap_thresh=0.1, force_suppress=False, **base)),
        dict(name="ut_case5_in_center", data=corner_to_center(boxes), expected=e_noforce01, kept=[3, 0, 1, -1],
             approx=True, args=dict(overlap_thresh=0.1, force_suppress=False, in_format="center",
                                    out_format="corner", **base)),
        dict(name="ut_case5_out_center", data=boxes, expected=corner_to_center(e_noforce01), kept=[3, 0, 1, -1],
             approx=True, args=dict(overlap_thresh=0.1, force_suppress=False, in_format="corner",
                                    out_format="center", **base)),
        dict(name="ut_case5_in_out_center", data=corner_to_center(boxes), expected=corner_to_center(e_noforce01),
             kept=[3, 0, 1, -1], approx=True,
             args=dict(overlap_thresh=0.1, force_suppress=False, in_format="center", out_format="center", **base)),
        dict(name="ut_case7_no_id_equals_force", data=[boxes] * 3, expected=[e_force05] * 3,
             kept=[[3, 0, 2, -1]] * 3,
             args=dict(overlap_thresh=0.5, force_suppress=False, coord_start=2, score_index=1, id_index=-1)),
        dict(name="ut_case8_multibatch_valid_topk",
             data=[[[1, 1, 0, 0, 10, 10], [1, 0.4, 0, 0, 10, 10], [1, 0.3, 0, 0, 10, 10]],
                   [[2, 1, 0, 0, 10, 10], [2, 0.4, 0, 0, 10, 10], [2, 0.3, 0, 0, 10, 10]],
                   [[3, 1, 0, 0, 10, 10], [3, 0.4, 0, 0, 10, 10], [3, 0.3, 0, 0, 10, 10]]],
             expected=[[[1, 1, 0, 0, 10, 10], pad, pad], [[2, 1, 0, 0, 10, 10], pad, pad],
                       [[3, 1, 0, 0, 10, 10], pad, pad]],
             kept=[[0, -1, -1]] * 3,
             args=dict(overlap_thresh=0.5, force_suppress=False, valid_thresh=0.5, topk=2, **base)),
        # SURVEY.md Appendix B.4 variants (hand-verified there with a scratch fp32 restatement)
        dict(name="b4_force_topk2", data=boxes, expected=e_force01, kept=[3, 0, -1, -1],
             args=dict(overlap_thresh=0.5, force_suppress=True, topk=2, **base)),
        dict(name="b4_valid0.45", data=boxes, expected=e_force01, kept=[3, 0, -1, -1],
             args=dict(overlap_thresh=0.5, valid_thresh=0.45, **base)),
    ]
    return cases


def ref_bbox_iou():
    spec = importlib.util.spec_from_file_location("ref_bbox", "/root/reference/utils/bbox.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.RandomState(20261017)
    out = {}

    def boxes(n, scale):
        xy = rng.uniform(0, scale, size=(n, 2))
        wh = rng.uniform(0, scale / 2, size=(n, 2))
        return np.concatenate([xy, xy + wh], axis=1)

    cases = {
        "pix": (boxes(37, 416.0), boxes(23, 416.0), 0),
        "pix_off1": (np.round(boxes(19, 416.0)), np.round(boxes(11, 416.0)), 1),
        "norm": (boxes(16, 1.0), boxes(1, 1.0), 0),                 # the reference's only call shape (M=1)
        "extra_cols": (np.concatenate([boxes(9, 100.0), rng.uniform(size=(9, 2))], 1), boxes(5, 100.0), 0),
    }
    deg_a = np.array([[0, 0, 10, 10], [5, 5, 5, 5], [10, 10, 0, 0], [0, 0, 10, 10], [20, 20, 30, 30.]])
    deg_b = np.array([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 5, 5], [2, 2, 4, 4.]])
    cases["degenerate"] = (deg_a, deg_b, 0)
    for k, (a, b, off) in cases.items():
        with np.errstate(all="ignore"):
            out[k + "_a"], out[k + "_b"], out[k + "_off"] = a, b, np.array(off)
            out[k + "_iou"] = mod.bbox_iou(a, b, off)
    np.savez_compressed(os.path.join(HERE, "bbox_iou_ref.npz"), **out)


from videoyolo_b200.synth import trained_like_heads  # noqa: E402


def oracle_regress():
    import oracle
    for name, (B, C, size, regime, seed) in {
        "voc416_random": (2, 20, 416, "R", 1235),
        "vid320_trained": (3, 30, 320, "T", 1239),
        "coco_small_trained": (2, 80, 160, "T", 1236),
    }.items():
        rng = np.random.RandomState(seed)
        if regime == "R":
            heads = [rng.normal(0, 1, size=(B, 3 * (5 + C), g, g)).astype(np.float32)
                     for g in (size // 32, size // 16, size // 8)]
        else:
            heads = trained_like_heads(rng, B, C, size)
        ids, scores, bboxes, rec = oracle.yolov3_postprocess(heads, C, return_record=True)
        np.savez_compressed(os.path.join(HERE, "postproc_regress_%s.npz" % name),
                            h0=heads[0], h1=heads[1], h2=heads[2], C=np.array(C),
                            ids=ids, scores=scores, bboxes=bboxes, kept_rows=rec)


REF_ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]  # wrappers.py:80-83
REF_STRIDES = [8, 16, 32]                                                                           # wrappers.py:84

# name -> (B, C, (H, W) of the stride-32 map, regime, seed, agnostic, row_step (0 = store every row))
DECODE_REF_CASES = {
    "voc416": (1, 20, (13, 13), "R", 1234, False, 61),
    "coco608": (1, 80, (19, 19), "R", 1235, False, 211),
    "vid320": (2, 30, (10, 10), "T", 1238, False, 37),
    "agnostic": (2, 30, (13, 13), "R", 1240, True, 3),
    "small": (2, 20, (3, 3), "R15", 1241, False, 0),
    "nonsquare": (2, 7, (3, 5), "R15", 1242, False, 0),      # H != W: pins the (x, y) offset order
}


def decode_ref_heads(name):
    """The case's head maps, regenerated from its seed (used by the generator AND by the tests)."""
    B, C, (H, W), regime, seed, agnostic, step = DECODE_REF_CASES[name]
    rng = np.random.RandomState(seed)
    if regime == "T":
        assert H == W
        return trained_like_heads(rng, B, C, H * 32)
    std = 1.5 if regime == "R15" else 1.0
    return [rng.normal(0, std, size=(B, 3 * (5 + C), H * m, W * m)).astype(np.float32) for m in (1, 2, 4)]


def reference_decode():
    """Runs the reference's own YOLOOutputV3 / YOLOV3.hybrid_forward (see module docstring)."""
    import importlib
    import oracle
    sys.path.insert(0, HERE)
    sys.path.insert(0, "/root/reference")
    import mx_shim
    mx_shim.install()
    ref = importlib.import_module("models.definitions.yolo.yolo3")
    F = mx_shim.F

    class Route(mx_shim.HybridBlock):            # backbone stage: out of scope, a placeholder feature map
        def hybrid_forward(self, F, x):
            return x

    class TipIs(mx_shim.HybridBlock):            # YOLODetectionBlockV3 stand-in: the tip IS the head map
        def __init__(self, tip):                 # (the 1x1 `prediction` conv is mx_shim.Identity)
            super().__init__()
            self.tip = mx_shim.NDArray(tip)

        def hybrid_forward(self, F, x):
            return x, self.tip

    for name, (B, C, hw, regime, seed, agnostic, step) in DECODE_REF_CASES.items():
        heads = decode_ref_heads(name)
        net = ref.YOLOV3([Route(), Route(), Route()], [512, 256, 128], REF_ANCHORS, REF_STRIDES,
                         ["c%d" % i for i in range(C)], agnostic=agnostic)
        net.set_nms(nms_thresh=0.45, nms_topk=400)                      # detect_yolo3.py:200
        net.yolo_blocks._children_list[:] = [TipIs(h) for h in heads]
        seen = {}

        def nms(data, **kw):
            seen["dets"], seen["kw"] = data.copy(), dict(kw)
            return oracle.box_nms_c(data, **kw)

        F.contrib.box_nms_impl = nms
        ids, scores, bboxes = net(mx_shim.NDArray(np.zeros((B, 3, 2, 2))))
        dets = seen["dets"]
        assert seen["kw"] == dict(overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1,
                                  coord_start=2, force_suppress=False), seen["kw"]
        # per-scale outputs straight from YOLOOutputV3 (no YOLOV3 around it) must be the same rows
        per = [ref.YOLOOutputV3(i, C, REF_ANCHORS[::-1][i], REF_STRIDES[::-1][i], agnostic=agnostic)(mx_shim.NDArray(h)).a
               for i, h in enumerate(heads)]
        assert np.array_equal(np.concatenate(per, axis=1), dets)
        R = dets.shape[1]
        rows = np.arange(0, R, step) if step else np.arange(R)
        np.savez_compressed(
            os.path.join(HERE, "decode_ref_%s.npz" % name),
            B=np.array(B), C=np.array(C), agnostic=np.array(int(agnostic)), R=np.array(R),
            heads_sum=np.array([h.astype(np.float64).sum() for h in heads]),
            rows=rows.astype(np.int64), dets_rows=dets[:, rows],
            col_sums=dets.astype(np.float64).sum(axis=1),
            ids=ids.a, scores=scores.a, bboxes=bboxes.a)
        print("decode_ref_%s: dets %s, %d rows stored, %d detections in frame 0"
              % (name, dets.shape, len(rows), int((ids.a[0] >= 0).sum())))


if __name__ == "__main__":
    with open(os.path.join(HERE, "box_nms_mxnet_doc.json"), "w") as f:
        json.dump({"provenance": "hand-transcribed from MXNet public box_nms docs + test_box_nms_op; "
                                 "see make_golden.py docstring", "cases": mxnet_doc_cases()}, f, indent=1)
    ref_bbox_iou()
    oracle_regress()
    reference_decode()
    print("golden fixtures written to", HERE)
