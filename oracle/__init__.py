"""CPU ORACLE -- test infrastructure, NOT product code.

A CPU restatement of the arithmetic on VideoYOLO's per-frame detection
post-processing path (anchor decode -> temporal fusion conv -> box_nms).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``videoyolo_b200`` never does.

PARITY: /root/reference holds no tests or golden vectors of its own, and its
arithmetic runs in un-vendored, un-pinned ``mxnet-cu100`` / ``gluoncv``
(requirements.txt:1-2), neither installable here.  What pins the oracle:
  * decode, the YOLOV3 tail, bbox_iou, hierarchical_nms, VOCMApMetric.update:
    PINNED BY THE REFERENCE'S OWN CODE, executed in the build container --
    tests/golden/make_golden.py imports /root/reference's yolo3.py /
    utils/bbox.py / metrics/pascalvoc.py unmodified (and cuts hierarchical_nms
    out of detect_yolo3.py) under a numpy-fp32 stand-in for the few MXNet array
    ops they call (tests/golden/mx_shim.py); outputs committed as
    tests/golden/decode_ref_*.npz, bbox_iou_ref.npz, consumer_ref.npz;
  * box_nms: PARITY UNPINNED by the reference (the operator's source lives in
    MXNet, absent from /root/reference).  Restated from MXNet's published
    algorithm, anchored on the reference's call site
    (yolo3.py:526-528) and pinned by tests/golden/box_nms_mxnet_doc.json (known
    answers of MXNet's public documentation / unit test, hand transcribed), an
    independently structured python twin (box_nms_py) and torchvision.ops.nms
    per class (same IoU formula).

Functions and what they follow:
  decode_numpy        models/definitions/yolo/yolo3.py:151-199 op for op (numpy fp32)
  decode_c            same, C (oracle/vy_oracle.c), writes the concatenated tensor (yolo3.py:523)
  box_nms_c/_py       F.contrib.box_nms as called at yolo3.py:525-530
  yolov3_tail         yolo3.py:523-534 (concat -> box_nms -> slice post_nms -> split)
  bbox_iou            utils/bbox.py:11-38
  conv_bn_leaky       models/definitions/layers.py:63-89,135-158 (torch CPU fp32 engine)
  bbox_batch_iou      gluoncv BBoxBatchIOU as called at models/definitions/yolo/yolo_target.py:171,202
  anchor_match        yolo_target.py:86-94 (nd.contrib.box_iou of zero-centred anchors / ground truths + argmax)
  decode_torch_graph  the Gluon graph of yolo3.py:158-197,:523 op by op on torch CPU (bench.py's graph-faithful CPU figure)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# reference constants: models/definitions/yolo/wrappers.py:80-84, used reversed (yolo3.py:416-417)
ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
STRIDES = [8, 16, 32]


def build(force: bool = False) -> str:
    """Compile oracle/vy_oracle.c -> oracle/libvy_oracle.so (gcc, seconds)."""
    so = os.path.join(_HERE, "libvy_oracle.so")
    src = os.path.join(_HERE, "vy_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        L.vy_oracle_decode_f32.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_float, f32p, ctypes.c_int, f32p,
                                           ctypes.c_long, ctypes.c_long]
        L.vy_oracle_decode_f32.restype = None
        L.vy_oracle_box_nms_f32.argtypes = [f32p, ctypes.c_int, ctypes.c_long, ctypes.c_int,
                                            ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            f32p, ctypes.POINTER(ctypes.c_int32)]
        L.vy_oracle_box_nms_f32.restype = ctypes.c_int
        f64p = ctypes.POINTER(ctypes.c_double)
        L.vy_oracle_bbox_iou_f64.argtypes = [f64p, ctypes.c_int, f64p, ctypes.c_int, ctypes.c_double, f64p]
        L.vy_oracle_bbox_iou_f64.restype = None
        L.vy_oracle_set_threads.argtypes = [ctypes.c_int]
        L.vy_oracle_get_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def set_threads(n: int) -> None:
    lib().vy_oracle_set_threads(int(n))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t))


# --------------------------------------------------------------------------- decode
def grid_sizes(size: int):
    """Head grids in the order the network emits them: strides 32, 16, 8 (yolo3.py:416-417)."""
    return [size // 32, size // 16, size // 8]


def n_boxes(grids: Sequence[int], A: int = 3) -> int:
    return sum(g * g * A for g in grids)


def decode_numpy(pred: np.ndarray, anchors, stride: float, num_class: int, agnostic: bool = False):
    """Op-for-op numpy fp32 restatement of YOLOOutputV3.hybrid_forward, yolo3.py:151-199.

    pred: (B, A*P, H, W) -- the output of the 1x1 ``prediction`` conv (:157).
    returns (B, C*H*W*A, 6) rows [cls, score, x1, y1, x2, y2]; agnostic: (B, H*W*A, 6).
    """
    pred = _f32(pred)
    anchors = np.asarray(anchors, dtype=np.float32).reshape(1, 1, -1, 2)            # :46,:64
    A = anchors.shape[2]
    P = 5 + num_class
    B, _, H, W = pred.shape
    p = pred.reshape(B, A * P, H * W)                                              # :158
    p = p.transpose(0, 2, 1).reshape(B, H * W, A, P)                               # :160
    raw_centers, raw_scales = p[..., 0:2], p[..., 2:4]                             # :162-163
    objness, class_pred = p[..., 4:5], p[..., 5:]                                  # :164-165
    gx, gy = np.meshgrid(np.arange(W), np.arange(H))                               # :67-69
    offsets = np.concatenate((gx[:, :, None], gy[:, :, None]), axis=-1)            # :71
    offsets = offsets.reshape(1, -1, 1, 2).astype(np.float32)                      # :168-170
    one = np.float32(1.0)
    sig = lambda v: one / (one + np.exp(-v, dtype=np.float32))                     # mshadow sigmoid
    box_centers = (sig(raw_centers) + offsets) * np.float32(stride)                # :172
    box_scales = np.exp(raw_scales, dtype=np.float32) * anchors                    # :173
    confidence = sig(objness)                                                      # :174
    class_score = sig(class_pred) * confidence                                     # :175
    wh = box_scales / np.float32(2.0)                                              # :176
    bbox = np.concatenate((box_centers - wh, box_centers + wh), axis=-1)           # :177
    if agnostic:                                                                   # :184-188
        ids = confidence * 0
        return np.concatenate((ids, confidence, bbox), axis=-1).reshape(B, -1, 6)
    bboxes = np.tile(bbox, (num_class, 1, 1, 1, 1))                                # :191
    scores = class_score.transpose(3, 0, 1, 2)[..., None]                          # :192
    ids = scores * 0 + np.arange(num_class, dtype=np.float32).reshape(-1, 1, 1, 1, 1)  # :194
    det = np.concatenate((ids, scores, bboxes), axis=-1)                           # :195
    return det.transpose(1, 0, 2, 3, 4).reshape(B, -1, 6).astype(np.float32)       # :197


def decode_torch_graph(heads, num_class: int, strides=None, anchors=None):
    """The Gluon GRAPH of the reference restated operator by operator on torch CPU tensors (one multi-threaded library
    op per MXNet op, every intermediate materialised, the C-times tiled box tensor and the scale concat included:
    yolo3.py:158-197, :523) -- what an MXNet CPU run of the same graph executes, as opposed to the fused C loop of
    decode_c.  Used by bench.py's cpu_baseline as the "graph-faithful" figure.  Returns (B, R, 6)."""
    import torch
    strides = list(strides) if strides is not None else STRIDES[::-1]
    anchors = list(anchors) if anchors is not None else ANCHORS[::-1]
    dets = []
    for h, stride, an in zip(heads, strides, anchors):
        pred = torch.as_tensor(h, dtype=torch.float32)
        anc = torch.tensor(an, dtype=torch.float32).reshape(1, 1, -1, 2)
        A, P = anc.shape[2], 5 + num_class
        B, _, H, W = pred.shape
        p = pred.reshape(B, A * P, H * W).transpose(1, 2).reshape(B, H * W, A, P)          # :158-160
        raw_centers, raw_scales = p[..., 0:2], p[..., 2:4]                                 # :162-163
        objness, class_pred = p[..., 4:5], p[..., 5:]                                      # :164-165
        gy, gx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        offsets = torch.stack((gx, gy), dim=-1).reshape(1, -1, 1, 2)                       # :67-74, :168-170
        box_centers = (torch.sigmoid(raw_centers) + offsets) * float(stride)               # :172
        box_scales = torch.exp(raw_scales) * anc                                           # :173
        confidence = torch.sigmoid(objness)                                                # :174
        class_score = torch.sigmoid(class_pred) * confidence                               # :175
        wh = box_scales / 2.0                                                              # :176
        bbox = torch.cat((box_centers - wh, box_centers + wh), dim=-1)                     # :177
        bboxes = bbox.unsqueeze(0).repeat(num_class, 1, 1, 1, 1)                           # :191 tile
        scores = class_score.permute(3, 0, 1, 2).unsqueeze(-1)                             # :192
        ids = scores * 0 + torch.arange(num_class, dtype=torch.float32).reshape(-1, 1, 1, 1, 1)   # :194
        det = torch.cat((ids, scores, bboxes), dim=-1)                                     # :195
        dets.append(det.permute(1, 0, 2, 3, 4).reshape(B, -1, 6))                          # :197
    return torch.cat(dets, dim=1).numpy()                                                  # :523


def decode_c(heads: Sequence[np.ndarray], num_class: int, strides=None, anchors=None,
             agnostic: bool = False) -> np.ndarray:
    """Decode the three head maps (order: stride 32, 16, 8) and concatenate along rows
    exactly like yolo3.py:523.  Returns (B, R, 6) fp32."""
    strides = list(strides) if strides is not None else STRIDES[::-1]
    anchors = list(anchors) if anchors is not None else ANCHORS[::-1]
    B = heads[0].shape[0]
    A = len(anchors[0]) // 2
    C = 1 if agnostic else num_class
    R = sum(C * h.shape[2] * h.shape[3] * A for h in heads)
    out = np.empty((B, R, 6), dtype=np.float32)
    off = 0
    L = lib()
    for h, st, an in zip(heads, strides, anchors):
        h = _f32(h)
        an = _f32(an)
        assert h.shape[1] == A * (5 + num_class), (h.shape, A, num_class)
        L.vy_oracle_decode_f32(_p(h), B, A, num_class, h.shape[2], h.shape[3], float(st), _p(an),
                               int(agnostic), _p(out), R, off)
        off += C * h.shape[2] * h.shape[3] * A
    return out


# --------------------------------------------------------------------------- box_nms
_FMT = {"corner": 0, "center": 1}


def box_nms_c(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1,
              id_index=-1, background_id=-1, force_suppress=False, in_format="corner",
              out_format="corner", return_record=False):
    """MXNet ``_contrib_box_nms`` semantics (C).  Leading dims are batch (yolo3_temporal.py:545)."""
    d = _f32(data)
    shp = d.shape
    R, W = shp[-2], shp[-1]
    B = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
    out = np.empty((B, R, W), dtype=np.float32)
    rec = np.empty((B, R), dtype=np.int32)
    rc = lib().vy_oracle_box_nms_f32(_p(d), B, R, W, overlap_thresh, valid_thresh, int(topk),
                                     coord_start, score_index, id_index, background_id,
                                     int(bool(force_suppress)), _FMT[in_format], _FMT[out_format],
                                     _p(out), _p(rec, ctypes.c_int32))
    if rc != 0:
        raise ValueError("vy_oracle_box_nms_f32 rc=%d" % rc)
    out = out.reshape(shp)
    return (out, rec.reshape(shp[:-1])) if return_record else out


def box_nms_py(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1,
               id_index=-1, background_id=-1, force_suppress=False, in_format="corner",
               out_format="corner", return_record=False):
    """Independently structured python twin of box_nms_c (numpy fp32 scalars; small inputs only)."""
    d = _f32(data)
    shp = d.shape
    R, W = shp[-2], shp[-1]
    d3 = d.reshape(-1, R, W)
    out = np.full_like(d3, -1.0)
    rec = np.full(d3.shape[:2], -1, dtype=np.int32)
    f = np.float32

    def corners(b):
        if in_format == "corner":
            return b[0], b[1], b[2], b[3]
        hw, hh = b[2] / f(2), b[3] / f(2)
        return b[0] - hw, b[1] - hh, b[0] + hw, b[1] + hh

    for bi in range(d3.shape[0]):
        img = d3[bi]
        sc = img[:, score_index]
        valid = sc > f(valid_thresh)
        if id_index >= 0 and background_id >= 0:
            valid &= img[:, id_index].astype(np.int64) != background_id
        rows = np.nonzero(valid)[0]
        order = rows[np.argsort(-sc[rows], kind="stable")]      # desc, ties -> lower row first
        k = R if topk < 0 else min(R, topk)
        order = order[:k]
        boxes = img[order, coord_start:coord_start + 4]
        if in_format == "corner":
            w_, h_ = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        else:
            w_, h_ = boxes[:, 2], boxes[:, 3]
        area = np.where((w_ < 0) | (h_ < 0), f(0), w_ * h_).astype(np.float32)
        ids = img[order, id_index].astype(np.int64) if id_index >= 0 else None
        alive = np.ones(len(order), dtype=bool)
        for r in range(len(order)):
            if not alive[r]:
                continue
            rx1, ry1, rx2, ry2 = corners(boxes[r])
            for t in range(r + 1, len(order)):
                if not alive[t]:
                    continue
                if not force_suppress and ids is not None and ids[r] != ids[t]:
                    continue
                tx1, ty1, tx2, ty2 = corners(boxes[t])
                iw = max(f(0), min(rx2, tx2) - max(rx1, tx1))
                ih = max(f(0), min(ry2, ty2) - max(ry1, ty1))
                inter = f(iw * ih)
                with np.errstate(divide="ignore", invalid="ignore"):
                    iou = inter / f(f(area[r] + area[t]) - inter)
                if iou > f(overlap_thresh):
                    alive[t] = False
        kept = order[alive]
        out[bi, :len(kept)] = img[kept]
        rec[bi, :len(kept)] = kept
        if in_format != out_format:
            for j in range(len(kept)):
                c = out[bi, j, coord_start:coord_start + 4]
                if c[0] < 0:
                    continue
                if out_format == "center":
                    l, t, r2, b2 = c.copy()
                    c[:] = ((l + r2) / f(2), (t + b2) / f(2), r2 - l, b2 - t)
                else:
                    x, y, w2, h2 = c.copy()
                    c[:] = (x - w2 / f(2), y - h2 / f(2), x + w2 / f(2), y + h2 / f(2))
    out = out.reshape(shp)
    return (out, rec.reshape(shp[:-1])) if return_record else out


def yolov3_tail(dets: np.ndarray, nms_thresh=0.45, nms_topk=400, post_nms=100,
                valid_thresh=0.01, force_suppress=False, return_record=False):
    """yolo3.py:523-534: box_nms on the concatenated detections, slice post_nms, split."""
    result, rec = dets, None
    if 0 < nms_thresh < 1:                                                          # :525
        result, rec = box_nms_c(dets, overlap_thresh=nms_thresh, valid_thresh=valid_thresh,
                                topk=nms_topk, id_index=0, score_index=1, coord_start=2,
                                force_suppress=force_suppress, return_record=True)  # :526-528
        if post_nms > 0:                                                            # :529-530
            result, rec = result[:, :post_nms], rec[:, :post_nms]
    ids, scores, bboxes = result[..., 0:1], result[..., 1:2], result[..., 2:6]      # :531-533
    return (ids, scores, bboxes, rec) if return_record else (ids, scores, bboxes)


def yolov3_postprocess(heads, num_class, nms_thresh=0.45, nms_topk=400, post_nms=100,
                       valid_thresh=0.01, agnostic=False, strides=None, anchors=None,
                       return_record=False):
    """heads (stride 32,16,8 order) -> (ids, scores, bboxes): decode + yolo3.py:523-534."""
    dets = decode_c(heads, num_class, strides=strides, anchors=anchors, agnostic=agnostic)
    return yolov3_tail(dets, nms_thresh, nms_topk, post_nms, valid_thresh, return_record=return_record)


# --------------------------------------------------------------------------- bbox_iou
def bbox_iou(bbox_a: np.ndarray, bbox_b: np.ndarray, offset=0) -> np.ndarray:
    """utils/bbox.py:11-38 restated (C, float64 like numpy float64 inputs)."""
    a = np.ascontiguousarray(bbox_a, dtype=np.float64)
    b = np.ascontiguousarray(bbox_b, dtype=np.float64)
    if a.shape[1] < 4 or b.shape[1] < 4:                                            # :29-30
        raise IndexError("Bounding boxes axis 1 must have at least length 4")
    a4 = np.ascontiguousarray(a[:, :4])
    b4 = np.ascontiguousarray(b[:, :4])
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    f64p = ctypes.POINTER(ctypes.c_double)
    with np.errstate(all="ignore"):
        lib().vy_oracle_bbox_iou_f64(a4.ctypes.data_as(f64p), a.shape[0], b4.ctypes.data_as(f64p),
                                     b.shape[0], float(offset), out.ctypes.data_as(f64p))
    return out


# --------------------------------------------------------------------------- fusion conv
def bbox_batch_iou(a: np.ndarray, b: np.ndarray, offset=0.0, eps=1e-15) -> np.ndarray:
    """gluoncv.nn.bbox.BBoxBatchIOU (corner format) as called at models/definitions/yolo/yolo_target.py:171,202.
    gluoncv is an un-vendored, unpinned dependency of the reference (requirements.txt:2): restated from its published
    source -- clip(min(r) - max(l) + offset, 0, float16.max) per axis, i / (area_a + area_b - i + eps) -- fp32 op for op.
    a (B, N, 4), b (B, M, 4) -> (B, N, M)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    off, e = np.float32(offset), np.float32(eps)
    al, at, ar, ab = (a[..., k] for k in range(4))
    bl, bt, br, bb = (b[..., k] for k in range(4))
    left = np.maximum(al[..., :, None], bl[..., None, :])
    right = np.minimum(ar[..., :, None], br[..., None, :])
    top = np.maximum(at[..., :, None], bt[..., None, :])
    bot = np.minimum(ab[..., :, None], bb[..., None, :])
    iw = np.clip(right - left + off, np.float32(0), np.float32(6.55040e+04))
    ih = np.clip(bot - top + off, np.float32(0), np.float32(6.55040e+04))
    i = iw * ih
    area_a = ((ar - al + off) * (ab - at + off))[..., :, None]
    area_b = ((br - bl + off) * (bb - bt + off))[..., None, :]
    union = (area_a + area_b) - i
    return (i / (union + e)).astype(np.float32)


def box_iou_mxnet(lhs: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    """MXNet ``nd.contrib.box_iou(lhs, rhs, format='corner')`` -> lhs.shape[:-1] + rhs.shape[:-1].  The operator's source
    is in MXNet (un-vendored, requirements.txt:1), not in /root/reference: restated from its published algorithm
    (src/operator/contrib/bounding_box-inl.h: Intersect / BoxArea / compute_overlap) -- per axis
    w = min(right) - max(left), 0 if negative; i = w_x * w_y; 0 if i <= 0 else i / (area_l + area_r - i), area = 0 for a
    negative extent -- fp32 op for op.  PARITY UNPINNED by the reference (no MXNet here)."""
    lhs = np.asarray(lhs, np.float32)
    rhs = np.asarray(rhs, np.float32)
    L = lhs.reshape(-1, 4)[:, None, :]
    Rr = rhs.reshape(-1, 4)[None, :, :]
    z = np.float32(0)
    wx = np.minimum(L[..., 2], Rr[..., 2]) - np.maximum(L[..., 0], Rr[..., 0])
    wx = np.where(wx < z, z, wx).astype(np.float32)
    wy = np.minimum(L[..., 3], Rr[..., 3]) - np.maximum(L[..., 1], Rr[..., 1])
    wy = np.where(wy < z, z, wy).astype(np.float32)
    inter = (wx * wy).astype(np.float32)

    def area(b):
        w, h = b[..., 2] - b[..., 0], b[..., 3] - b[..., 1]
        return np.where((w < z) | (h < z), z, w * h).astype(np.float32)

    union = ((area(L) + area(Rr)).astype(np.float32) - inter).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(inter > z, inter / union, z).astype(np.float32)
    return out.reshape(lhs.shape[:-1] + rhs.shape[:-1])


def anchor_match(gt_boxes: np.ndarray, anchors: np.ndarray):
    """The anchor matching of YOLOV3PrefetchTargetGenerator.forward, models/definitions/yolo/yolo_target.py:86-94, line for
    line in numpy fp32: bbox2center = gluoncv BBoxCornerToCenter (width = xmax - xmin, height = ymax - ymin), bbox2corner
    = BBoxCenterToCorner (x -+ w/2), ``nd.contrib.box_iou`` = box_iou_mxnet, argmax = first maximum.
    gt_boxes (B, M, 4), anchors (A, 2) = all_anchors -> matches (B, M) int32, ious (B, A, M)."""
    gt = np.asarray(gt_boxes, np.float32)
    all_anchors = np.asarray(anchors, np.float32).reshape(-1, 2)
    gtw = (gt[..., 2:3] - gt[..., 0:1]).astype(np.float32)                                        # :86 bbox2center
    gth = (gt[..., 3:4] - gt[..., 1:2]).astype(np.float32)
    h = np.float32(0.5)
    shift_gt_boxes = np.concatenate([-h * gtw, -h * gth, h * gtw, h * gth], axis=-1).astype(np.float32)      # :89
    anchor_boxes = np.concatenate([np.float32(0) * all_anchors, all_anchors], axis=-1)            # :90 zero center anchors
    hw, hh = anchor_boxes[:, 2:3] / np.float32(2), anchor_boxes[:, 3:4] / np.float32(2)           # :91 bbox2corner
    shift_anchor_boxes = np.concatenate([anchor_boxes[:, 0:1] - hw, anchor_boxes[:, 1:2] - hh,
                                         anchor_boxes[:, 0:1] + hw, anchor_boxes[:, 1:2] + hh], axis=-1).astype(np.float32)
    ious = box_iou_mxnet(shift_anchor_boxes, shift_gt_boxes).transpose((1, 0, 2))                 # :92  (B, A, M)
    matches = ious.argmax(axis=1).astype(np.int32)                                                # :94  (B, M)
    return matches, ious


def conv_bn_leaky(x, w, gamma, beta, mean, var, padding, stride=1, eps=1e-5, slope=0.1, groups=1):
    """LeakyReLU_0.1(BN_eps1e-5(ConvND(x))), use_bias=False -- layers.py:63-79.

    x: (B, Cin, [T,] H, W) fp32 (NCHW / NCDHW like the reference), w: (Cout, Cin, [kt,] kh, kw).
    Arithmetic engine: torch CPU fp32 (the reference's MXNet conv is not installable).
    """
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(np.asarray(x), dtype=torch.float32)
    w = torch.as_tensor(np.asarray(w), dtype=torch.float32)
    conv = F.conv3d if x.dim() == 5 else F.conv2d
    y = conv(x, w, bias=None, stride=stride, padding=padding, groups=groups)
    shape = [1, -1] + [1] * (x.dim() - 2)
    g, b_, m, v = (torch.as_tensor(np.asarray(t), dtype=torch.float32).reshape(shape)
                   for t in (gamma, beta, mean, var))
    y = (y - m) / torch.sqrt(v + eps) * g + b_
    return F.leaky_relu(y, slope).numpy()


def conv21d_bn_leaky(x, w_s, bn_s, w_t, bn_t, padding=1, stride=1):
    """_conv21d (layers.py:82-89): (1,d,d) conv+BN+LReLU then (t,1,1) conv+BN+LReLU."""
    y = conv_bn_leaky(x, w_s, *bn_s, padding=(0, padding, padding), stride=stride)
    return conv_bn_leaky(y, w_t, *bn_t, padding=(padding, 0, 0), stride=stride)


def conv1d_bn_leaky(x, w, gamma, beta, mean, var):
    """_conv1d (layers.py:50-60) on one window: x (B, C, T, H, W), w (C, 1, T, 1, 1); depthwise
    Conv3D(kernel (T,1,1), groups=C, padding 0) + BN + LeakyReLU -> (B, C, 1, H, W)."""
    return conv_bn_leaky(x, w, gamma, beta, mean, var, padding=0, groups=np.asarray(w).shape[0])


def temporal_pool(x: np.ndarray, type: str = "max") -> np.ndarray:
    """TemporalPooling 'direct' style, layers.py:201-205: reduce axis 1 of (B,K,C,H,W)."""
    return x.max(axis=1) if type == "max" else x.mean(axis=1)


# --------------------------------------------------------------------------- consumer step
def detect_consume(ids: np.ndarray, bboxes: np.ndarray, size: float):
    """detect_yolo3.py:226,254-258 restated: clip every box to [0, size]; per image the rows with id >= 0
    and their boxes / size.  Returns (clipped (B,P,4), list of (valid_rows, normalised boxes))."""
    clipped = np.clip(bboxes, 0, size).astype(np.float32)                  # :226
    per_image = []
    for i in range(ids.shape[0]):
        valid = np.where(ids[i].flat >= 0)[0]                              # :256
        per_image.append((valid, (clipped[i, valid, :] / np.float32(size)).astype(np.float32)))   # :257
    return clipped, per_image


# --------------------------------------------------------------------------- device-side consumers
def voc_match(pred_bboxes, pred_labels, pred_scores, gt_bboxes, gt_labels, gt_difficults=None, iou_thresh=0.5):
    """What ``VOCMApMetric.update`` appends for ONE image (metrics/pascalvoc.py:116-184, class_map = None), as arrays
    instead of per-class python lists: the valid predictions ordered by (class ascending, score descending -- ties in the
    order ``argsort()[::-1]`` of a stable sort gives: later row first), each with its ``match`` value (1 true positive,
    0 false positive, -1 matched to a difficult ground truth), and the number of non-difficult ground truths per class.
    Returns (labels (n,), scores (n,), match (n,), n_pos {class: count})."""
    pred_bboxes, gt_bboxes = np.asarray(pred_bboxes, dtype=np.float32), np.asarray(gt_bboxes, dtype=np.float32)
    pred_label, pred_score = np.asarray(pred_labels).reshape(-1), np.asarray(pred_scores, dtype=np.float32).reshape(-1)
    gt_label = np.asarray(gt_labels).reshape(-1)
    valid_pred = np.where(pred_label >= 0)[0]                         # :118
    pred_bbox = pred_bboxes[valid_pred, :]
    pred_label = pred_label[valid_pred].astype(int)
    pred_score = pred_score[valid_pred]
    valid_gt = np.where(gt_label >= 0)[0]                             # :127
    gt_bbox = gt_bboxes[valid_gt, :]
    gt_label = gt_label[valid_gt].astype(int)
    gt_difficult = np.zeros(gt_bbox.shape[0]) if gt_difficults is None else np.asarray(gt_difficults).reshape(-1)[valid_gt]
    out_l, out_s, out_m, n_pos = [], [], [], {}
    for l in np.unique(np.concatenate((pred_label, gt_label)).astype(int)):                        # :136
        pred_mask_l = pred_label == l
        pred_bbox_l, pred_score_l = pred_bbox[pred_mask_l], pred_score[pred_mask_l]
        order = pred_score_l.argsort(kind="stable")[::-1]             # :141 (stable: what the default gives for <= 16 rows)
        pred_bbox_l, pred_score_l = pred_bbox_l[order], pred_score_l[order]
        gt_mask_l = gt_label == l
        gt_bbox_l, gt_difficult_l = gt_bbox[gt_mask_l], gt_difficult[gt_mask_l]
        n_pos[int(l)] = n_pos.get(int(l), 0) + int(np.logical_not(gt_difficult_l).sum())           # :149
        out_l.extend([l] * len(pred_score_l)); out_s.extend(pred_score_l)                           # :150
        if len(pred_bbox_l) == 0:
            continue
        if len(gt_bbox_l) == 0:
            out_m.extend((0,) * pred_bbox_l.shape[0])                 # :155
            continue
        iou = bbox_iou_f32(pred_bbox_l, gt_bbox_l)                    # :166 (gluoncv bbox_iou == utils/bbox.py:11-38, fp32 in)
        gt_index = iou.argmax(axis=1)
        gt_index[iou.max(axis=1) < iou_thresh] = -1                   # :169
        selec = np.zeros(gt_bbox_l.shape[0], dtype=bool)
        for gt_idx in gt_index:                                       # :173-184
            if gt_idx >= 0:
                if gt_difficult_l[gt_idx]:
                    out_m.append(-1)
                else:
                    out_m.append(0 if selec[gt_idx] else 1)
                selec[gt_idx] = True
            else:
                out_m.append(0)
    return (np.array(out_l, dtype=np.int32), np.array(out_s, dtype=np.float32), np.array(out_m, dtype=np.int32), n_pos)


def bbox_iou_f32(bbox_a, bbox_b, offset=0):
    """utils/bbox.py:11-38 evaluated in the dtype numpy gives float32 inputs (what the metric feeds it)."""
    a, b = np.asarray(bbox_a, dtype=np.float32), np.asarray(bbox_b, dtype=np.float32)
    tl = np.maximum(a[:, None, :2], b[:, :2])
    br = np.minimum(a[:, None, 2:4], b[:, 2:4])
    area_i = np.prod(br - tl + np.float32(offset), axis=2) * (tl < br).all(axis=2)
    area_a = np.prod(a[:, 2:4] - a[:, :2] + np.float32(offset), axis=1)
    area_b = np.prod(b[:, 2:4] - b[:, :2] + np.float32(offset), axis=1)
    with np.errstate(all="ignore"):
        return (area_i / (area_a[:, None] + area_b - area_i)).astype(np.float32)


def hier_iou_f32(bb, bbgt):
    """``iou`` of detect_yolo3.py:712-733 on float32 boxes (detect() hands it np.float32 scalars; the +1 / +1. are weak
    python scalars, so every operation rounds to float32)."""
    f = np.float32
    ov = f(0)
    iw = f(f(min(bb[2], bbgt[2]) - max(bb[0], bbgt[0])) + f(1))
    ih = f(f(min(bb[3], bbgt[3]) - max(bb[1], bbgt[1])) + f(1))
    if iw > 0 and ih > 0:
        intersect = f(iw * ih)
        ua = f(f(f(f(f(bb[2] - bb[0]) + f(1)) * f(f(bb[3] - bb[1]) + f(1))) +
                 f(f(f(bbgt[2] - bbgt[0]) + f(1)) * f(f(bbgt[3] - bbgt[1]) + f(1)))) - intersect)
        ov = f(intersect / ua)
    return ov


def hierarchical_nms(boxes, lifted, branch, ov_thresh=0.5, conf_thresh=0.0):
    """``hierarchical_nms`` of detect_yolo3.py:736-789 for ONE image.  ``boxes`` (n, 6) float32 rows [cls, conf, x1, y1,
    x2, y2] in the order detect() collected them; ``lifted[c]`` = the class the ``while levels[cls] > level_thresh``
    loop (:766-767) ends at for class c; ``branch[i][j]`` = ``dataset.on_branch(i, j)`` (:743-747).  Returns the new
    prediction rows (m, 6) in the order they were appended."""
    boxes = np.asarray(boxes, dtype=np.float32)
    order = sorted(range(len(boxes)), key=lambda i: boxes[i][0], reverse=True)      # :758 (stable: ties keep their order)
    new = []
    for i in order:
        cls, conf, coords = int(boxes[i][0]), boxes[i][1], boxes[i][2:6]
        if conf < conf_thresh:                                        # :763
            continue
        cls = int(lifted[cls])                                        # :766-767
        max_ov, max_idx = np.float32(0), -1
        for idx, boxb in enumerate(new):                              # :772-776
            overlap = hier_iou_f32(coords, boxb[2:6])
            if overlap > ov_thresh and overlap > max_ov:
                max_ov, max_idx = overlap, idx
        if max_idx == -1:
            new.append([cls, conf] + list(coords))                    # :779
        else:
            boxb = new[max_idx]
            if not branch[cls][int(boxb[0])]:                         # :783
                new.append([cls, conf] + list(coords))
            elif cls == int(boxb[0]):                                 # :786
                new[max_idx][1] = max(new[max_idx][1], conf)
    return np.array(new, dtype=np.float32).reshape(-1, 6)
