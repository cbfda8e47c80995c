"""Operator-level Python surface over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Names and argument meaning follow what the reference calls:
  box_nms(...)         mx.nd.contrib.box_nms as called at models/definitions/yolo/yolo3.py:526-528
  bbox_iou(a, b, off)  utils/bbox.py:11-38
  yolo3_decode(...)    YOLOOutputV3.hybrid_forward decode, yolo3.py:151-199 (+ concat :523)
  yolo3_decode_nms()   the fused inference tail, yolo3.py:496,523-534
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib

_FMT = {"corner": 0, "center": 1}
_WS_CACHE = {}


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor on a CUDA device (no CPU fallback)" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: videoyolo_b200 has no CPU path, move it to a CUDA device" % (name, t.device))
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _check_result(t: Optional[torch.Tensor], name: str, shape, dtype, device) -> torch.Tensor:
    """A caller-supplied result buffer is written through its raw pointer: it must be exactly what the kernel
    expects (a copy made here would throw the results away, so nothing is converted -- it raises)."""
    if t is None:
        return torch.empty(shape, dtype=dtype, device=device)
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != device:
        raise RuntimeError("%s must be a CUDA tensor on %s (the device of the head maps)" % (name, device))
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous (it is written in place)" % name)
    return t


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes: int, device) -> torch.Tensor:
    """Cached per-(device, stream) scratch buffer; grows monotonically."""
    key = (torch.device(device).index, _stream())
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


class workspace_scope:
    """Route ``workspace()`` to a caller-owned store for the duration of the block: an object that captures library
    calls in a CUDA graph keeps the scratch buffers its kernels write to (pipeline.GraphedModule)."""

    def __init__(self, store: dict):
        self.store = store

    def __enter__(self):
        global _WS_CACHE
        self._prev, _WS_CACHE = _WS_CACHE, self.store
        return self

    def __exit__(self, *exc):
        global _WS_CACHE
        _WS_CACHE = self._prev
        return False


def _head_args(heads: Sequence[torch.Tensor], anchors, strides):
    n = len(heads)
    if n < 1 or n > 4 or len(anchors) != n or len(strides) != n:
        raise ValueError("need 1..4 scales with matching anchors/strides")
    heads = [_need_cuda(h, "heads[%d]" % i) for i, h in enumerate(heads)]
    B = heads[0].shape[0]
    A = len(anchors[0]) // 2
    for a in anchors:
        if len(a) != 2 * A:
            raise ValueError("every scale needs the same number of anchors (%d), got %d values" % (A, len(a)))
    for h in heads:
        if h.dim() != 4 or h.shape[0] != B or h.device != heads[0].device:
            raise ValueError("heads must be (B, A*(5+C), H, W) on one device")
    ptrs = (ctypes.c_void_p * n)(*[h.data_ptr() for h in heads])
    H = (ctypes.c_int * n)(*[h.shape[2] for h in heads])
    W = (ctypes.c_int * n)(*[h.shape[3] for h in heads])
    st = (ctypes.c_float * n)(*[float(s) for s in strides])
    flat = [float(v) for a in anchors for v in a]
    an = (ctypes.c_float * len(flat))(*flat)
    return heads, ptrs, H, W, st, an, n, B, A


def n_rows(heads: Sequence[torch.Tensor], num_class: int, A: int = 3, agnostic: bool = False) -> int:
    ceff = 1 if agnostic else num_class
    return ceff * sum(h.shape[2] * h.shape[3] * A for h in heads)


def yolo3_decode(heads, num_class: int, anchors, strides, agnostic: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """heads (network order: stride 32,16,8) -> (B, R, 6) detections in the reference's row order."""
    heads, ptrs, H, W, st, an, n, B, A = _head_args(heads, anchors, strides)
    for h in heads:
        if h.shape[1] != A * (5 + num_class):
            raise ValueError("head has %d channels, expected A*(5+C)=%d" % (h.shape[1], A * (5 + num_class)))
    R = n_rows(heads, num_class, A, agnostic)
    out = _check_result(out, "out", (B, R, 6), torch.float32, heads[0].device)
    if B == 0:
        return out
    with torch.cuda.device(heads[0].device):
        _lib.check(_lib.lib().vy_decode_f32(ptrs, H, W, st, an, n, B, A, num_class, int(agnostic),
                                            out.data_ptr(), _stream()))
    return out


def box_nms(data: torch.Tensor, overlap_thresh: float = 0.5, valid_thresh: float = 0, topk: int = -1,
            coord_start: int = 2, score_index: int = 1, id_index: int = -1, background_id: int = -1,
            force_suppress: bool = False, in_format: str = "corner", out_format: str = "corner",
            out_rows: Optional[int] = None, return_kept: bool = False, _exhaustive: bool = False):
    """MXNet ``contrib.box_nms`` on a CUDA tensor; all leading dims are batch.

    Returns a tensor of the same shape as ``data`` (or with ``out_rows`` rows when given, which fuses
    the ``slice_axis(axis=1, 0, post_nms)`` of yolo3.py:529-530); with ``return_kept`` also the int32
    source-row index of every output row (MXNet's hidden ``record`` output).
    """
    data = _need_cuda(data, "data")
    if data.dim() < 2:
        raise ValueError("data must be (..., R, W)")
    shp = tuple(data.shape)
    R, Wd = shp[-2], shp[-1]
    B = 1
    for d in shp[:-2]:
        B *= d
    rows = R if out_rows is None else int(out_rows)
    out = torch.empty(shp[:-2] + (rows, Wd), dtype=torch.float32, device=data.device)
    kept = torch.empty(shp[:-2] + (rows,), dtype=torch.int32, device=data.device)
    if B == 0 or R == 0:
        return (out, kept) if return_kept else out
    L = _lib.lib()
    with torch.cuda.device(data.device):
        need = L.vy_box_nms_workspace_bytes(B, R, Wd, int(topk))
        ws = workspace(need, data.device)
        _lib.check(L.vy_box_nms_f32(data.data_ptr(), B, R, Wd, float(overlap_thresh), float(valid_thresh),
                                    int(topk), int(coord_start), int(score_index), int(id_index),
                                    int(background_id), int(bool(force_suppress)) | (0x100 if _exhaustive else 0), _FMT[in_format],
                                    _FMT[out_format], rows, out.data_ptr(), kept.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream()))
    return (out, kept) if return_kept else out


class DecodeNmsPlan:
    """Owns a ``vy_decode_nms_plan_t`` (include/vyolo.h): the fused call's shape-dependent planning done once, like the
    cached graph of a hybridized Gluon block (detect_yolo3.py:204).  ``launch`` only fills in this call's pointers."""

    def __init__(self, shapes, num_class, anchors, strides, nms_thresh, valid_thresh, topk, post_nms, force_suppress,
                 agnostic):
        n = len(shapes)
        if n < 1 or n > 4 or len(anchors) != n or len(strides) != n:
            raise ValueError("need 1..4 scales with matching anchors/strides")
        A = len(anchors[0]) // 2
        for a in anchors:
            if len(a) != 2 * A:
                raise ValueError("every scale needs the same number of anchors (%d), got %d values" % (A, len(a)))
        B = shapes[0][0]
        for sh in shapes:
            if len(sh) != 4 or sh[0] != B:
                raise ValueError("heads must be (B, A*(5+C), H, W)")
            if sh[1] != A * (5 + num_class):
                raise ValueError("head has %d channels, expected A*(5+C)=%d" % (sh[1], A * (5 + num_class)))
        self.B, self.A, self.n, self.post_nms = B, A, n, int(post_nms)
        H = (ctypes.c_int * n)(*[sh[2] for sh in shapes])
        W = (ctypes.c_int * n)(*[sh[3] for sh in shapes])
        st = (ctypes.c_float * n)(*[float(v) for v in strides])
        flat = [float(v) for a in anchors for v in a]
        an = (ctypes.c_float * len(flat))(*flat)
        L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(L.vy_decode_nms_plan_create(H, W, st, an, n, B, A, num_class, int(agnostic), float(nms_thresh),
                                               float(valid_thresh), int(topk), int(bool(force_suppress)), int(post_nms),
                                               ctypes.byref(h)))
        self._h, self._destroy = h, L.vy_decode_nms_plan_destroy
        self._launch = L.vy_decode_nms_plan_launch
        self.workspace_bytes = int(L.vy_decode_nms_plan_workspace_bytes(h))
        self._ptrs = (ctypes.c_void_p * n)()

    def launch(self, heads, out, kept, ws, stream: int):
        for i, t in enumerate(heads):
            self._ptrs[i] = t.data_ptr()
        _lib.check(self._launch(self._h, self._ptrs, out.data_ptr(), kept.data_ptr(), ws.data_ptr(), ws.numel(), stream))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._destroy(h)


_PLANS = {}


def decode_nms_plan(shapes, num_class, anchors, strides, nms_thresh=0.45, valid_thresh=0.01, topk=400, post_nms=100,
                    force_suppress=False, agnostic=False) -> DecodeNmsPlan:
    """The (cached) plan of a fused call on head maps of these shapes with these arguments."""
    key = (tuple(tuple(int(v) for v in sh) for sh in shapes), int(num_class), tuple(tuple(a) for a in anchors),
           tuple(strides), float(nms_thresh), float(valid_thresh), int(topk), int(post_nms), bool(force_suppress),
           bool(agnostic))
    p = _PLANS.get(key)
    if p is None:
        if len(_PLANS) > 256:
            _PLANS.clear()
        p = _PLANS[key] = DecodeNmsPlan(key[0], num_class, anchors, strides, nms_thresh, valid_thresh, topk, post_nms,
                                        force_suppress, agnostic)
    return p


def yolo3_decode_nms(heads, num_class: int, anchors, strides, nms_thresh: float = 0.45,
                     valid_thresh: float = 0.01, topk: int = 400, post_nms: int = 100,
                     force_suppress: bool = False, agnostic: bool = False,
                     out: Optional[torch.Tensor] = None, kept: Optional[torch.Tensor] = None,
                     workspace_buf: Optional[torch.Tensor] = None):
    """Fused decode + box_nms + post_nms slice.  Returns (out (B, post_nms, 6), kept (B, post_nms)).
    ``workspace_buf``: a caller-owned uint8 CUDA scratch tensor (a captured CUDA graph must own the buffer its
    kernels write to); default: the per-(device, stream) cached one."""
    if len(heads) < 1:
        raise ValueError("need 1..4 scales with matching anchors/strides")
    heads = [_need_cuda(h, "heads[%d]" % i) for i, h in enumerate(heads)]
    dev = heads[0].device
    for h in heads:
        if h.dim() != 4 or h.device != dev:
            raise ValueError("heads must be (B, A*(5+C), H, W) on one device")
    B = heads[0].shape[0]
    out = _check_result(out, "out", (B, post_nms, 6), torch.float32, dev)
    kept = _check_result(kept, "kept", (B, post_nms), torch.int32, dev)
    if B == 0:                                           # an empty batch is an empty result (MXNet operators agree)
        return out, kept
    plan = decode_nms_plan([h.shape for h in heads], num_class, anchors, strides, nms_thresh, valid_thresh, topk,
                           post_nms, force_suppress, agnostic)
    with torch.cuda.device(dev):
        if workspace_buf is None:
            ws = workspace(plan.workspace_bytes, dev)
        else:
            ws = workspace_buf
            if not ws.is_cuda or ws.device != dev or ws.dtype != torch.uint8 or ws.numel() < plan.workspace_bytes:
                raise ValueError("workspace_buf must be a uint8 CUDA tensor of >= %d bytes on %s" % (plan.workspace_bytes, dev))
        plan.launch(heads, out, kept, ws, _stream())
    return out, kept


def bbox_iou(bbox_a: torch.Tensor, bbox_b: torch.Tensor, offset=0) -> torch.Tensor:
    """utils/bbox.py:11-38 on CUDA tensors (float32 or float64): (N,>=4) x (M,>=4) -> (N, M)."""
    if not (isinstance(bbox_a, torch.Tensor) and isinstance(bbox_b, torch.Tensor)):
        raise TypeError("bbox_iou needs torch CUDA tensors (no CPU fallback)")
    if bbox_a.dim() != 2 or bbox_b.dim() != 2 or bbox_a.shape[1] < 4 or bbox_b.shape[1] < 4:
        raise IndexError("Bounding boxes axis 1 must have at least length 4")      # utils/bbox.py:29-30
    dt = torch.float64 if (bbox_a.dtype == torch.float64 or bbox_b.dtype == torch.float64) else torch.float32
    a = _need_cuda(bbox_a.to(dt), "bbox_a", dt)
    b = _need_cuda(bbox_b.to(dt), "bbox_b", dt)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=dt, device=a.device)
    fn = _lib.lib().vy_bbox_iou_f64 if dt == torch.float64 else _lib.lib().vy_bbox_iou_f32
    with torch.cuda.device(a.device):
        _lib.check(fn(a.data_ptr(), a.shape[0], a.shape[1], b.data_ptr(), b.shape[0], b.shape[1],
                      float(offset), out.data_ptr(), _stream()))
    return out


def bbox_batch_iou(a: torch.Tensor, b: torch.Tensor, offset=0, eps=1e-15, ignore_iou_thresh: Optional[float] = None,
                   return_ious: bool = True):
    """gluoncv ``BBoxBatchIOU()(a, b)`` as the dynamic-target step calls it (yolo_target.py:171,202): corner boxes
    a (B, N, 4), b (B, M, 4) -> ious (B, N, M).  With ``ignore_iou_thresh`` the two lines that follow in the
    reference are fused into the same pass and returned as well: ``ious_max`` (B, N, 1) (:203) and
    ``objness_t = (ious_max > thresh) * -1`` (B, N, 1) (:204); ``return_ious=False`` then skips the (B, N, M) write.
    Returns ious | (ious_or_None, ious_max, objness_t)."""
    a = _need_cuda(a, "a")
    b = _need_cuda(b, "b")
    if a.dim() != 3 or b.dim() != 3 or a.shape[2] != 4 or b.shape[2] != 4 or a.shape[0] != b.shape[0]:
        raise ValueError("bbox_batch_iou: a (B, N, 4) and b (B, M, 4) expected")
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    if M < 1:
        raise ValueError("bbox_batch_iou: b needs at least one box per image")
    fused = ignore_iou_thresh is not None
    if not fused and not return_ious:
        raise ValueError("bbox_batch_iou: nothing to return")
    ious = torch.empty((B, N, M), dtype=torch.float32, device=a.device) if return_ious else None
    imax = torch.empty((B, N, 1), dtype=torch.float32, device=a.device) if fused else None
    obj = torch.empty((B, N, 1), dtype=torch.float32, device=a.device) if fused else None
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().vy_bbox_batch_iou_f32(
            a.data_ptr(), b.data_ptr(), B, N, M, float(offset), float(eps), float(ignore_iou_thresh if fused else 0.0),
            ious.data_ptr() if ious is not None else None, imax.data_ptr() if fused else None,
            obj.data_ptr() if fused else None, _stream()))
    return (ious, imax, obj) if fused else ious


def anchor_match(gt_boxes: torch.Tensor, anchors, return_ious: bool = False):
    """The anchor matching of YOLOV3PrefetchTargetGenerator (yolo_target.py:86-94): ``nd.contrib.box_iou`` of the
    zero-centred anchors against the zero-centred ground-truth boxes and ``argmax`` over the anchors, one kernel.
    gt_boxes (B, M, 4) corner boxes (fp32 CUDA), anchors: the (A, 2) ``all_anchors`` as (w, h).
    Returns matches (B, M) int32 [, ious (B, A, M)]."""
    gt = _need_cuda(gt_boxes, "gt_boxes")
    if gt.dim() != 3 or gt.shape[2] != 4:
        raise ValueError("anchor_match: gt_boxes (B, M, 4) expected")
    an = torch.as_tensor(anchors, dtype=torch.float32).reshape(-1, 2).to(gt.device).contiguous()
    B, M, A = gt.shape[0], gt.shape[1], an.shape[0]
    matches = torch.empty((B, M), dtype=torch.int32, device=gt.device)
    ious = torch.empty((B, A, M), dtype=torch.float32, device=gt.device) if return_ious else None
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().vy_anchor_match_f32(gt.data_ptr(), B, M, an.data_ptr(), A, matches.data_ptr(),
                                                  ious.data_ptr() if return_ious else None, _stream()))
    return (matches, ious) if return_ious else matches


def detect_consume(dets: torch.Tensor, size: float):
    """The device part of what detect()/validate() do with net(x)'s output (detect_yolo3.py:226,254-258):
    returns (bboxes clipped to [0, size] (B,P,4), normalised boxes of the valid rows (B,P,4; -1 elsewhere),
    number of valid rows per image (B,) int32).  ``dets`` is the (B, P, 6) tensor of yolo3_decode_nms."""
    dets = _need_cuda(dets, "dets")
    if dets.dim() != 3 or dets.shape[2] != 6:
        raise ValueError("dets must be (B, P, 6)")
    B, P = dets.shape[0], dets.shape[1]
    clipped = torch.empty((B, P, 4), dtype=torch.float32, device=dets.device)
    normed = torch.empty((B, P, 4), dtype=torch.float32, device=dets.device)
    counts = torch.empty((B,), dtype=torch.int32, device=dets.device)
    with torch.cuda.device(dets.device):
        _lib.check(_lib.lib().vy_detect_consume_f32(dets.data_ptr(), B, P, float(size), float(size), clipped.data_ptr(),
                                                    normed.data_ptr(), counts.data_ptr(), _stream()))
    return clipped, normed, counts


def hierarchical_nms(boxes: torch.Tensor, lifted, branch, ov_thresh: float = 0.5, conf_thresh: float = 0.0):
    """``hierarchical_nms`` of detect_yolo3.py:736-789 on the device.  ``boxes`` (B, N, 6) float32 CUDA rows
    [cls, conf, x1, y1, x2, y2] (cls < 0 = padding); ``lifted`` (n_cls,) the class every class is raised to by the
    ``level_thresh`` walk (:766-767); ``branch`` (n_cls, n_cls) ``dataset.on_branch(i, j)``.  Returns
    (new rows (B, N, 6) in the reference's append order, -1 padded; rows kept per image (B,) int32)."""
    boxes = _need_cuda(boxes, "boxes")
    if boxes.dim() != 3 or boxes.shape[2] != 6:
        raise ValueError("boxes must be (B, N, 6)")
    dev = boxes.device
    lifted = torch.as_tensor(lifted, dtype=torch.int32, device=dev).contiguous()
    branch = torch.as_tensor(branch, device=dev).to(torch.uint8).contiguous()
    n_cls = lifted.numel()
    if tuple(branch.shape) != (n_cls, n_cls):
        raise ValueError("branch must be (n_cls, n_cls)")
    B, N = boxes.shape[0], boxes.shape[1]
    out = torch.empty_like(boxes)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    if B == 0 or N == 0:
        return out, counts
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vy_hier_nms_f32(boxes.data_ptr(), B, N, lifted.data_ptr(), branch.data_ptr(), n_cls,
                                              float(ov_thresh), float(conf_thresh), out.data_ptr(), counts.data_ptr(), _stream()))
    return out, counts


def voc_match(dets: torch.Tensor, gt_bboxes: torch.Tensor, gt_labels: torch.Tensor, gt_difficults: Optional[torch.Tensor] = None,
              n_class: int = 1, iou_thresh: float = 0.5):
    """The per-image part of ``VOCMApMetric.update`` (metrics/pascalvoc.py:116-184) on the device.  ``dets`` (B, P, 6) as
    returned by the fused tail; ``gt_bboxes`` (B, M, 4), ``gt_labels`` (B, M[, 1]) (< 0 = padding), ``gt_difficults``
    (B, M[, 1]) or None.  Returns (labels (B, P) int32, scores (B, P), match (B, P) int32 in {1, 0, -1}; -2 padding;
    valid predictions per image (B,), n_pos (B, n_class)): per image the valid predictions in the metric's order
    (class ascending, score descending, later row first among equal scores)."""
    dets = _need_cuda(dets, "dets")
    gb = _need_cuda(gt_bboxes, "gt_bboxes")
    gl = _need_cuda(gt_labels.reshape(gt_labels.shape[0], -1).float(), "gt_labels")
    gd = _need_cuda(gt_difficults.reshape(gt_difficults.shape[0], -1).float(), "gt_difficults") if gt_difficults is not None else None
    if dets.dim() != 3 or dets.shape[2] != 6 or gb.dim() != 3 or gb.shape[2] != 4 or gb.shape[0] != dets.shape[0]:
        raise ValueError("dets (B, P, 6) and gt_bboxes (B, M, 4) expected")
    B, P, M = dets.shape[0], dets.shape[1], gb.shape[1]
    if tuple(gl.shape) != (B, M) or (gd is not None and tuple(gd.shape) != (B, M)):
        raise ValueError("gt_labels / gt_difficults must be (B, M)")
    dev = dets.device
    lab = torch.empty((B, P), dtype=torch.int32, device=dev)
    sc = torch.empty((B, P), dtype=torch.float32, device=dev)
    mt = torch.empty((B, P), dtype=torch.int32, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    npos = torch.empty((B, n_class), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vy_voc_match_f32(dets.data_ptr(), gb.data_ptr(), gl.data_ptr(), gd.data_ptr() if gd is not None else None,
                                               B, P, M, int(n_class), float(iou_thresh), lab.data_ptr(), sc.data_ptr(),
                                               mt.data_ptr(), cnt.data_ptr(), npos.data_ptr(), _stream()))
    return lab, sc, mt, cnt, npos


# ----------------------------------------------------------------------------------- temporal fusion conv
class PTensor:
    """An activation in the library's P layout: ``data`` is a bf16 (or fp32) CUDA tensor of shape
    (T, B, H+2, W+2, C) whose one-pixel spatial border is zero (include/vyolo.h)."""

    __slots__ = ("data", "B", "T", "H", "W", "C")

    def __init__(self, data, B, T, H, W, C):
        self.data, self.B, self.T, self.H, self.W, self.C = data, B, T, H, W, C

    @property
    def shape(self):                       # the reference's NCDHW view of the same tensor
        return (self.B, self.C, self.T, self.H, self.W)


def pack_p(x: torch.Tensor, layout: str = "NCDHW") -> PTensor:
    """fp32 CUDA tensor in a reference layout -> P layout bf16.  ``layout``: 'NCDHW' (B,C,T,H,W), the
    block-internal layout after ``swapaxes(1, 2)`` (yolo3.py:256); 'NTCHW' (B,K,C,H,W), what the
    temporal models carry between blocks; 'NCHW' (B,C,H,W) for 2-D cells."""
    x = _need_cuda(x, "x")
    if layout == "NCHW":
        B, C, H, W = x.shape
        T, sb, sc, st = 1, C * H * W, H * W, 0
    elif layout == "NCDHW":
        B, C, T, H, W = x.shape
        sb, sc, st = C * T * H * W, T * H * W, H * W
    elif layout == "NTCHW":
        B, T, C, H, W = x.shape
        sb, st, sc = T * C * H * W, C * H * W, H * W
    else:
        raise ValueError("layout must be NCHW, NCDHW or NTCHW")
    out = torch.empty((T, B, H + 2, W + 2, C), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vy_pack_f32_to_p_bf16(x.data_ptr(), sb, sc, st, B, C, T, H, W, out.data_ptr(), _stream()))
    return PTensor(out, B, T, H, W, C)


def pack_p_split(x: torch.Tensor, layout: str = "NCHW") -> PTensor:
    """fp32 CUDA tensor -> P layout bf16 with 3*Cpad channels [hi | hi | lo] (hi = bf16(v), lo = bf16(v - hi), Cpad = C
    rounded up to 64, padding zero): the activation operand of a split-precision 1x1 conv (``split_weight(w, 3)``)."""
    x = _need_cuda(x, "x")
    if layout == "NCHW":
        B, C, H, W = x.shape
        T, sb, sc, st = 1, C * H * W, H * W, 0
    elif layout == "NTCHW":
        B, T, C, H, W = x.shape
        sb, st, sc = T * C * H * W, C * H * W, H * W
    else:
        raise ValueError("layout must be NCHW or NTCHW")
    Cpad = (C + 63) // 64 * 64
    out = torch.empty((T, B, H + 2, W + 2, 3 * Cpad), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vy_pack_f32_split_to_p_bf16(x.data_ptr(), sb, sc, st, B, C, Cpad, T, H, W, out.data_ptr(), _stream()))
    return PTensor(out, B, T, H, W, 3 * Cpad)


def cat_repeat(x: PTensor, rep: int = 1) -> PTensor:
    """The 'cat' late join (reshape (0,-3,-2), yolo3.py:1135-1136) of a P-layout activation: (T, B, ., ., C) ->
    (1, B, ., ., rep*T*C), channel r*T*C + t*C + c = frame t, channel c; ``rep`` repeats the joined channels
    (rep = 2 pairs with ``split_weight(w, 2)``)."""
    if not isinstance(x, PTensor) or x.data.dtype != torch.bfloat16:
        raise TypeError("cat_repeat takes a bf16 PTensor")
    y = torch.empty((1, x.B, x.H + 2, x.W + 2, rep * x.T * x.C), dtype=torch.bfloat16, device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_cat_repeat_bf16(x.data.data_ptr(), x.B, x.T, x.H, x.W, x.C, int(rep), y.data_ptr(), _stream()))
    return PTensor(y, x.B, 1, x.H, x.W, rep * x.T * x.C)


def split_weight(w: torch.Tensor, parts: int) -> torch.Tensor:
    """1x1 conv weight (Cout, Cin[, 1, 1]) fp32 -> the fusion-conv operand (Cout_pad, 1, 1, 1, parts*Cin_pad) bf16 that
    carries it at fp32 grade: [w_hi | w_lo] against an activation repeated twice (``cat_repeat(x, 2)``), or
    [w_hi | w_lo | w_hi] against ``pack_p_split``'s [hi | hi | lo].  Cout / Cin are zero-padded to multiples of 64."""
    w = w.reshape(w.shape[0], -1).float()
    n, cin = w.shape
    npad, cpad = (n + 63) // 64 * 64, (cin + 63) // 64 * 64
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    out = torch.zeros((npad, 1, 1, 1, parts * cpad), dtype=torch.bfloat16, device=w.device)
    for i, part in enumerate([hi, lo, hi][:parts]):
        out[:n, 0, 0, 0, i * cpad: i * cpad + cin] = part
    return out


def unpack_p(p: PTensor, layout: str = "NCDHW", channels: Optional[int] = None) -> torch.Tensor:
    """P layout (bf16 or fp32) -> fp32 CUDA tensor in a reference layout (see pack_p).  ``channels``: only the
    first ``channels`` of the tensor's channels (a conv whose Cout was padded to a multiple of 64)."""
    B, T, H, W, Cp = p.B, p.T, p.H, p.W, p.C
    C = Cp if channels is None else int(channels)
    if not 1 <= C <= Cp:
        raise ValueError("channels must be in [1, %d]" % Cp)
    if layout == "NCHW":
        if T != 1:
            raise ValueError("NCHW needs T == 1")
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=p.data.device)
        sb, sc, st = C * H * W, H * W, 0
    elif layout == "NCDHW":
        out = torch.empty((B, C, T, H, W), dtype=torch.float32, device=p.data.device)
        sb, sc, st = C * T * H * W, T * H * W, H * W
    elif layout == "NTCHW":
        out = torch.empty((B, T, C, H, W), dtype=torch.float32, device=p.data.device)
        sb, st, sc = T * C * H * W, C * H * W, H * W
    else:
        raise ValueError("layout must be NCHW, NCDHW or NTCHW")
    with torch.cuda.device(p.data.device):
        _lib.check(_lib.lib().vy_unpack_p_channels_to_f32(p.data.data_ptr(), int(p.data.dtype == torch.float32), B, Cp, C,
                                                          T, H, W, out.data_ptr(), sb, sc, st, _stream()))
    return out


def fusion_conv(x: PTensor, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                slope: float = 0.1, out_f32: bool = False, pool_max: bool = False) -> PTensor:
    """One Conv+BN+LeakyReLU cell (layers.py:63-79) on a P-layout activation.

    weight: (Cout, kt, kh, kw, Cin) bf16 CUDA (see ``conv_weight``); scale/shift: folded inference
    BatchNorm per output channel, fp32 CUDA.  'same' padding, stride 1.
    ``pool_max``: the cell followed by ``TemporalPooling(k, 'max')`` (the late join, yolo3.py:1134-1138) in one call:
    returns the pooled frame (T = 1); the un-pooled output is never written."""
    if not isinstance(x, PTensor):
        raise TypeError("fusion_conv takes a PTensor (ops.pack_p)")
    w = _need_cuda(weight, "weight", torch.bfloat16)
    scale = _need_cuda(scale, "scale")
    shift = _need_cuda(shift, "shift")
    Cout, kt, kh, kw, Cin = w.shape
    if Cin != x.C:
        raise ValueError("weight has %d input channels, activation has %d" % (Cin, x.C))
    if scale.numel() != Cout or shift.numel() != Cout:
        raise ValueError("scale/shift must have Cout elements")
    if pool_max:
        if out_f32:
            raise ValueError("pool_max writes bf16")
        y = torch.empty((1, x.B, x.H + 2, x.W + 2, Cout), dtype=torch.bfloat16, device=x.data.device)
        with torch.cuda.device(x.data.device):
            _lib.check(_lib.lib().vy_fusion_conv_bf16_maxpool(x.data.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                              float(slope), x.B, x.T, x.H, x.W, Cin, Cout, kt, kh, kw,
                                                              y.data_ptr(), _stream()))
        return PTensor(y, x.B, 1, x.H, x.W, Cout)
    y = torch.empty((x.T, x.B, x.H + 2, x.W + 2, Cout), dtype=torch.float32 if out_f32 else torch.bfloat16,
                    device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_fusion_conv_bf16(x.data.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                  float(slope), x.B, x.T, x.H, x.W, Cin, Cout, kt, kh, kw,
                                                  y.data_ptr(), int(out_f32), None, 0, _stream()))
    return PTensor(y, x.B, x.T, x.H, x.W, Cout)


def fusion_conv_nchw(x: PTensor, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                     slope: float = 1.0, channels: Optional[int] = None) -> torch.Tensor:
    """The same cell for a 2-D input (T == 1, (1, kh, kw) weight) with the result in the reference's own layout: the fp32
    (B, channels, H, W) tensor, first ``channels`` of the Cout output channels.  The 1x1 ``prediction`` conv uses it
    (slope 1 = no activation, bias as shift): its head maps go to the decode without a layout pass."""
    if not isinstance(x, PTensor):
        raise TypeError("fusion_conv_nchw takes a PTensor (ops.pack_p)")
    if x.T != 1:
        raise ValueError("fusion_conv_nchw needs T == 1")
    w = _need_cuda(weight, "weight", torch.bfloat16)
    scale = _need_cuda(scale, "scale")
    shift = _need_cuda(shift, "shift")
    Cout, kt, kh, kw, Cin = w.shape
    if kt != 1:
        raise ValueError("fusion_conv_nchw takes a (Cout, 1, kh, kw, Cin) weight")
    if Cin != x.C:
        raise ValueError("weight has %d input channels, activation has %d" % (Cin, x.C))
    if scale.numel() != Cout or shift.numel() != Cout:
        raise ValueError("scale/shift must have Cout elements")
    C = Cout if channels is None else int(channels)
    if not 1 <= C <= Cout:
        raise ValueError("channels must be in [1, %d]" % Cout)
    y = torch.empty((x.B, C, x.H, x.W), dtype=torch.float32, device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_fusion_conv_bf16_nchw(x.data.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                       float(slope), x.B, x.H, x.W, Cin, Cout, kh, kw,
                                                       y.data_ptr(), C, _stream()))
    return y


def fusion_conv_nchw_joined(x: PTensor, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                            rep: int = 1, slope: float = 1.0, channels: Optional[int] = None) -> torch.Tensor:
    """``fusion_conv_nchw(cat_repeat(x, rep), ...)`` without materialising the joined / repeated activation: the 1x1 conv
    reads frame t, channel block c of ``x`` for k-block (r, t, c) of the (Cout, 1, 1, 1, rep*T*C) weight."""
    if not isinstance(x, PTensor) or x.data.dtype != torch.bfloat16:
        raise TypeError("fusion_conv_nchw_joined takes a bf16 PTensor")
    w = _need_cuda(weight, "weight", torch.bfloat16)
    scale = _need_cuda(scale, "scale")
    shift = _need_cuda(shift, "shift")
    Cout, kt, kh, kw, Cin = w.shape
    if (kt, kh, kw) != (1, 1, 1) or Cin != rep * x.T * x.C:
        raise ValueError("weight must be (Cout, 1, 1, 1, rep*T*C = %d), got %s" % (rep * x.T * x.C, tuple(w.shape)))
    if scale.numel() != Cout or shift.numel() != Cout:
        raise ValueError("scale/shift must have Cout elements")
    C = Cout if channels is None else int(channels)
    if not 1 <= C <= Cout:
        raise ValueError("channels must be in [1, %d]" % Cout)
    y = torch.empty((x.B, C, x.H, x.W), dtype=torch.float32, device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_fusion_conv_bf16_nchw_joined(x.data.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                              float(slope), x.B, x.T, x.H, x.W, x.C, int(rep), Cout,
                                                              y.data_ptr(), C, _stream()))
    return y


def conv_weight(w_ref: torch.Tensor) -> torch.Tensor:
    """The reference's conv weight (Cout, Cin, [kt,] kh, kw) fp32 -> (Cout, kt, kh, kw, Cin) bf16 CUDA."""
    if w_ref.dim() == 4:
        w_ref = w_ref[:, :, None]
    return w_ref.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def fold_bn(gamma, beta, mean, var, eps: float = 1e-5):
    """Inference BatchNorm (layers.py:68,77: epsilon=1e-5) as y = x*scale + shift."""
    scale = gamma / torch.sqrt(var + eps)
    return scale.float().contiguous(), (beta - mean * scale).float().contiguous()


def temporal_pool(x: PTensor, type: str = "max") -> PTensor:
    """TemporalPooling 'direct' style (layers.py:201-205): reduce the K frames to one."""
    if type not in ("max", "mean"):
        raise ValueError("type must be max or mean")
    if x.data.dtype != torch.bfloat16:
        raise TypeError("temporal_pool takes a bf16 P-layout activation")
    inner = x.B * (x.H + 2) * (x.W + 2) * x.C
    y = torch.empty((1, x.B, x.H + 2, x.W + 2, x.C), dtype=torch.bfloat16, device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_temporal_pool_bf16(x.data.data_ptr(), x.T, inner, 0 if type == "max" else 1,
                                                    y.data_ptr(), _stream()))
    return PTensor(y, x.B, 1, x.H, x.W, x.C)


def upsample_concat(up: PTensor, route: PTensor) -> PTensor:
    """``concat(slice_like(_upsample(up, 2), route), route)`` along channels (layers.py:11-20, yolo3.py:1170-1177):
    the join between two scales of the YOLO neck, P layout in and out."""
    if up.data.dtype != torch.bfloat16 or route.data.dtype != torch.bfloat16:
        raise TypeError("upsample_concat takes bf16 P-layout activations")
    if up.B != route.B or up.T != route.T:
        raise ValueError("upsample_concat: batch / frames differ")
    if 2 * up.H < route.H or 2 * up.W < route.W:
        raise ValueError("upsample_concat: the upsampled map must cover the route")
    y = torch.empty((route.T, route.B, route.H + 2, route.W + 2, up.C + route.C), dtype=torch.bfloat16, device=route.data.device)
    with torch.cuda.device(route.data.device):
        _lib.check(_lib.lib().vy_upsample_concat_bf16(up.data.data_ptr(), route.data.data_ptr(), route.B, route.T,
                                                      route.H, route.W, up.H, up.W, up.C, route.C, y.data_ptr(), _stream()))
    return PTensor(y, route.B, route.T, route.H, route.W, up.C + route.C)


def temporal_dwconv(x: PTensor, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                    slope: float = 0.1) -> PTensor:
    """``_conv1d`` (layers.py:50-60) over a window of exactly ``x.T`` frames: depthwise Conv3D with kernel
    (T,1,1) + BN + LeakyReLU, the temporal merge of HDarknet (h_darknet.py:97-119).
    weight: the reference's (C, 1, T, 1, 1) (or (C, T)) fp32 CUDA tensor."""
    if not isinstance(x, PTensor) or x.data.dtype != torch.bfloat16:
        raise TypeError("temporal_dwconv takes a bf16 PTensor (ops.pack_p)")
    w = _need_cuda(weight.reshape(weight.shape[0], -1), "weight")
    if tuple(w.shape) != (x.C, x.T):
        raise ValueError("weight must be (C, 1, T, 1, 1) with C=%d, T=%d" % (x.C, x.T))
    scale = _need_cuda(scale, "scale")
    shift = _need_cuda(shift, "shift")
    y = torch.empty((1, x.B, x.H + 2, x.W + 2, x.C), dtype=torch.bfloat16, device=x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(_lib.lib().vy_temporal_dwconv_bf16(x.data.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                      float(slope), x.B, x.T, x.H, x.W, x.C, y.data_ptr(), _stream()))
    return PTensor(y, x.B, 1, x.H, x.W, x.C)
