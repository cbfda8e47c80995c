"""Host-side mirror of the reference's Gluon block surface for the detection post-processing path.

Same names, constructor arguments, return shapes and error behaviour as
models/definitions/yolo/yolo3.py (YOLOOutputV3 :25-199, YOLOV3_noback :1686-1870, set_nms :536-556);
the arithmetic runs in the CUDA library behind include/vyolo.h.  Backbone stages, training branches
and reset_class are out of scope.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops

# models/definitions/yolo/wrappers.py:80-84
ANCHORS = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
STRIDES = [8, 16, 32]
CHANNELS = [512, 256, 128]


class Prediction(torch.nn.Module):
    """The 1x1 ``prediction`` conv of YOLOOutputV3 (``nn.Conv2D(all_pred, kernel_size=1, padding=0, strides=1)``,
    yolo3.py:62, applied at :157): weight (all_pred, in_channels, 1, 1) + bias, MXNet's default initialisation
    (Uniform(0.07) weight, zero bias).  It runs in the library's tcgen05 fusion-conv kernel -- no cuDNN / cuBLAS -- with
    SPLIT operands so that it keeps the reference's fp32 grade: a value is carried as hi = bf16(v) and lo = bf16(v - hi)
    in separate channels and the products w_hi*v_hi + w_lo*v_hi + w_hi*v_lo are accumulated in fp32 in TMEM (what is
    dropped is ~2^-16 relative).  Inputs: an fp32 NCHW tensor (the tip of a 2-D model), or a bf16 ``ops.PTensor`` (the
    tip of the temporal models: exact in bf16 already, so only the weight is split; K frames are joined 'cat'-wise,
    yolo3.py:1135-1136).  Returns the fp32 NCHW head map (B, all_pred, H, W)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty((out_channels, in_channels, 1, 1)).uniform_(-0.07, 0.07))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self._packed = {}

    def _operands(self, parts: int):
        key = (self.weight._version, self.bias._version, self.weight.device)
        hit = self._packed.get(parts)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                w = ops.split_weight(self.weight.detach(), parts)
                npad = w.shape[0]
                shift = torch.zeros(npad, dtype=torch.float32, device=w.device)
                shift[: self.bias.numel()] = self.bias.detach().float()
                hit = (key, w, torch.ones(npad, dtype=torch.float32, device=w.device), shift)
            self._packed[parts] = hit
        return hit[1:]

    def forward(self, x):
        n, cin = self.weight.shape[0], self.weight.shape[1]
        if isinstance(x, ops.PTensor):
            if x.T * x.C != cin:
                raise ValueError("prediction conv expects %d input channels, got %d x %d frames" % (cin, x.C, x.T))
            w, scale, shift = self._operands(2)
            if x.C % 64 == 0:
                # the join ('cat': K frames side by side in the channels) and the repeat for the split weight happen in
                # the conv's operand addressing: nothing is materialised
                return ops.fusion_conv_nchw_joined(x, w, scale, shift, rep=2, slope=1.0, channels=n)
            xp = ops.cat_repeat(x, 2)
        else:
            if x.dim() != 4 or x.shape[1] != cin:
                raise ValueError("prediction conv expects (B, %d, H, W)" % cin)
            w, scale, shift = self._operands(3)
            xp = ops.pack_p_split(x, "NCHW")
        # identity activation, bias as shift; the GEMM's epilogue writes the reference's NCHW head map itself
        return ops.fusion_conv_nchw(xp, w, scale, shift, slope=1.0, channels=n)


class YOLOOutputV3(torch.nn.Module):
    """YOLO output layer V3 (yolo3.py:25-199), inference branch.

    ``__call__(x)`` applies the 1x1 ``prediction`` conv (yolo3.py:62,157; ``Prediction``: the library's own kernel) to
    the tip feature map and decodes it with the CUDA kernel into ``(B, C*H*W*A, 6)`` detections in the
    reference's row order.  ``decode(pred)`` skips the conv for an already computed head map.
    """

    def __init__(self, index, num_class, anchors, stride, alloc_size=(128, 128), agnostic=False,
                 in_channels: Optional[int] = None, **kwargs):
        super().__init__()
        anchors = np.array(anchors).astype("float32")                      # :46
        self._index = index
        self._classes = num_class
        self._num_pred = 1 + 4 + num_class                                  # :48
        self._num_anchors = anchors.size // 2                               # :49
        self._stride = stride
        self._agnostic = agnostic
        self._alloc_size = tuple(alloc_size)
        self._anchors = [float(v) for v in anchors.reshape(-1)]
        all_pred = self._num_pred * self._num_anchors                       # :57
        self.prediction = Prediction(in_channels, all_pred) if in_channels else None     # :62

    def _check(self, pred: torch.Tensor):
        if pred.shape[2] > self._alloc_size[0] or pred.shape[3] > self._alloc_size[1]:
            # the reference crops a (128,128) offset map (yolo3.py:67-74,168): larger maps cannot broadcast
            raise ValueError("feature map %s exceeds alloc_size %s" % (tuple(pred.shape[2:]), self._alloc_size))

    def decode(self, pred: torch.Tensor) -> torch.Tensor:
        self._check(pred)
        return ops.yolo3_decode([pred], self._classes, [self._anchors], [self._stride], self._agnostic)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.prediction is None:
            raise RuntimeError("YOLOOutputV3 built without in_channels: call decode(pred) with the head map")
        return self.decode(self.prediction(x))


class YOLOV3(torch.nn.Module):
    """Inference tail of the YOLOV3 family (yolo3.py:350-556; head-level like YOLOV3_noback :1686-1870).

    ``net(x1, x2, x3)`` takes the three inputs of the output layers in network order (stride 32, 16, 8:
    anchors/strides are used reversed, yolo3.py:416-417) and returns
    ``ids (B, post_nms, 1), scores (B, post_nms, 1), bboxes (B, post_nms, 4)`` like yolo3.py:531-534.
    With ``in_channels`` given the inputs are tip feature maps and go through the 1x1 prediction convs
    first; otherwise they are the head maps themselves.
    """

    def __init__(self, anchors=None, strides=None, classes: Sequence[str] = (), alloc_size=(128, 128),
                 nms_thresh=0.45, nms_topk=400, post_nms=100, agnostic=False,
                 in_channels: Optional[Sequence[int]] = None, **kwargs):
        super().__init__()
        anchors = ANCHORS if anchors is None else anchors
        strides = STRIDES if strides is None else strides
        self._classes = list(classes)
        self.nms_thresh = nms_thresh
        self.nms_topk = nms_topk
        self.post_nms = post_nms
        self._agnostic = agnostic
        self.valid_thresh = 0.01                                            # hard-coded at yolo3.py:527
        self.yolo_outputs = torch.nn.ModuleList()
        # note that anchors and strides should be used in reverse order (yolo3.py:415-417)
        for i, (anchor, stride) in enumerate(zip(anchors[::-1], strides[::-1])):
            ic = in_channels[i] if in_channels is not None else None
            self.yolo_outputs.append(YOLOOutputV3(i, len(self._classes), anchor, stride,
                                                  alloc_size=alloc_size, agnostic=agnostic, in_channels=ic))
        self.last_kept_rows = None

    @property
    def num_class(self):
        return len(self._classes)

    @property
    def classes(self):
        return self._classes

    def set_nms(self, nms_thresh=0.45, nms_topk=400, post_nms=100):
        """yolo3.py:536-556."""
        self.nms_thresh = nms_thresh
        self.nms_topk = nms_topk
        self.post_nms = post_nms

    def _heads(self, xs):
        if len(xs) != len(self.yolo_outputs):
            raise ValueError("expected %d inputs (stride 32,16,8 order)" % len(self.yolo_outputs))
        heads = []
        for x, out in zip(xs, self.yolo_outputs):
            h = out.prediction(x) if out.prediction is not None else x
            out._check(h)
            heads.append(h)
        return heads

    def forward(self, *xs):
        return self.forward_heads(*self._heads(xs))

    def forward_heads(self, *heads):
        """the tail from the head maps on (outputs of the prediction convs): decode, concat, box_nms, slice, split"""
        for h, out in zip(heads, self.yolo_outputs):
            out._check(h)
        C = len(self._classes)
        anchors = [o._anchors for o in self.yolo_outputs]
        strides = [o._stride for o in self.yolo_outputs]
        R = ops.n_rows(heads, C, self.yolo_outputs[0]._num_anchors, self._agnostic)
        nms_on = 0 < self.nms_thresh < 1                                     # yolo3.py:525
        k_eff = R if self.nms_topk < 0 else min(R, self.nms_topk)
        if nms_on and self.post_nms > 0 and 1 <= k_eff <= 1024:
            # fused: the (B, R, 6) tensor of yolo3.py:523 is never materialised
            result, kept = ops.yolo3_decode_nms(heads, C, anchors, strides, self.nms_thresh, self.valid_thresh,
                                                self.nms_topk, min(self.post_nms, R), agnostic=self._agnostic)
            self.last_kept_rows = kept
        else:
            result = ops.yolo3_decode(heads, C, anchors, strides, self._agnostic)        # :496,:523
            self.last_kept_rows = None
            if nms_on:
                rows = min(self.post_nms, R) if self.post_nms > 0 else None              # :529-530
                result, kept = ops.box_nms(result, overlap_thresh=self.nms_thresh, valid_thresh=self.valid_thresh,
                                           topk=self.nms_topk, id_index=0, score_index=1, coord_start=2,
                                           force_suppress=False, out_rows=rows, return_kept=True)  # :526-528
                self.last_kept_rows = kept
        ids = result[..., 0:1]                                                           # :531
        scores = result[..., 1:2]                                                        # :532
        bboxes = result[..., 2:6]                                                        # :533
        return ids, scores, bboxes


# the reference's YOLOV3_noback takes pre-extracted features; at head level the two coincide
YOLOV3_noback = YOLOV3


class YOLOV3Temporal(YOLOV3):
    """Inference tail of ``YOLOV3Temporal`` with ``t_out`` (yolo3_temporal.py:447-468, 540-555): the output layers run
    ``TimeDistributed`` over the window (:468), the detections of the scales are concatenated on axis -2 (:540) and
    ``box_nms`` runs over ``(B, T, R, 6)`` -- MXNet's operator treats every leading dimension as batch (:543-545) -- then
    ``slice_axis(axis=-2, 0, post_nms)`` (:548) and the split (:550-552).

    ``net(x1, x2, x3)`` takes the three inputs of the output layers, each (B, T, C_i, H_i, W_i) (head maps, or tip
    features with ``in_channels``; a 4-D input is T = 1 and behaves like ``YOLOV3``), and returns
    ``ids (B, T, post_nms, 1), scores (B, T, post_nms, 1), bboxes (B, T, post_nms, 4)``.  TimeDistributed folds T into
    the batch (layers.py:241-250), so this is the fused decode + box_nms of ``YOLOV3`` on B*T frames; ``last_kept_rows``
    is (B, T, post_nms)."""

    def forward(self, *xs):
        if xs[0].dim() != 5:
            return super().forward(*xs)
        B, T = xs[0].shape[0], xs[0].shape[1]
        for x in xs:
            if x.dim() != 5 or x.shape[0] != B or x.shape[1] != T:
                raise ValueError("every input must be (B=%d, T=%d, C, H, W)" % (B, T))
        flat = [x.reshape((B * T,) + tuple(x.shape[2:])) for x in xs]            # TimeDistributed: reshape (-3, -2)
        ids, scores, bboxes = super().forward(*flat)
        if self.last_kept_rows is not None:
            self.last_kept_rows = self.last_kept_rows.reshape(B, T, -1)
        unfold = lambda t: t.reshape((B, T) + tuple(t.shape[1:]))
        return unfold(ids), unfold(scores), unfold(bboxes)


class YOLOV3T(torch.nn.Module):
    """Post-backbone tail of the temporal detector YOLOV3T with a late join (yolo3.py:915-1302; the loop
    :1126-1177 and the NMS tail :1195-1206), i.e. BASELINE configs[2]: per scale

        tip  = Conv(block_conv_type, 2*channel, 3, 1, 1)(x)        the block's tip conv     yolo3.py:250-251,:1132
        tip  = 'max' | 'mean' TemporalPooling(k) or 'cat' reshape     late join                :1134-1138
        dets = YOLOOutputV3(tip)                                      1x1 prediction + decode  :1159
      then concat -> box_nms -> slice -> (ids, scores, bboxes)                                 :1195-1206

    ``net(x32, x16, x8)`` takes the three block-body outputs, each (B, K, channel_i, H_i, W_i) fp32 on a CUDA
    device (network order, channel_i = 512, 256, 128: wrappers.py:91-103).  The reference asserts that 3-D /
    2+1-D blocks need k > 1 and a late join (yolo3.py:979-985); so does this class.  The tip convs run in the
    tcgen05 fusion-conv kernel (bf16 operands and activations, fp32 accumulation); the 1x1 prediction conv runs in the
    same kernel on the joined bf16 tip with split, fp32-grade weights (``Prediction``) for all three join types -- there
    is no cuDNN / cuBLAS call on this path; decode + box_nms is the fused kernel path of YOLOV3.  Against the fp32
    reference the head logits therefore differ by what the bf16 ACTIVATIONS of the fusion conv cost (the stated fusion-conv
    tolerance, 1e-2 of the tensor's range), not by the prediction conv.
    """

    def __init__(self, classes: Sequence[str], k: int = 3, k_join_type: str = "max", block_conv_type: str = "3",
                 channels: Sequence[int] = (512, 256, 128), anchors=None, strides=None,
                 nms_thresh=0.45, nms_topk=400, post_nms=100, agnostic=False, k_join_pos: str = "late", **kwargs):
        super().__init__()
        from .layers import Conv, TemporalPooling
        assert k_join_pos in ("late", "early")                            # yolo3.py:984
        assert k_join_type in ("max", "mean", "cat")                      # yolo3.py:984
        self._late = k_join_pos == "late"
        if self._late:
            assert k > 1, "3-D and 2+1-D convolutions need a temporal window (yolo3.py:981-983)"
            assert block_conv_type in ("3", "21")
        else:
            # early join (yolo3.py:1107-1123): the window is joined right behind the backbone stages, everything after
            # it is the 2-D detector on one frame's worth of activations
            assert block_conv_type == "2", "after an early join the blocks are 2-D (yolo3.py:979-985)"
            k = 1
        self._k, self._join = k, k_join_type
        self.tips = torch.nn.ModuleList([Conv(block_conv_type, 2 * c, 3, 1, 1, in_channels=c) for c in channels])
        self.pools = (torch.nn.ModuleList([TemporalPooling(k, k_join_type) for _ in channels])
                      if k_join_type != "cat" and self._late else None)
        mult = k if k_join_type == "cat" else 1
        self.tail = YOLOV3(anchors, strides, classes=classes, nms_thresh=nms_thresh, nms_topk=nms_topk,
                           post_nms=post_nms, agnostic=agnostic, in_channels=[2 * c * mult for c in channels])

    @property
    def classes(self):
        return self.tail.classes

    def set_nms(self, nms_thresh=0.45, nms_topk=400, post_nms=100):
        self.tail.set_nms(nms_thresh, nms_topk, post_nms)

    #: the late 'max' join runs in the tip conv's epilogue (ops.fusion_conv(pool_max=True)): the un-pooled tip is never
    #: written and the TemporalPooling kernel does not run; False = conv, then TemporalPooling (same bits)
    fuse_max_join = True

    def _fused_join(self):
        return self.fuse_max_join and self._late and self._join == "max"

    def _tips(self, xs, joined=False):
        """tip conv outputs; with ``joined`` and a late 'max' join, the already joined frame (T = 1)"""
        if len(xs) != len(self.tips):
            raise ValueError("expected %d inputs (stride 32,16,8 order)" % len(self.tips))
        pool = joined and self._fused_join()
        tips = []
        for i, x in enumerate(xs):
            if isinstance(x, ops.PTensor):                 # already in the library's layout (YOLOV3TNeck)
                if x.T != self._k:
                    raise ValueError("input %d must hold K=%d frames" % (i, self._k))
                tips.append(self.tips[i](x, pool_max=pool))
                continue
            if x.dim() != 5 or x.shape[1] != self._k:
                raise ValueError("input %d must be (B, K=%d, C, H, W)" % (i, self._k))
            tips.append(self.tips[i](ops.pack_p(x, "NTCHW"), pool_max=pool))
        return tips

    def tip_features(self, *xs):
        """the three joined tip feature maps (B, C', H, W) fp32 that feed the output layers"""
        feats = []
        for i, tip in enumerate(self._tips(xs, joined=True)):
            if self._join == "cat" or not self._late:
                t = ops.unpack_p(tip, "NTCHW")                                  # (B, K, C, H, W)
                feats.append(t.reshape(t.shape[0], -1, t.shape[3], t.shape[4]))  # reshape (0,-3,-2): yolo3.py:1136
            else:
                feats.append(ops.unpack_p(tip if self._fused_join() else self.pools[i](tip), "NCHW"))
        return feats

    def head_maps(self, *xs):
        """the three NCHW head maps (B, A*(5+C), H, W) fp32 = outputs of the prediction convs (yolo3.py:157), computed
        by ``Prediction`` directly on the joined bf16 tip ('max' / 'mean': the pooled frame; 'cat': the K frames side by
        side in the channels); only the head map is converted to NCHW."""
        heads = []
        for i, tip in enumerate(self._tips(xs, joined=True)):
            joined = tip if self._join == "cat" or not self._late or self._fused_join() else self.pools[i](tip)
            heads.append(self.tail.yolo_outputs[i].prediction(joined))
        return heads

    def forward(self, *xs):
        return self.tail.forward_heads(*self.head_maps(*xs))

    @property
    def last_kept_rows(self):
        return self.tail.last_kept_rows


class YOLOV3TNeck(torch.nn.Module):
    """Everything of ``YOLOV3T.hybrid_forward`` after the backbone stages (yolo3.py:1104-1206), late join (default) or,
    with ``k_join_pos='early'`` and ``block_conv_type='2'``, the early join of :1107-1123 followed by the 2-D neck: per scale the
    detection block (:1131-1132), the late join of its tip (:1134-1138), the output layer (:1159); between scales the
    1x1 transition (:1167, a TimeDistributed 2-D cell :1050), ``_upsample`` x2 + ``slice_like`` + channel concat with the
    backbone's route of the next scale (:1170-1175); then concat -> box_nms -> slice (:1195-1206).

    ``net(r32, r16, r8)`` takes the three stage outputs, each (B, K, C_i, H_i, W_i) fp32 on a CUDA device (deep to
    shallow; Darknet-53: 1024, 512, 256 channels).  Every convolution runs in the tcgen05 fusion-conv kernel on bf16
    P-layout activations; the join between scales is ``vy_upsample_concat_bf16``; tip / join / prediction / decode / NMS
    are the ``YOLOV3T`` tail above."""

    def __init__(self, classes: Sequence[str], k: int = 3, k_join_type: str = "max", block_conv_type: str = "3",
                 stage_channels: Sequence[int] = (1024, 512, 256), channels: Sequence[int] = (512, 256, 128),
                 k_join_pos: str = "late", **kwargs):
        super().__init__()
        from .layers import Conv, TemporalPooling, TimeDistributed, YOLODetectionBlockV3
        assert len(stage_channels) == len(channels)
        assert k_join_pos in ("late", "early")
        self._k, self._early, self._join = k, k_join_pos == "early", k_join_type
        if self._early:
            # early join (yolo3.py:1107-1123): every stage output is joined over the window first -- 'cat': reshape
            # (0,-3,-2) = K*C channels (:1110), 'max' / 'mean': TemporalPooling (:1112) -- and the rest is the 2-D neck
            self.join_pool = TemporalPooling(k, k_join_type) if k_join_type != "cat" else None
            stage_channels = [c * (k if k_join_type == "cat" else 1) for c in stage_channels]
        self.head = YOLOV3T(classes, k=k, k_join_type=k_join_type, block_conv_type=block_conv_type, channels=channels,
                            k_join_pos=k_join_pos, **kwargs)
        blocks, transitions, cin = [], [], stage_channels[0]
        for i, c in enumerate(channels):
            blocks.append(YOLODetectionBlockV3(c, block_conv_type, in_channels=cin))
            if i + 1 < len(channels):
                transitions.append(TimeDistributed(Conv("2", channels[i + 1], 1, 0, 1, in_channels=c)))   # yolo3.py:1048-1051
                cin = channels[i + 1] + stage_channels[i + 1]
        self.blocks = torch.nn.ModuleList(blocks)
        self.transitions = torch.nn.ModuleList(transitions)
        for i, b in enumerate(self.blocks):               # the block's tip conv IS the head's tip conv (one set of weights)
            b.tip = self.head.tips[i]

    @property
    def classes(self):
        return self.head.classes

    def set_nms(self, nms_thresh=0.45, nms_topk=400, post_nms=100):
        self.head.set_nms(nms_thresh, nms_topk, post_nms)

    def routes(self, *rs):
        """the three block-body outputs (``route`` of yolo3.py:258) as P-layout activations, deep to shallow"""
        if len(rs) != len(self.blocks):
            raise ValueError("expected %d stage outputs (deep to shallow)" % len(self.blocks))
        outs = []
        x = self._stage(rs[0])
        for i, block in enumerate(self.blocks):
            for cell in block.body:
                x = cell(x)
            outs.append(x)
            if i + 1 < len(self.blocks):
                x = ops.upsample_concat(self.transitions[i](x), self._stage(rs[i + 1]))
        return outs

    def _stage(self, r):
        """a stage output (B, K, C, H, W) as a P-layout activation; with an early join, joined over the window"""
        if r.dim() != 5 or r.shape[1] != self._k:
            raise ValueError("stage outputs must be (B, K=%d, C, H, W)" % self._k)
        x = ops.pack_p(r, "NTCHW")
        if not self._early:
            return x
        return ops.cat_repeat(x, 1) if self._join == "cat" else self.join_pool(x)

    def forward(self, *rs):
        return self.head(*self.routes(*rs))

    @property
    def last_kept_rows(self):
        return self.head.last_kept_rows


def get_yolov3_postprocess(classes, agnostic=False, in_channels=None, **kwargs) -> YOLOV3:
    """Head-level counterpart of wrappers.yolo3_darknet53 (wrappers.py:9-110): reference anchors/strides."""
    return YOLOV3(ANCHORS, STRIDES, classes=classes, agnostic=agnostic, in_channels=in_channels, **kwargs)
