"""videoyolo_b200 -- B200-native (sm_100a) detection post-processing path of VideoYOLO.

decode (YOLOOutputV3) -> temporal fusion conv (Conv) -> box_nms -> (ids, scores, bboxes), behind the
reference's own block/operator names.  Host code is Python over a C ABI (include/vyolo.h, ctypes);
PyTorch supplies device buffers, streams and torch.distributed only.  No CPU fallback.
"""
from . import _lib, ops, parallel
from .ops import bbox_batch_iou, bbox_iou, box_nms, yolo3_decode, yolo3_decode_nms
from .layers import Conv, Conv1D, TemporalPooling, TimeDistributed, YOLODetectionBlockV3
from .yolo3 import (ANCHORS, STRIDES, YOLOOutputV3, YOLOV3, YOLOV3T, YOLOV3Temporal, YOLOV3TNeck, YOLOV3_noback,
                    get_yolov3_postprocess)

__all__ = ["bbox_batch_iou", "bbox_iou", "box_nms", "yolo3_decode", "yolo3_decode_nms", "YOLOOutputV3", "YOLOV3",
           "YOLOV3_noback", "YOLOV3T", "YOLOV3Temporal", "YOLOV3TNeck", "YOLODetectionBlockV3", "Conv", "Conv1D", "TemporalPooling", "TimeDistributed", "get_yolov3_postprocess",
           "ANCHORS", "STRIDES", "ops", "parallel"]
