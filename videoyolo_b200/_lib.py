"""ctypes binding of libvyolo.so (the C ABI declared in include/vyolo.h).

There is no fallback: if the CUDA library is missing or a tensor is not on a CUDA device the call
raises.  PyTorch is used for device buffers and streams only.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VYOLO_LIB_VARIANT: development builds of tools/ (videoyolo_b200.build.build(variant=...)); unset in production
_VARIANT = os.environ.get("VYOLO_LIB_VARIANT", "")
SO_PATH = os.path.join(_HERE, "libvyolo%s.so" % ("_" + _VARIANT if _VARIANT else ""))
_LIB = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int)
c_vp = ctypes.c_void_p

# every symbol include/vyolo.h declares, with its signature (tests check the export list against this)
SIGNATURES = {
    "vy_version": (ctypes.c_int, []),
    "vy_last_error": (ctypes.c_char_p, []),
    "vy_kernel_name": (ctypes.c_char_p, [ctypes.c_int]),
    "vy_launch_counts": (ctypes.c_int, [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]),
    "vy_prof_enable": (ctypes.c_int, [ctypes.c_int]),
    "vy_prof_read": (ctypes.c_int, [c_f64p, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]),
    "vy_decode_f32": (ctypes.c_int, [ctypes.POINTER(c_vp), c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "vy_box_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int]),
    "vy_box_nms_f32": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                      c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "vy_decode_nms_workspace_bytes": (ctypes.c_size_t, [c_i32p, c_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                        ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "vy_decode_nms_f32": (ctypes.c_int, [ctypes.POINTER(c_vp), c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "vy_decode_nms_plan_create": (ctypes.c_int, [c_i32p, c_i32p, c_f32p, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "vy_decode_nms_plan_workspace_bytes": (ctypes.c_size_t, [c_vp]),
    "vy_decode_nms_plan_launch": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "vy_decode_nms_plan_destroy": (None, [c_vp]),
    "vy_bbox_iou_f32": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, c_vp, c_vp]),
    "vy_bbox_iou_f64": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, c_vp, c_vp]),
    "vy_fusion_conv_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 9),
    "vy_fusion_conv_bf16": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_float] + [ctypes.c_int] * 9 +
                            [c_vp, ctypes.c_int, c_vp, ctypes.c_size_t, c_vp]),
    "vy_fusion_conv_bf16_nchw_joined": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_float] + [ctypes.c_int] * 7 +
                                        [c_vp, ctypes.c_int, c_vp]),
    "vy_fusion_conv_bf16_maxpool": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_float] + [ctypes.c_int] * 9 + [c_vp, c_vp]),
    "vy_fusion_conv_bf16_nchw": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_float] + [ctypes.c_int] * 7 +
                                 [c_vp, ctypes.c_int, c_vp]),
    "vy_anchor_match_f32": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp, c_vp]),
    "vy_bbox_batch_iou_f32": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "vy_upsample_concat_bf16": (ctypes.c_int, [c_vp, c_vp] + [ctypes.c_int] * 8 + [c_vp, c_vp]),
    "vy_detect_consume_f32": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp,
                                             c_vp, c_vp]),
    "vy_hier_nms_f32": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                       c_vp, c_vp, c_vp]),
    "vy_voc_match_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp] + [ctypes.c_int] * 4 + [ctypes.c_float] + [c_vp] * 6),
    "vy_temporal_pool_bf16": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_long, ctypes.c_int, c_vp, c_vp]),
    "vy_temporal_dwconv_bf16": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "vy_p_layout_elems": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "vy_pack_f32_to_p_bf16": (ctypes.c_int, [c_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong] +
                              [ctypes.c_int] * 5 + [c_vp, c_vp]),
    "vy_pack_f32_split_to_p_bf16": (ctypes.c_int, [c_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong] +
                                    [ctypes.c_int] * 6 + [c_vp, c_vp]),
    "vy_cat_repeat_bf16": (ctypes.c_int, [c_vp] + [ctypes.c_int] * 6 + [c_vp, c_vp]),
    "vy_unpack_p_to_f32": (ctypes.c_int, [c_vp] + [ctypes.c_int] * 6 + [c_vp, ctypes.c_longlong, ctypes.c_longlong,
                                                                        ctypes.c_longlong, c_vp]),
    "vy_unpack_p_channels_to_f32": (ctypes.c_int, [c_vp] + [ctypes.c_int] * 7 + [c_vp, ctypes.c_longlong, ctypes.c_longlong,
                                                                                 ctypes.c_longlong, c_vp]),
}

VY_OK = 0
ERRORS = {-1: "VY_EINVAL", -2: "VY_EALIGN", -3: "VY_EWORKSPACE", -4: "VY_ECUDA", -5: "VY_EUNSUPPORTED"}


class VyoloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "VY_E?"), code, msg))
        self.code = code


def lib() -> ctypes.CDLL:
    """Load libvyolo.so.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "videoyolo_b200: %s is missing - build it with `python -m videoyolo_b200.build` "
                "(or __graft_entry__.build()); there is no CPU fallback" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.vy_version() != 1:
            raise RuntimeError("videoyolo_b200: ABI version mismatch")
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != VY_OK:
        raise VyoloError(rc, lib().vy_last_error().decode("utf-8", "replace"))


N_KERNEL_IDS = 13


def launch_counts() -> dict:
    """Cumulative kernel launches of the library per kernel name (include/vyolo.h: vy_launch_counts)."""
    c = (ctypes.c_longlong * N_KERNEL_IDS)()
    lib().vy_launch_counts(c, N_KERNEL_IDS)
    return {lib().vy_kernel_name(i).decode(): int(c[i]) for i in range(N_KERNEL_IDS)}


def prof_enable(on: bool) -> None:
    check(lib().vy_prof_enable(int(bool(on))))


def prof_read() -> dict:
    """{kernel name: (summed ms, launches)} of the launches recorded since the last read."""
    ms = (ctypes.c_double * N_KERNEL_IDS)()
    n = (ctypes.c_longlong * N_KERNEL_IDS)()
    check(lib().vy_prof_read(ms, n, N_KERNEL_IDS))
    return {lib().vy_kernel_name(i).decode(): (float(ms[i]), int(n[i])) for i in range(N_KERNEL_IDS) if n[i]}
