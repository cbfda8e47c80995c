"""Host-facing detection call: pinned host head maps in, host (ids, scores, bboxes) out.

This is the call a user of the reference makes around ``net(x)``: ``split_and_load`` (host -> device,
detect_yolo3.py:211-213), the forward tail, and ``as_numpy`` (device -> host, utils/general.py:6-18,
detect_yolo3.py:233).  The batch is cut into chunks that travel through a small ring of device
buffers: the copy engine brings chunk i+1 in on one stream while the decode+NMS kernels of chunk i
run on another, and the (chunk, post_nms, 6) results leave on the compute stream.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class HostDetector:
    """Fused decode + box_nms for head maps that live in (pinned) host memory."""

    def __init__(self, num_class: int, anchors, strides, device, chunk: int = 8, ring: int = 3,
                 nms_thresh: float = 0.45, valid_thresh: float = 0.01, nms_topk: int = 400,
                 post_nms: int = 100, agnostic: bool = False):
        self.C = num_class
        self.anchors, self.strides = anchors, strides
        self.device = torch.device(device)
        self.chunk, self.ring = int(chunk), int(ring)
        self.nms = dict(nms_thresh=nms_thresh, valid_thresh=valid_thresh, topk=nms_topk,
                        post_nms=post_nms, agnostic=agnostic)
        self.post_nms = post_nms
        self._copy = torch.cuda.Stream(self.device)
        self._compute = torch.cuda.Stream(self.device)
        self._bufs: Optional[List[List[torch.Tensor]]] = None
        self._shapes = None
        self._out_dev = self._kept_dev = self._out_host = self._kept_host = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _prepare(self, heads: Sequence[torch.Tensor]):
        shapes = [tuple(h.shape) for h in heads]
        if shapes == self._shapes:
            return
        B = shapes[0][0]
        self._bufs = [[torch.empty((self.chunk,) + s[1:], dtype=torch.float32, device=self.device) for s in shapes]
                      for _ in range(self.ring)]
        self._free = [torch.cuda.Event() for _ in range(self.ring)]      # chunk buffer consumed
        self._ready = [torch.cuda.Event() for _ in range(self.ring)]     # chunk buffer filled
        self._out_dev = torch.empty((B, self.post_nms, 6), dtype=torch.float32, device=self.device)
        self._kept_dev = torch.empty((B, self.post_nms), dtype=torch.int32, device=self.device)
        self._out_host = torch.empty((B, self.post_nms, 6), dtype=torch.float32).pin_memory()
        self._kept_host = torch.empty((B, self.post_nms), dtype=torch.int32).pin_memory()
        self._shapes = shapes
        self.h2d_bytes = sum(h.numel() * 4 for h in heads)
        self.d2h_bytes = self._out_host.numel() * 4 + self._kept_host.numel() * 4

    def __call__(self, heads: Sequence[torch.Tensor], copy: bool = True):
        """heads: host tensors (B, A*(5+C), H, W) fp32 in network order (stride 32, 16, 8); pinned memory
        makes the copies asynchronous.  Returns host tensors ids (B,post,1), scores (B,post,1),
        bboxes (B,post,4) and leaves the kept source rows in ``self.kept_rows`` (host, int32).

        ``copy=True`` (default) returns FRESH host tensors, like the reference's ``as_numpy``
        (utils/general.py:6-18): detect_yolo3.py:233-262 appends every batch's results to a list, which
        must not alias the next batch.  ``copy=False`` returns views of the detector's persistent pinned
        result buffer, which the next call overwrites (the zero-allocation loop of bench.py)."""
        for h in heads:
            if h.is_cuda or h.dtype != torch.float32 or not h.is_contiguous():
                raise TypeError("HostDetector takes contiguous fp32 host tensors; device tensors go through YOLOV3")
        self._prepare(heads)
        B = heads[0].shape[0]
        start = torch.cuda.current_stream(self.device)
        self._copy.wait_stream(start)
        self._compute.wait_stream(start)
        n_chunks = (B + self.chunk - 1) // self.chunk
        for i in range(n_chunks):
            s, e = i * self.chunk, min(B, (i + 1) * self.chunk)
            slot = i % self.ring
            with torch.cuda.stream(self._copy):
                if i >= self.ring:
                    self._copy.wait_event(self._free[slot])
                for buf, h in zip(self._bufs[slot], heads):
                    buf[: e - s].copy_(h[s:e], non_blocking=True)
                self._ready[slot].record(self._copy)
            with torch.cuda.stream(self._compute):
                self._compute.wait_event(self._ready[slot])
                ops.yolo3_decode_nms([b[: e - s] for b in self._bufs[slot]], self.C, self.anchors, self.strides,
                                     out=self._out_dev[s:e], kept=self._kept_dev[s:e], **self.nms)
                self._free[slot].record(self._compute)
        with torch.cuda.stream(self._compute):
            self._out_host.copy_(self._out_dev, non_blocking=True)
            self._kept_host.copy_(self._kept_dev, non_blocking=True)
        self._compute.synchronize()                                     # as_numpy: the host needs the values
        start.wait_stream(self._compute)
        self.kept_rows = self._kept_host.clone() if copy else self._kept_host
        r = self._out_host.clone() if copy else self._out_host
        return r[..., 0:1], r[..., 1:2], r[..., 2:6]


class GraphedDetector:
    """Fused decode + box_nms for device-resident head maps of a FIXED shape, captured once in a CUDA graph:
    a call is one ``cudaGraphLaunch`` instead of four kernel launches through ctypes.  Meant for
    the launch-bound end of the path (small batches: BASELINE configs[0] is batch 1, where the kernels take
    tens of microseconds).  The library needs nothing special for this: it never allocates or synchronises
    and every launch goes to the stream it is handed, so its calls are capturable as they are.

    ``det(heads)`` copies the heads into the graph's static inputs (device-to-device) and replays;
    ``det()`` replays on whatever ``det.heads`` currently hold.  Returns ``(out (B, post_nms, 6), kept)``,
    static tensors that the next replay overwrites.  The replay is enqueued on the CALLER's current stream; one
    object = one workspace = one replay in flight at a time (use one object per stream for concurrent batches).
    """

    def __init__(self, num_class: int, anchors, strides, shapes: Sequence[Sequence[int]], device,
                 nms_thresh: float = 0.45, valid_thresh: float = 0.01, nms_topk: int = 400,
                 post_nms: int = 100, agnostic: bool = False):
        self.device = torch.device(device)
        self.heads = [torch.zeros(tuple(s), dtype=torch.float32, device=self.device) for s in shapes]
        B = self.heads[0].shape[0]
        self.out = torch.empty((B, post_nms, 6), dtype=torch.float32, device=self.device)
        self.kept = torch.empty((B, post_nms), dtype=torch.int32, device=self.device)
        kw = dict(nms_thresh=nms_thresh, valid_thresh=valid_thresh, topk=nms_topk, post_nms=post_nms, agnostic=agnostic)
        # the graph's kernels write to THIS buffer: owned here, not the per-stream cache of ops (which another caller
        # on the same stream handle may outgrow and replace while the graph still points at it)
        plan = ops.decode_nms_plan([h.shape for h in self.heads], num_class, anchors, strides, nms_thresh, valid_thresh,
                                   nms_topk, post_nms, False, agnostic)
        self.workspace = torch.empty(max(plan.workspace_bytes, 256), dtype=torch.uint8, device=self.device)
        kw["workspace_buf"] = self.workspace
        self._stream = torch.cuda.Stream(self.device)
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):                 # warm-up: loads the kernels, sets their attributes
            ops.yolo3_decode_nms(self.heads, num_class, anchors, strides, out=self.out, kept=self.kept, **kw)
        self._stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._stream):
            ops.yolo3_decode_nms(self.heads, num_class, anchors, strides, out=self.out, kept=self.kept, **kw)

    def __call__(self, heads: Optional[Sequence[torch.Tensor]] = None):
        if heads is not None:
            for dst, src in zip(self.heads, heads):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out, self.kept


class GraphedModule:
    """Any block of this package (``YOLOV3T``, ``YOLOV3TNeck``, ...) for inputs of a FIXED shape, captured once in a CUDA
    graph: the temporal tail / neck is a chain of ~15-45 short launches through ctypes, and a replay removes the host
    from between them.  The library never allocates or synchronises and the blocks' intermediate tensors come from the
    graph's private pool (stable addresses, so the TMA descriptors built at capture time stay valid).

    ``g(*inputs)`` copies the inputs into the static ones (device-to-device) and replays; returns the static outputs
    (overwritten by the next replay)."""

    def __init__(self, module, example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        self.module = module
        self.inputs = [x.clone() for x in example_inputs]
        dev = self.inputs[0].device
        self._stream = torch.cuda.Stream(dev)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        self._workspaces = {}                                  # scratch buffers the captured kernels write to: owned here
        with ops.workspace_scope(self._workspaces):
            with torch.cuda.stream(self._stream), torch.no_grad():
                for _ in range(max(1, warmup)):                # sizes the workspaces, packs weights, folds BN
                    module(*self.inputs)
            self._stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self._stream), torch.no_grad():
                self.outputs = module(*self.inputs)

    def __call__(self, *inputs):
        if inputs:
            if len(inputs) != len(self.inputs):
                raise ValueError("expected %d inputs" % len(self.inputs))
            for dst, src in zip(self.inputs, inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
