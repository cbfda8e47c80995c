"""Host-side mirror of the reference's temporal-fusion layers (models/definitions/layers.py), inference only.

    Conv(type, channel, kernel, padding, stride)     layers.py:135-158  ('2' | '3' | '21' = conv + BN + LeakyReLU cells)
    TemporalPooling(k, type)                          layers.py:161-205  ('direct' style: max / mean over the K frames)
    TimeDistributed(model)                            layers.py:208-264  (fold K into the batch axis)
    Conv1D(channel, kernel)                           layers.py:50-60    (_conv1d: depthwise temporal merge of a window)
    YOLODetectionBlockV3(channel, conv_type)          yolo3.py:200-263   (body of five cells -> route, tip conv -> tip)

Activations travel between these blocks as ``ops.PTensor`` (the library's P layout, include/vyolo.h): pack the
reference's (B, K, C, H, W) / NCDHW fp32 tensor once with ``ops.pack_p`` and unpack the last output with
``ops.unpack_p``.  The arithmetic is the tcgen05/TMEM implicit-GEMM kernel behind ``vy_fusion_conv_bf16``: bf16
operands, fp32 accumulation, BatchNorm folded at call time from the block's (gamma, beta, running_mean, running_var).
Gluon infers input channels lazily; here ``in_channels`` is an explicit constructor argument.
"""
from __future__ import annotations

import torch

from . import ops


class _Cell(torch.nn.Module):
    """One ConvND(use_bias=False) + BatchNorm(eps 1e-5) + LeakyReLU(0.1) cell (layers.py:63-79)."""

    def __init__(self, in_channels, channel, k3, slope=0.1):
        super().__init__()
        self.k3 = tuple(k3)
        # MXNet's default initialiser for conv weights is Uniform(0.07); BN gamma 1, beta 0, mean 0, var 1
        self.weight = torch.nn.Parameter(torch.empty((channel, in_channels) + self.k3).uniform_(-0.07, 0.07))
        self.gamma = torch.nn.Parameter(torch.ones(channel))
        self.beta = torch.nn.Parameter(torch.zeros(channel))
        self.register_buffer("running_mean", torch.zeros(channel))
        self.register_buffer("running_var", torch.ones(channel))
        self.slope = slope
        self._packed = None

    def packed(self):
        """(weight in the kernel's (Cout, kt, kh, kw, Cin) bf16 layout, folded BN scale, shift), cached."""
        key = (self.weight._version, self.gamma._version, self.beta._version,
               self.running_mean._version, self.running_var._version, self.weight.device)
        if self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                w = ops.conv_weight(self.weight.detach())
                scale, shift = ops.fold_bn(self.gamma.detach(), self.beta.detach(), self.running_mean, self.running_var)
            self._packed = (key, w, scale, shift)
        return self._packed[1:]

    def forward(self, x: ops.PTensor, out_f32: bool = False, pool_max: bool = False) -> ops.PTensor:
        w, scale, shift = self.packed()
        return ops.fusion_conv(x, w, scale, shift, self.slope, out_f32=out_f32, pool_max=pool_max)


class Conv(torch.nn.Module):
    """Convolution helper layer, 2d / 3d / 2+1d (layers.py:135-158), on P-layout activations.

    '2'  : Conv2D(kernel, padding) + BN + LeakyReLU                      (layers.py:63-70); T must be 1
    '3'  : Conv3D(kernel^3, padding^3) + BN + LeakyReLU                  (layers.py:73-79)
    '21' : (1,k,k) conv + BN + LReLU then (k,1,1) conv + BN + LReLU, m = channel   (layers.py:82-89,154-155)
    Only what the reference uses is served: stride 1, padding = kernel // 2, kernel in {1, 3}.
    """

    def __init__(self, type, channel, kernel, padding, stride, in_channels=None, **kwargs):
        super().__init__()
        assert type in ["2", "3", "21"]                                   # layers.py:142
        if in_channels is None:
            raise ValueError("in_channels is required (Gluon infers it at the first call; this mirror does not)")
        if stride != 1 or padding != kernel // 2 or kernel not in (1, 3):
            raise ValueError("the fusion-conv kernel serves stride 1, 'same' padding, kernel 1 or 3")
        self._type = type
        k = kernel
        if type == "2":
            self.cells = torch.nn.ModuleList([_Cell(in_channels, channel, (1, k, k))])
        elif type == "3":
            self.cells = torch.nn.ModuleList([_Cell(in_channels, channel, (k, k, k))])
        else:
            self.cells = torch.nn.ModuleList([_Cell(in_channels, channel, (1, k, k)), _Cell(channel, channel, (k, 1, 1))])

    def forward(self, x: ops.PTensor, out_f32: bool = False, pool_max: bool = False) -> ops.PTensor:
        """``pool_max``: the conv followed by ``TemporalPooling(k, 'max')`` -- the last cell's epilogue does the join."""
        if self._type == "2" and x.T != 1:
            raise ValueError("Conv('2') takes a 2-D activation (T == 1); fold time first (TimeDistributed / 'cat')")
        for i, cell in enumerate(self.cells):
            last = i == len(self.cells) - 1
            x = cell(x, out_f32=out_f32 and last, pool_max=pool_max and last)
        return x


class TemporalPooling(torch.nn.Module):
    """'direct'-style temporal pooling (layers.py:161-205): max or mean over the K frames."""

    def __init__(self, k, type="max", pool_size=None, strides=None, padding=0, style="direct", **kwargs):
        super().__init__()
        assert type in ["max", "mean"]                                    # layers.py:168
        assert style in ["direct", "layer"]
        if style == "layer" or pool_size is not None:
            raise NotImplementedError("only the 'direct' style (what the detection models use) is served")
        self._k, self._type = k, type

    def forward(self, x: ops.PTensor) -> ops.PTensor:
        if x.T != self._k:
            raise ValueError("TemporalPooling(k=%d) got %d frames" % (self._k, x.T))
        return ops.temporal_pool(x, self._type)


class TimeDistributed(torch.nn.Module):
    """Apply a 2-D block to every frame (layers.py:208-264): in the P layout time is the outermost axis, so
    folding K into the batch (layers.py:241-250) is a view, not a copy."""

    def __init__(self, model, **kwargs):
        super().__init__()
        self.model = model

    def forward(self, x: ops.PTensor) -> ops.PTensor:
        folded = ops.PTensor(x.data.view((1, x.T * x.B) + tuple(x.data.shape[2:])), x.T * x.B, 1, x.H, x.W, x.C)
        y = self.model(folded)
        return ops.PTensor(y.data.view((x.T, x.B) + tuple(y.data.shape[2:])), x.B, x.T, y.H, y.W, y.C)


class Conv1D(torch.nn.Module):
    """_conv1d (layers.py:50-60): depthwise Conv3D(kernel (k,1,1), groups=channels, zero-initialised) + BN +
    LeakyReLU over a window of exactly k frames, the temporal merge of HDarknet (h_darknet.py:97-119)."""

    def __init__(self, out_channels, kernel, padding=0, strides=1, **kwargs):
        super().__init__()
        if padding != 0 or strides != 1:
            raise ValueError("the reference only builds _conv1d(c, w, 0, 1) (h_darknet.py:100)")
        self.weight = torch.nn.Parameter(torch.zeros((out_channels, 1, kernel, 1, 1)))   # weight_initializer='zeros'
        self.gamma = torch.nn.Parameter(torch.ones(out_channels))
        self.beta = torch.nn.Parameter(torch.zeros(out_channels))
        self.register_buffer("running_mean", torch.zeros(out_channels))
        self.register_buffer("running_var", torch.ones(out_channels))

    def forward(self, x: ops.PTensor) -> ops.PTensor:
        with torch.no_grad():
            scale, shift = ops.fold_bn(self.gamma.detach(), self.beta.detach(), self.running_mean, self.running_var)
        return ops.temporal_dwconv(x, self.weight.detach(), scale, shift, 0.1)


class YOLODetectionBlockV3(torch.nn.Module):
    """YOLO V3 detection block (models/definitions/yolo/yolo3.py:200-263): body = 2 x [1x1 reduce to ``channel``,
    3x3 expand to ``2*channel``] + 1x1 reduce (:225-247), tip = 3x3 expand (:250-253); returns ``(route, tip)``.
    ``conv_type`` '3' / '21': the 1x1 cells are 1x1x1 3-D convs and the 3x3 cells ``Conv(conv_type, ...)`` (:229-243);
    '2': all cells 2-D (the block then sits under ``TimeDistributed`` in the temporal models, yolo3.py:1035-1037).
    P-layout activations in and out (the reference's swapaxes(1, 2) :256-262 is a no-op in that layout)."""

    def __init__(self, channel, conv_type="2", in_channels=None, **kwargs):
        super().__init__()
        assert channel % 2 == 0, "channel {} cannot be divided by 2".format(channel)      # yolo3.py:222
        assert conv_type in ("2", "3", "21")
        if in_channels is None:
            raise ValueError("in_channels is required (Gluon infers it at the first call; this mirror does not)")
        self._conv_type = conv_type
        one = "3" if conv_type in ("3", "21") else "2"
        cells, cin = [], in_channels
        for _ in range(2):
            cells.append(Conv(one, channel, 1, 0, 1, in_channels=cin))
            cells.append(Conv(conv_type, channel * 2, 3, 1, 1, in_channels=channel))
            cin = channel * 2
        cells.append(Conv(one, channel, 1, 0, 1, in_channels=cin))
        self.body = torch.nn.ModuleList(cells)
        self.tip = Conv(conv_type, channel * 2, 3, 1, 1, in_channels=channel)

    def forward(self, x: ops.PTensor):
        route = x
        for cell in self.body:
            route = cell(route)
        return route, self.tip(route)
