"""numpy-fp32 stand-in for the slice of MXNet/Gluon that the reference's YOLO output path touches.

TEST INFRASTRUCTURE, build container only.  MXNet and gluoncv are not installable here, but the
reference's decode (`models/definitions/yolo/yolo3.py:130-199`) and tail (`:448-534`) are plain
Python over ``F.*`` / NDArray methods.  ``install()`` puts stub ``mxnet`` / ``gluoncv`` modules in
``sys.modules`` so that the reference file IMPORTS UNMODIFIED from /root/reference, and ``F`` /
``NDArray`` below execute the handful of array ops it calls with MXNet's semantics in numpy fp32:

  NDArray.reshape with MXNet's special codes 0 / -1 / -2 / -3 / -4, transpose(axes=), slice_axis,
  expand_dims, repeat, arithmetic;  F.sigmoid = 1/(1+exp(-x)) (mshadow_op::sigmoid), F.exp,
  broadcast_add/mul, concat(dim=), tile(reps=), transpose, arange, reshape, slice_like(axes=),
  zeros_like;  F.contrib.box_nms is NOT MXNet's kernel (its source is not in /root/reference):
  it is handed in by the caller (the oracle's box_nms) and flagged as such in the fixture provenance.

Nothing here is imported by the product or by the GPU tests; make_golden.py is the only user.
"""
from __future__ import annotations

import sys
import types

import numpy as np

f32 = np.float32


def _mx_reshape(shape_in, spec):
    """MXNet NDArray.reshape special values (python/mxnet/ndarray/ndarray.py docstring):
    0 copy this dim; -1 infer; -2 copy all remaining dims; -3 merge two consecutive dims;
    -4 split one dim into the next two values (one of which may be -1)."""
    out, i, j, spec = [], 0, 0, list(spec)
    infer = None
    while j < len(spec):
        s = spec[j]
        if s > 0:
            out.append(s); i += 1
        elif s == 0:
            out.append(shape_in[i]); i += 1
        elif s == -1:
            assert infer is None
            infer = len(out); out.append(-1); i += 1
        elif s == -2:
            out.extend(shape_in[i:]); i = len(shape_in)
        elif s == -3:
            out.append(shape_in[i] * shape_in[i + 1]); i += 2
        elif s == -4:
            a, b = spec[j + 1], spec[j + 2]
            d = shape_in[i]
            if a == -1:
                a = d // b
            if b == -1:
                b = d // a
            assert a * b == d
            out.extend([a, b]); i += 1; j += 2
        else:
            raise ValueError(s)
        j += 1
    return tuple(out)


class NDArray:
    """float32 array with the NDArray methods yolo3.py calls."""
    __array_priority__ = 100

    def __init__(self, a):
        self.a = np.asarray(a.a if isinstance(a, NDArray) else a, dtype=f32)

    @property
    def shape(self):
        return self.a.shape

    def asnumpy(self):
        return self.a.copy()

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if "shape" in kw:
            shape = tuple(kw["shape"])
        return NDArray(self.a.reshape(_mx_reshape(self.a.shape, shape)))

    def transpose(self, *axes, **kw):
        if "axes" in kw:
            axes = kw["axes"]
        elif len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = axes[0]
        return NDArray(np.ascontiguousarray(np.transpose(self.a, axes if len(axes) else None)))

    def slice_axis(self, axis, begin, end):
        sl = [slice(None)] * self.a.ndim
        sl[axis] = slice(begin, end)
        return NDArray(self.a[tuple(sl)])

    def expand_dims(self, axis):
        return NDArray(np.expand_dims(self.a, axis))

    def repeat(self, repeats, axis=None):
        return NDArray(np.repeat(self.a, repeats, axis=axis))

    def swapaxes(self, dim1, dim2):
        return NDArray(np.swapaxes(self.a, dim1, dim2))

    def clip(self, a_min, a_max):
        return NDArray(np.clip(self.a, f32(a_min), f32(a_max)))

    @staticmethod
    def _v(o):
        return o.a if isinstance(o, NDArray) else f32(o)

    # MXNet's elementwise operators on fp32 arrays: every result is rounded to fp32
    def __add__(self, o): return NDArray(self.a + self._v(o))
    def __radd__(self, o): return NDArray(self._v(o) + self.a)
    def __sub__(self, o): return NDArray(self.a - self._v(o))
    def __rsub__(self, o): return NDArray(self._v(o) - self.a)
    def __mul__(self, o): return NDArray(self.a * self._v(o))
    def __rmul__(self, o): return NDArray(self._v(o) * self.a)
    def __truediv__(self, o): return NDArray(self.a / self._v(o))
    def __neg__(self): return NDArray(-self.a)


def _expf(a):
    """libm expf as MXNet's CPU kernels call it: glibc's expf evaluates in double and is correctly
    rounded (but for ~1e-9 of arguments), so exp in float64 rounded once to fp32 reproduces it and
    does not depend on which SIMD exp the installed numpy ships for float32."""
    return np.exp(a.astype(np.float64)).astype(f32)


class _Contrib:
    def __init__(self):
        self.box_nms_impl = None

    def box_nms(self, data, **kw):
        assert self.box_nms_impl is not None, "box_nms implementation not supplied"
        return NDArray(self.box_nms_impl(data.a, **kw))


class _F:
    """the `F` namespace (mxnet.nd) as yolo3.py uses it"""
    def __init__(self):
        self.contrib = _Contrib()

    @staticmethod
    def sigmoid(x):
        # mshadow_op::sigmoid: 1.0f / (1.0f + expf(-a))
        return NDArray(f32(1.0) / (f32(1.0) + _expf(-x.a)))

    @staticmethod
    def exp(x):
        return NDArray(_expf(x.a))

    @staticmethod
    def broadcast_add(a, b):
        return NDArray(a.a + b.a)

    @staticmethod
    def broadcast_mul(a, b):
        return NDArray(a.a * b.a)

    @staticmethod
    def concat(*arrs, dim=1):
        return NDArray(np.concatenate([x.a for x in arrs], axis=dim))

    @staticmethod
    def tile(x, reps):
        return NDArray(np.tile(x.a, reps))

    @staticmethod
    def transpose(x, axes=None):
        return x.transpose(axes=axes) if axes is not None else x.transpose()

    @staticmethod
    def arange(start, stop=None, step=1.0):
        if stop is None:
            start, stop = 0, start
        return NDArray(np.arange(start, stop, step, dtype=f32))

    @staticmethod
    def reshape(x, shape):
        return x.reshape(shape)

    @staticmethod
    def slice_like(x, like, axes=None):
        sl = [slice(None)] * x.a.ndim
        for ax in (axes if axes is not None else range(x.a.ndim)):
            sl[ax] = slice(0, like.a.shape[ax])
        return NDArray(x.a[tuple(sl)])

    @staticmethod
    def zeros_like(x):
        return NDArray(np.zeros_like(x.a))

    @staticmethod
    def squeeze(x, axis=None):
        return NDArray(np.squeeze(x.a, axis=axis))

    @staticmethod
    def max(x, axis=None, keepdims=False):
        return NDArray(np.max(x.a, axis=axis, keepdims=keepdims))

    @staticmethod
    def mean(x, axis=None, keepdims=False):
        return NDArray(np.mean(x.a, axis=axis, keepdims=keepdims, dtype=f32))


F = _F()


# ---------------------------------------------------------------- gluon stand-ins
class _Const:
    def __init__(self, value):
        self.value = NDArray(np.asarray(value))


class _Params:
    def __init__(self):
        self.consts = {}

    def get_constant(self, name, value):
        c = _Const(value)
        self.consts[name] = c
        return c


class _Scope:
    def __enter__(self): return self
    def __exit__(self, *a): return False


class HybridBlock:
    """gluon.HybridBlock as far as `__call__ -> hybrid_forward(F, x, **registered constants)` goes"""
    def __init__(self, prefix=None, params=None, **kwargs):
        object.__setattr__(self, "_reg_consts", {})
        object.__setattr__(self, "params", _Params())

    def __setattr__(self, k, v):
        if isinstance(v, _Const):
            self._reg_consts[k] = v
        object.__setattr__(self, k, v)

    def name_scope(self):
        return _Scope()

    def _clear_cached_op(self):
        pass

    def __call__(self, *args):
        kw = {k: c.value for k, c in self._reg_consts.items()}
        return self.hybrid_forward(F, *args, **kw)

    def hybrid_forward(self, F, *a, **k):
        raise NotImplementedError


class HybridSequential(HybridBlock):
    def __init__(self, *a, **k):
        super().__init__()
        object.__setattr__(self, "_children_list", [])

    def add(self, *blocks):
        self._children_list.extend(blocks)

    def __iter__(self): return iter(self._children_list)
    def __len__(self): return len(self._children_list)
    def __getitem__(self, i): return self._children_list[i]

    def hybrid_forward(self, F, x):
        for b in self._children_list:
            x = b(x)
        return x


class Identity(HybridBlock):
    """stands in for every learned layer that is NOT on the path being pinned (the 1x1 prediction
    conv, the transition convs): the head maps are fed in as the 'tip'."""
    def __init__(self, *a, **k):
        super().__init__()

    def hybrid_forward(self, F, x):
        return x


class _Auto(types.ModuleType):
    """module whose unknown attributes are inert placeholder classes (for names only imported)"""
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (Identity,), {})
        setattr(self, name, cls)
        return cls


def install():
    """Put stub `mxnet` / `gluoncv` packages into sys.modules; returns the `autograd` stub."""
    def mod(name):
        m = _Auto(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    mx = mod("mxnet")
    gluon = mod("mxnet.gluon")
    nn = mod("mxnet.gluon.nn")
    autograd = mod("mxnet.autograd")
    nd = mod("mxnet.nd")
    mx.gluon, mx.autograd, mx.nd, gluon.nn = gluon, autograd, nd, nn
    mod("mxnet.gluon.contrib"); mod("mxnet.gluon.contrib.nn"); mod("mxnet.gluon.rnn"); mod("mxnet.initializer")
    mod("mxnet.gluon.contrib.rnn"); mod("mxnet.gluon.contrib.rnn.conv_rnn_cell")
    gluon.HybridBlock = HybridBlock
    gluon.Block = HybridBlock
    nn.HybridBlock = HybridBlock
    nn.HybridSequential = HybridSequential
    nn.Conv2D = type("Conv2D", (Identity,), {})
    autograd.is_training = lambda: False
    autograd.is_recording = lambda: False
    for n in ("gluoncv", "gluoncv.loss", "gluoncv.nn", "gluoncv.nn.bbox", "gluoncv.model_zoo",
              "gluoncv.utils", "gluoncv.data", "gluoncv.nn.coder", "gluoncv.nn.feature"):
        mod(n)
    sys.modules["gluoncv"].loss = sys.modules["gluoncv.loss"]
    sys.modules["gluoncv"].nn = sys.modules["gluoncv.nn"]
    return autograd
