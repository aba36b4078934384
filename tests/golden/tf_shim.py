"""A tiny eager stand-in for the `tensorflow` 1.x API surface the reference's hot path touches.

PURPOSE: lets tests/golden/make_golden.py import the UNMODIFIED reference modules
(/root/reference/Models/*.py, Utils.py) and execute their own graph-building code, so the wiring of
the network (layer order, crops, concat order, gather interleave, output algebra) in the committed
fixtures comes from the reference's source, not from our restatement.  Leaf ops are implemented on
torch-CPU tensors following the TF-1.8 op definitions; torch autograd then differentiates the graph
the reference code built (TF would have used tf.gradients).

Used ONLY by make_golden.py, in the build container.  Never shipped, never imported by tests that run
on the GPU box.
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


class Shape(list):
    def as_list(self):
        return list(self)


class T:
    """Tensor wrapper with the few TF tensor methods/operators the reference uses."""

    def __init__(self, t, name=None):
        self.t = t
        self.name = name

    def get_shape(self):
        return Shape(self.t.shape)

    def __getitem__(self, idx):  # noqa: E301
        return T(self.t[idx])

    def _v(self, o):
        return o.t if isinstance(o, T) else o

    def __add__(self, o): return T(self.t + self._v(o))
    def __radd__(self, o): return T(self._v(o) + self.t)
    def __sub__(self, o): return T(self.t - self._v(o))
    def __rsub__(self, o): return T(self._v(o) - self.t)
    def __mul__(self, o): return T(self.t * self._v(o))
    def __rmul__(self, o): return T(self._v(o) * self.t)
    def __truediv__(self, o): return T(self.t / self._v(o))


class _State:
    def __init__(self):
        self.reset()

    def reset(self, seed=1337, dtype=torch.float64):
        self.vars = {}          # name -> torch leaf (creation order preserved)
        self.scope = []
        self.counters = {}
        self.rng = np.random.default_rng(seed)
        self.dtype = dtype
        self.preset = None      # optional dict name -> np array to use instead of random init
        self.uniform_queue = []  # values tf.random_uniform returns, in call order (Utils.random_amplify)


STATE = _State()


def _unique(base):
    scope = "/".join(STATE.scope)
    key = scope + "/" + base
    n = STATE.counters.get(key, 0)
    STATE.counters[key] = n + 1
    return key if n == 0 else "%s_%d" % (key, n)


def _make_var(name, shape, glorot_fans=None, zeros=False):
    if name in STATE.vars:
        return STATE.vars[name]
    if STATE.preset is not None and name in STATE.preset:
        arr = np.asarray(STATE.preset[name], dtype=np.float64)
        assert tuple(arr.shape) == tuple(shape), (name, arr.shape, shape)
    elif zeros:
        arr = np.zeros(shape)
    else:
        fan_in, fan_out = glorot_fans
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        arr = STATE.rng.uniform(-lim, lim, size=shape)
    t = torch.tensor(arr, dtype=STATE.dtype, requires_grad=True)
    STATE.vars[name] = t
    return t


class _VarScope:
    def __init__(self, name, reuse=None):
        self.name = name

    def __enter__(self):
        STATE.scope.append(self.name)
        return self

    def __exit__(self, *a):
        STATE.scope.pop()


def _same_pad(k):
    left = (k - 1) // 2
    return left, k - 1 - left


def _conv1d_layer(inputs, filters, kernel_size, strides=1, activation=None, padding="valid", **kw):
    assert strides == 1
    x = inputs.t
    cin = x.shape[2]
    lname = _unique("conv1d")
    kern = _make_var(lname + "/kernel", (kernel_size, cin, filters),
                     glorot_fans=(kernel_size * cin, kernel_size * filters))
    bias = _make_var(lname + "/bias", (filters,), zeros=True)
    xt = x.transpose(1, 2)
    if padding.lower() == "same":
        xt = F.pad(xt, _same_pad(kernel_size))
    y = F.conv1d(xt, kern.permute(2, 1, 0), bias).transpose(1, 2)
    out = T(y)
    return activation(out) if activation is not None else out


def _get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    full = "/".join(STATE.scope + [name])
    n = int(np.prod(shape))
    return T(_make_var(full, tuple(shape), glorot_fans=(n, n)), name=full)


def _resize_bilinear(images, size, align_corners=False):
    """TF-1.8 ResizeBilinear on NHWC (legacy: no half-pixel centres)."""
    x = images.t
    B, H, W, C = x.shape
    oh, ow = int(size[0]), int(size[1])
    assert H == 1 and oh == 1

    def scale(i, o):
        return (i - 1) / float(o - 1) if (align_corners and o > 1) else i / float(o)

    ws = scale(W, ow)
    cols = []
    for xo in range(ow):
        src = xo * ws
        lo = int(math.floor(src))
        hi = min(lo + 1, W - 1)
        frac = src - lo
        left = x[:, :, lo, :]
        right = x[:, :, hi, :]
        cols.append(left + (right - left) * frac)
    return T(torch.stack(cols, dim=2))


def _conv2d(inp, filt, strides, padding):
    x = inp.t                      # NHWC
    w = filt.t                     # [kh, kw, cin, cout]
    kh, kw = w.shape[0], w.shape[1]
    xt = x.permute(0, 3, 1, 2)     # NCHW
    if padding.upper() == "SAME":
        pl, pr = _same_pad(kw)
        pt, pb = _same_pad(kh)
        xt = F.pad(xt, (pl, pr, pt, pb))
    y = F.conv2d(xt, w.permute(3, 2, 0, 1))
    return T(y.permute(0, 2, 3, 1))


def _concat(values, axis):
    return T(torch.cat([v.t for v in values], dim=axis))


def _gather(params, indices):
    return T(params.t[torch.as_tensor(list(indices), dtype=torch.long)])


def install():
    """Register fake `tensorflow` and `librosa` modules in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.Tensor = T
    tf.variable_scope = _VarScope
    tf.get_variable = _get_variable
    tf.expand_dims = lambda x, axis: T(x.t.unsqueeze(axis))
    tf.squeeze = lambda x, axis=None: T(x.t.squeeze(axis))
    tf.concat = _concat
    tf.transpose = lambda x, perm: T(x.t.permute(*perm))
    tf.gather = _gather
    tf.diag = lambda x: T(torch.diag(x.t))
    tf.tanh = lambda x: T(torch.tanh(x.t))
    # tf.maximum(a, b): gradient goes to `a` where a >= b (MaximumGrad) - made explicit here
    def _maximum(a, b):
        a_t = a.t if isinstance(a, T) else torch.as_tensor(a, dtype=b.t.dtype)
        b_t = b.t if isinstance(b, T) else torch.as_tensor(b, dtype=a_t.dtype)
        return T(torch.where(a_t >= b_t, a_t + 0 * b_t, b_t + 0 * a_t))

    def _minimum(a, b):
        a_t = a.t if isinstance(a, T) else torch.as_tensor(a, dtype=b.t.dtype)
        b_t = b.t if isinstance(b, T) else torch.as_tensor(b, dtype=a_t.dtype)
        return T(torch.where(a_t <= b_t, a_t + 0 * b_t, b_t + 0 * a_t))
    tf.maximum = _maximum
    tf.minimum = _minimum
    layers = types.ModuleType("tensorflow.layers")
    layers.conv1d = _conv1d_layer
    tf.layers = layers
    image = types.ModuleType("tensorflow.image")
    image.resize_bilinear = _resize_bilinear
    tf.image = image
    nn = types.ModuleType("tensorflow.nn")
    nn.sigmoid = lambda x: T(torch.sigmoid(x.t))
    nn.conv2d = _conv2d
    tf.nn = nn
    tf.trainable_variables = lambda: []
    # data pipeline ops used by Utils.random_amplify (Utils.py:33-35): the "random" gains are popped from a queue the
    # fixture generator fills (TF's stateful RNG cannot be reproduced; the values are inputs of the fixture instead)
    tf.random_uniform = lambda shape, minval=0.0, maxval=1.0, **kw: T(torch.as_tensor(STATE.uniform_queue.pop(0), dtype=STATE.dtype))

    def _add_n(vals):
        acc = vals[0].t if isinstance(vals[0], T) else torch.as_tensor(vals[0])
        for v in vals[1:]:
            acc = acc + (v.t if isinstance(v, T) else torch.as_tensor(v))      # left to right, like the kernel
        return T(acc)
    tf.add_n = _add_n
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.layers"] = layers
    sys.modules["tensorflow.image"] = image
    sys.modules["tensorflow.nn"] = nn
    if "librosa" not in sys.modules:
        sys.modules["librosa"] = types.ModuleType("librosa")   # Utils.py:3 imports it at module level
    return tf
