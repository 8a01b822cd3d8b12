"""TF1 API shim -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

TensorFlow 1.x cannot be installed in this image (Python 3.12, no network), so the
reference sources under /root/reference cannot run on their own runtime.  This
module implements just enough of the TF1 *eager-equivalent* API surface, backed by
torch CPU tensors, that the reference's own hot-path source files (utils.py,
utils_lr.py and selected functions of my_losses.py / Demon_Data_loader.py) can be
imported and executed UNMODIFIED from where they lie.  tests/golden/make_golden.py
uses it to produce the golden vectors that pin oracle/vsl_oracle.py.

What is pinned: the reference's op graph (which ops, in which order, with which
constants).  What is NOT pinned: the numerics of the TF C++ kernels themselves.
Where a TF kernel's arithmetic order matters the shim states the convention:

* tf.matmul with inner dim <= 8 is evaluated as a sequential sum over k of
  separate multiplies and adds (no FMA).  Stock TF1 CPU wheels were SSE-only
  Eigen builds (the well-known "AVX2 FMA" start-up warning), so no contraction.
* tf.linspace follows TF1's LinSpace kernel: start + step * i, step =
  (stop - start) / (num - 1), evaluated in the working dtype.
* tf.matrix_inverse follows Eigen PartialPivLU: row-pivoted LU then column-wise
  forward/back substitution.
* tf.add_n sums left to right.
* tf.image.resize_area supports integer shrink factors only (all the reference
  uses).  Order of TF's ResizeArea kernel (ComputePatchSum) when every overlap
  weight is 1: per contributing row the fx values are summed left to right, the
  fy row sums are accumulated top to bottom, the total is scaled by 1/(fy*fx).
* tf.gather raises on out-of-range indices like the TF CPU kernel.

`set_float(torch.float64)` makes 'float32' mean float64 so the same reference
source can also be evaluated in double precision (used for gradient ground truth).
"""
import contextlib

import numpy as np
import torch

_FLOAT = torch.float32


def set_float(dtype):
    """Select what the reference's 'float32' means (float32 or float64)."""
    global _FLOAT
    assert dtype in (torch.float32, torch.float64)
    _FLOAT = dtype


def get_float():
    return _FLOAT


class _Shape(list):
    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    """torch.Tensor with the handful of tf.Tensor methods the reference calls."""

    def get_shape(self):
        return _Shape(int(d) for d in self.shape)

    # TF tensors are immutable: `a /= b` rebinds, it never writes through a view.
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o


def _T(x):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, torch.Tensor):
        return x.as_subclass(Tensor)
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if t.is_floating_point():
            t = t.to(_FLOAT)
        return t.as_subclass(Tensor)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=_FLOAT).as_subclass(Tensor)
    if isinstance(x, (int, np.integer, bool)):
        return torch.tensor(int(x), dtype=torch.int32).as_subclass(Tensor)
    if isinstance(x, (list, tuple)):
        return _T(np.asarray(x))
    raise TypeError(type(x))


class _DType(object):
    def __init__(self, name):
        self.name = name


float32 = _DType('float32')
float64 = _DType('float64')
int32 = _DType('int32')
int64 = _DType('int64')


def _dtype(d):
    name = d.name if isinstance(d, _DType) else str(d)
    if name in ('float32', 'float'):
        return _FLOAT
    if name == 'float64':
        return torch.float64
    if name == 'int32':
        return torch.int32
    if name == 'int64':
        return torch.int64
    raise TypeError(d)


def _ishape(shape):
    if isinstance(shape, torch.Tensor):
        shape = shape.tolist()
    return [int(s) for s in shape]


# ---------------------------------------------------------------- creation
def convert_to_tensor(x, dtype=None):
    t = _T(x)
    return t if dtype is None else cast(t, dtype)


def constant(value, dtype=None, shape=None):
    arr = np.asarray(value)
    if dtype is None:
        dt = _FLOAT if arr.dtype.kind == 'f' else torch.int32
    else:
        dt = _dtype(dtype)
    t = torch.as_tensor(arr).to(dt)
    if shape is not None:
        shape = _ishape(shape)
        n = int(np.prod(shape))
        if t.numel() == 1 and n != 1:
            t = t.reshape(1).expand(n)
        t = t.reshape(shape)
    return _T(t.contiguous())


def zeros(shape, dtype=float32):
    return _T(torch.zeros(_ishape(shape), dtype=_dtype(dtype)))


def ones(shape, dtype=float32):
    return _T(torch.ones(_ishape(shape), dtype=_dtype(dtype)))


def zeros_like(x):
    return _T(torch.zeros_like(_T(x)))


def ones_like(x):
    return _T(torch.ones_like(_T(x)))


def eye(n, batch_shape=None, dtype=float32):
    e = torch.eye(n, dtype=_dtype(dtype))
    if batch_shape is not None:
        e = e.expand(*_ishape(batch_shape), n, n).contiguous()
    return _T(e)


def range(*args):  # noqa: A001 - mirrors tf.range
    return _T(torch.arange(*[int(a) for a in args], dtype=torch.int32))


def linspace(start, stop, num):
    num = int(num)
    start_t = torch.tensor(start, dtype=_FLOAT)
    stop_t = torch.tensor(stop, dtype=_FLOAT)
    step = (stop_t - start_t) / torch.tensor(num - 1, dtype=_FLOAT)
    i = torch.arange(num).to(_FLOAT)
    return _T(start_t + step * i)


# ---------------------------------------------------------------- shape ops
def shape(x):
    return _Shape(int(d) for d in _T(x).shape)


def reshape(x, shp):
    return _T(x).reshape(_ishape(shp))


def expand_dims(x, axis):
    return _T(x).unsqueeze(axis)


def squeeze(x, axis=None):
    x = _T(x)
    if axis is None:
        return x.squeeze()
    for a in sorted(_ishape(axis if isinstance(axis, (list, tuple)) else [axis]), reverse=True):
        assert x.shape[a] == 1
        x = x.squeeze(a)
    return x


def transpose(x, perm=None):
    x = _T(x)
    if perm is None:
        perm = list(reversed(list(np.arange(x.dim()))))
    return x.permute(*_ishape(perm)).contiguous()


def concat(values, axis):
    return _T(torch.cat([_T(v) for v in values], dim=axis))


def stack(values, axis=0):
    if all(isinstance(v, (int, np.integer)) for v in values):
        return [int(v) for v in values]  # used by the reference only as a shape
    return _T(torch.stack([_T(v) for v in values], dim=axis))


def tile(x, multiples):
    return _T(x).repeat(*_ishape(multiples))


def slice(x, begin, size):  # noqa: A001 - mirrors tf.slice
    x = _T(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        e = x.shape[d] if s == -1 else b + s
        idx.append(np.s_[b:e])
    return x[tuple(idx)].clone()


def split(x, num_or_size_splits, axis=0):
    x = _T(x)
    if isinstance(num_or_size_splits, int):
        return [p.clone() for p in torch.chunk(x, num_or_size_splits, dim=axis)]
    return [p.clone() for p in torch.split(x, list(num_or_size_splits), dim=axis)]


def cast(x, dtype):
    if isinstance(x, (list, _Shape)) or (isinstance(x, (int, float)) and not isinstance(x, bool)):
        x = torch.tensor(x)
    dt = _dtype(dtype)
    x = _T(x)
    if dt in (torch.int32, torch.int64) and x.is_floating_point():
        return torch.trunc(x).to(dt)  # C-style float->int
    return x.to(dt)


def to_float(x):
    return cast(x, float32)


# ---------------------------------------------------------------- math
def _seq_matmul(a, b):
    K = a.shape[-1]
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for k in np.arange(1, K):
        acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return acc


def matmul(a, b):
    a, b = _T(a), _T(b)
    if a.shape[-1] <= 8:
        return _T(_seq_matmul(a, b))
    return _T(torch.matmul(a, b))


def lu_inverse(m):
    """Eigen::PartialPivLU::inverse() on one small square matrix (numpy, working dtype)."""
    n = m.shape[0]
    dt = m.dtype.type
    lu = m.copy()
    perm = list(np.arange(n))
    for k in np.arange(n):
        p = k + int(np.argmax(np.abs(lu[k:, k])))
        if p != k:
            lu[[k, p], :] = lu[[p, k], :]
            perm[k], perm[p] = perm[p], perm[k]
        for i in np.arange(k + 1, n):
            lu[i, k] = dt(lu[i, k] / lu[k, k])
            for j in np.arange(k + 1, n):
                lu[i, j] = dt(lu[i, j] - dt(lu[i, k] * lu[k, j]))
    inv = np.zeros_like(m)
    for c in np.arange(n):
        y = np.zeros(n, dtype=m.dtype)
        for i in np.arange(n):  # forward: L y = P e_c  (unit diagonal)
            acc = dt(1.0) if perm[i] == c else dt(0.0)
            for j in np.arange(i):
                acc = dt(acc - dt(lu[i, j] * y[j]))
            y[i] = acc
        for i in np.arange(n - 1, -1, -1):  # backward: U x = y
            acc = y[i]
            for j in np.arange(i + 1, n):
                acc = dt(acc - dt(lu[i, j] * inv[j, c]))
            inv[i, c] = dt(acc / lu[i, i])
    return inv


def matrix_inverse(m):
    m = _T(m)
    if m.requires_grad:
        return _T(torch.linalg.inv(m))
    a = m.detach().numpy().reshape(-1, m.shape[-2], m.shape[-1])
    out = np.stack([lu_inverse(a[i]) for i in np.arange(a.shape[0])]).reshape(tuple(m.shape))
    return _T(torch.from_numpy(out))


def sin(x):
    return torch.sin(_T(x))


def cos(x):
    return torch.cos(_T(x))


def floor(x):
    return torch.floor(_T(x))


def abs(x):  # noqa: A001
    return torch.abs(_T(x))


def exp(x):
    return torch.exp(_T(x))


def log(x):
    return torch.log(_T(x))


def sqrt(x):
    return torch.sqrt(_T(x))


def square(x):
    x = _T(x)
    return x * x


def multiply(a, b):
    return _T(a) * _T(b)


def add_n(xs):
    acc = _T(xs[0])
    for x in xs[1:]:
        acc = acc + _T(x)
    return acc


def equal(a, b):
    return torch.eq(_T(a), _T(b))


def where(c, a, b):
    return torch.where(c, _T(a), _T(b))


def clip_by_value(x, lo, hi):
    x = _T(x)
    lo = _T(lo).to(x.dtype) if not isinstance(lo, (int, float)) else torch.tensor(lo, dtype=x.dtype)
    hi = _T(hi).to(x.dtype) if not isinstance(hi, (int, float)) else torch.tensor(hi, dtype=x.dtype)
    # TF1: minimum(maximum(x, lo), hi); gradient passes where lo <= x <= hi
    return _T(torch.minimum(torch.maximum(x, lo), hi))


def norm(x, axis=None):
    x = _T(x)
    return torch.sqrt(torch.sum(x * x, dim=axis))


def reduce_mean(x, axis=None):
    x = _T(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis)


def reduce_sum(x, axis=None):
    x = _T(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


def gather(params, indices):
    params, indices = _T(params), _T(indices).long()
    n = params.shape[0]
    if indices.numel() and (int(indices.min()) < 0 or int(indices.max()) >= n):
        raise IndexError('tf.gather: index out of range [0, %d)' % n)
    return params[indices]


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield


class _NN(object):
    @staticmethod
    def softmax(logits):
        return torch.softmax(_T(logits), dim=-1)

    @staticmethod
    def softmax_cross_entropy_with_logits(labels=None, logits=None):
        return -torch.sum(_T(labels) * torch.log_softmax(_T(logits), dim=-1), dim=-1)


nn = _NN()


class _Image(object):
    @staticmethod
    def resize_area(images, size):
        x = _T(images)
        B, H, W, C = x.shape
        oh, ow = int(size[0]), int(size[1])
        assert H % oh == 0 and W % ow == 0, 'shim supports integer shrink factors only'
        fy, fx = H // oh, W // ow
        blocks = x.reshape(B, oh, fy, ow, fx, C)
        acc = None
        for dy in np.arange(fy):
            row = blocks[:, :, dy, :, 0, :]
            for dx in np.arange(1, fx):
                row = row + blocks[:, :, dy, :, dx, :]
            acc = row if acc is None else acc + row
        scale = torch.tensor(1.0, dtype=_FLOAT) / torch.tensor(float(fy * fx), dtype=_FLOAT)
        return _T(acc * scale)


image = _Image()
