"""Tensor / Variable wrappers of the numpy stand-in (see the package docstring)."""
import numpy as np


class TensorShape(tuple):
    """`Tensor.shape`: indexable like a tuple, with the `as_list()` the reference calls (layers.py:80,94)."""

    def as_list(self):
        return [None if s is None else int(s) for s in self]

    @property
    def rank(self):
        return len(self)

    def __getitem__(self, k):
        r = tuple.__getitem__(self, k)
        return TensorShape(r) if isinstance(k, slice) else r

    def __add__(self, o):
        return TensorShape(tuple(self) + tuple(o))

    def __radd__(self, o):
        return TensorShape(tuple(o) + tuple(self))


def raw(x):
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (bool, int, float)):
        return x                      # Python scalars stay "weak": float32 tensor * 2. / 255 is float32, as in TF
    return np.asarray(x)


def convert(x):
    return x if isinstance(x, Tensor) else Tensor(x)


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a, keep_dtype=False):
        """Floating-point data is promoted to float64 unless `keep_dtype` (set by tf.cast(x, tf.float32): the record
        decode of sagan/dataset.py:34 is float32 arithmetic and is pinned bit for bit)."""
        a = np.asarray(a)
        if a.dtype.kind == "f" and not keep_dtype:
            a = a.astype(np.float64)
        self.a = a
        self.keep = bool(keep_dtype)

    def _w(self, arr):
        return Tensor(arr, keep_dtype=getattr(self, "keep", False))

    def set_shape(self, shape):
        assert self.a.size == int(np.prod([int(s) for s in (shape if isinstance(shape, (list, tuple)) else [shape])]))

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    @property
    def dtype(self):
        return str(self.a.dtype)

    def numpy(self):
        return self.a

    def __add__(self, o):
        return self._w(self.a + raw(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self._w(self.a - raw(o))

    def __rsub__(self, o):
        return self._w(raw(o) - self.a)

    def __mul__(self, o):
        return self._w(self.a * raw(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._w(self.a / raw(o))

    def __rtruediv__(self, o):
        return self._w(raw(o) / self.a)

    def __neg__(self):
        return self._w(-self.a)

    def __getitem__(self, k):
        return self._w(self.a[k])

    def __repr__(self):
        return "shim.Tensor(shape=%s)" % (tuple(self.a.shape),)


class Variable(Tensor):
    def __init__(self, a, name=None, trainable=True):
        super().__init__(a)
        self.name = name
        self.trainable = trainable

    def assign(self, value):
        v = np.asarray(raw(value))
        assert tuple(v.shape) == tuple(self.a.shape), (v.shape, self.a.shape)
        self.a = np.array(v, dtype=self.a.dtype)
        return self
