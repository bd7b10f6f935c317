"""Tensor / Variable wrappers of the numpy stand-in (see the package docstring)."""
import numpy as np


class TensorShape(tuple):
    """`Tensor.shape`: indexable like a tuple, with the `as_list()` the reference calls (layers.py:80,94)."""

    def as_list(self):
        return [int(s) for s in self]


def raw(x):
    if isinstance(x, Tensor):
        return x.a
    return np.asarray(x)


def convert(x):
    return x if isinstance(x, Tensor) else Tensor(x)


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        a = np.asarray(a)
        if a.dtype.kind == "f":
            a = a.astype(np.float64)
        self.a = a

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    @property
    def dtype(self):
        return str(self.a.dtype)

    def numpy(self):
        return self.a

    def __add__(self, o):
        return Tensor(self.a + raw(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Tensor(self.a - raw(o))

    def __rsub__(self, o):
        return Tensor(raw(o) - self.a)

    def __mul__(self, o):
        return Tensor(self.a * raw(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return Tensor(self.a / raw(o))

    def __rtruediv__(self, o):
        return Tensor(raw(o) / self.a)

    def __neg__(self):
        return Tensor(-self.a)

    def __getitem__(self, k):
        return Tensor(self.a[k])

    def __repr__(self):
        return "shim.Tensor(shape=%s)" % (tuple(self.a.shape),)


class Variable(Tensor):
    def __init__(self, a, name=None, trainable=True):
        super().__init__(a)
        self.name = name
        self.trainable = trainable

    def assign(self, value):
        v = raw(value)
        assert tuple(v.shape) == tuple(self.a.shape), (v.shape, self.a.shape)
        self.a = np.array(v, dtype=self.a.dtype)
        return self
