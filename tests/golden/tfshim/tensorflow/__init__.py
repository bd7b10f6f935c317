"""A numpy-backed stand-in for the handful of `tensorflow` symbols the reference's hot-path files touch.

TEST INFRASTRUCTURE ONLY.  TensorFlow cannot be installed in this environment (no network, not in the wheelhouse),
so the reference's own Python (`/root/reference/layers.py`, the model builders under `/root/reference/sagan/models/`
and `/root/reference/models/`, `/root/reference/sagan/layers.py`, `/root/reference/sagan/dataset.py`, the loss functions
of `/root/reference/sagan/main.py`) cannot run as shipped.  This package lets those files be
IMPORTED AND EXECUTED UNMODIFIED by `tests/golden/make_reference_vectors.py`: every line of control flow, every
reshape, transposition, normalisation and reduction executed is the reference's; only the primitive array operations
underneath (`tf.matmul`, `tf.norm`, `tf.reshape`, `tf.nn.softmax`, Keras `Conv2D` ...) are supplied here, in float64,
with the semantics TensorFlow documents for them.  It is eager: there is no graph, `tf.function` is the identity.

Nothing outside `tests/golden/make_reference_vectors.py` imports this package; it is never on the product path and
does not travel into the GPU tests (they read the `.npz` fixtures the script wrote).
"""
import numpy as _np

from . import _core
from ._core import Tensor, TensorShape, Variable, convert as convert_to_tensor  # noqa: F401
from . import keras  # noqa: F401
from . import nn  # noqa: F401
from . import random  # noqa: F401
from . import math  # noqa: F401
from . import linalg  # noqa: F401
from . import debugging  # noqa: F401
from . import dtypes  # noqa: F401
from . import compat  # noqa: F401
from . import io  # noqa: F401
from . import data  # noqa: F401

string = "string"
uint8 = "uint8"

float32 = "float32"
float64 = "float64"
int32 = "int32"
int64 = "int64"

__version__ = "0.0-numpy-shim"


def function(fn=None, **_kw):
    """tf.function: eager here, so the decorated Python runs as written (layers.py:50)."""
    if fn is None:
        return lambda f: f
    return fn


def _a(x):
    return _core.raw(x)


def reshape(t, shape):
    return Tensor(_np.reshape(_a(t), [int(s) for s in shape]), keep_dtype=getattr(t, "keep", False))


def transpose(t, perm=None):
    return Tensor(_np.transpose(_a(t), perm))


def matmul(a, b):
    """Batched matrix product over the last two axes with broadcasting of the leading ones (tf.matmul)."""
    return Tensor(_np.matmul(_a(a), _a(b)))


def norm(t):
    """tf.norm with default arguments: the Euclidean norm of ALL elements (Frobenius for matrices)."""
    return Tensor(_np.sqrt(_np.sum(_np.square(_a(t)))))


def reduce_sum(t, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return Tensor(_np.sum(_a(t), axis=axis, keepdims=keepdims))


def reduce_mean(t, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return Tensor(_np.mean(_a(t), axis=axis, keepdims=keepdims))


def one_hot(indices, depth):
    idx = _a(indices).astype(_np.int64)
    return Tensor(_np.eye(int(depth), dtype=_np.float64)[idx])


def ones_like(t):
    return Tensor(_np.ones_like(_a(t)))


def zeros_like(t):
    return Tensor(_np.zeros_like(_a(t)))


def tanh(t):
    return Tensor(_np.tanh(_a(t)))


def tile(t, multiples):
    return Tensor(_np.tile(_a(t), [int(m) for m in multiples]))


def identity(t):
    return Tensor(_np.array(_a(t)))


def cast(t, dtype):
    """tf.cast; a cast to float32 yields a tensor whose arithmetic STAYS float32 (numpy float32 with Python-float
    scalars, which is what TensorFlow does with `x * (2. / 255) - 1.` on a float32 tensor)."""
    return Tensor(_a(t).astype(dtype), keep_dtype=(dtype == "float32"))


def cond(pred, true_fn, false_fn):
    """Eager tf.cond: exactly one branch runs."""
    return true_fn() if bool(_a(pred)) else false_fn()


class _NullContext:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


name_scope = _NullContext


def control_dependencies(ops):
    """Eager: the operations in `ops` have already run by the time the list exists."""
    return _NullContext()


class CriticalSection:
    def __init__(self, name=None):
        self.name = name

    def execute(self, fn):
        return fn()


class VariableSynchronization:
    AUTO = "auto"


def constant(value, dtype=None):
    return Tensor(_np.asarray(value, dtype=_np.float64 if dtype in (None, "float32", "float64") else dtype))


class distribute:
    class ReduceOp:
        SUM = "sum"


class GradientTape:
    """No automatic differentiation here: the tape records WHAT the reference asks to differentiate (the target scalar,
    the variable list) and which model calls happened while it was open; `gradient` hands back one opaque token per
    variable.  Used to pin the step schedule and the loss scaling of sagan/main.py:171-211, not the gradients."""
    log = []            # shared event log, reset by the fixture script
    depth = 0

    def __enter__(self):
        GradientTape.depth += 1
        GradientTape.log.append(("tape_open",))
        return self

    def __exit__(self, *exc):
        GradientTape.depth -= 1
        GradientTape.log.append(("tape_close",))
        return False

    def gradient(self, target, sources):
        GradientTape.log.append(("gradient", float(_a(target)), tuple(getattr(v, "name", "?") for v in sources)))
        return [("grad", getattr(v, "name", "?")) for v in sources]
