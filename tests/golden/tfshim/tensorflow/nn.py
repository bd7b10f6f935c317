"""tf.nn subset."""
import numpy as np

from ._core import Tensor, raw


def softmax(logits, axis=-1):
    """tf.nn.softmax: exp(x - max) / sum over `axis` (default the last one, layers.py:109)."""
    x = raw(logits)
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return Tensor(e / np.sum(e, axis=axis, keepdims=True))


def relu(x):
    return Tensor(np.maximum(raw(x), 0.0))


def leaky_relu(x, alpha=0.2):
    a = raw(x)
    return Tensor(np.where(a >= 0, a, alpha * a))


def l2_normalize(x, axis=None, epsilon=1e-12):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2, axis), epsilon))."""
    a = raw(x)
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    ss = np.sum(np.square(a), axis=ax, keepdims=True)
    return Tensor(a / np.sqrt(np.maximum(ss, epsilon)))


def moments(x, axes):
    """tf.nn.moments: mean and (biased) variance over `axes`."""
    a = raw(x)
    ax = tuple(axes)
    return Tensor(a.mean(axis=ax)), Tensor(a.var(axis=ax))
