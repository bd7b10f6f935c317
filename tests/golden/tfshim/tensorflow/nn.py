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
