"""tf.linalg subset."""
import numpy as np

from ._core import Tensor, raw


def norm(t, axis=None):
    return Tensor(np.sqrt(np.sum(np.square(raw(t)), axis=axis)))
