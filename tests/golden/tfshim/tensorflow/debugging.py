"""tf.debugging subset."""
import numpy as np

from ._core import raw


def assert_equal(x, y, message=None):
    assert np.array_equal(raw(x), raw(y)), message
    return None
