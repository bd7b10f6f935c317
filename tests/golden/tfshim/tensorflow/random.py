"""tf.random subset: a seeded numpy generator the fixture script re-seeds before every case."""
import numpy as np

from ._core import Tensor

_rng = [np.random.Generator(np.random.PCG64(0))]


def set_seed(seed):
    _rng[0] = np.random.Generator(np.random.PCG64(seed))


def normal(shape, mean=0.0, stddev=1.0):
    return Tensor(mean + stddev * _rng[0].standard_normal([int(s) for s in shape]))


def uniform(shape, minval=0.0, maxval=1.0, dtype=None):
    if dtype in ("int32", "int64"):
        return Tensor(_rng[0].integers(minval, maxval, [int(s) for s in shape]).astype(dtype))
    return Tensor(_rng[0].uniform(minval, maxval, [int(s) for s in shape]))
