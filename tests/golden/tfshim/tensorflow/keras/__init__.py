"""tf.keras subset: eager `Input` / `Model` so the reference's functional-API builders run on concrete arrays.

`Input(...)` hands back the next concrete array queued with `feed(...)`, so "building" a model with the reference's
builder (sagan/models/generator.py:14-37, sagan/models/discriminator.py:13-36) IS one forward pass through exactly
the layers, in exactly the order, the builder wires up; `Model(inputs, outputs)` just keeps the result and the layers
that were created (in creation order) for the fixture script to read.
"""
from . import layers  # noqa: F401
from .._core import Tensor

_feed = []


def feed(*arrays):
    _feed[:] = list(arrays)
    layers.created[:] = []


def Input(shape=None, batch_size=None, dtype=None, name=None):
    assert _feed, "shim: no concrete array queued for Input(name=%r)" % (name,)
    t = Tensor(_feed.pop(0))
    want = ([batch_size] if batch_size is not None else [t.shape[0]]) + [int(s) for s in shape]
    assert list(t.shape) == want, (name, list(t.shape), want)
    return t


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs = inputs, outputs
        self.layers = list(layers.created)


class Sequential:                     # imported by the legacy builders (models/generator.py:2), never used by them
    pass


class optimizers:                     # likewise
    pass
