"""tf.io subset for the record reader of sagan/dataset.py:12-40.  A serialized tf.train.Example is stood in for by a
dict {feature name: value}: the protobuf framing is TensorFlow's, not the reference's; what the reference fixes -- and
what is pinned -- is which features it reads ('label' int64 scalar, 'image_raw' bytes) and what it does with the bytes."""
import numpy as np

from ._core import Tensor


class FixedLenFeature:
    def __init__(self, shape, dtype, default_value=None):
        self.shape, self.dtype = shape, dtype


def parse_single_example(serialized, features):
    out = {}
    for name, spec in features.items():
        assert name in serialized, "record lacks feature %r" % name
        val = serialized[name]
        if spec.dtype == "string":
            assert isinstance(val, (bytes, bytearray))
            out[name] = bytes(val)
        else:
            out[name] = Tensor(np.asarray(val, dtype=spec.dtype).reshape(spec.shape))
    return out


def decode_raw(input_bytes, out_type):
    return Tensor(np.frombuffer(input_bytes, dtype=out_type).copy())
