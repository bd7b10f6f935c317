"""Oracle: weight-normalisation wrapper and the uint8 record decode (numpy).  TEST INFRASTRUCTURE, see oracle/__init__.py.

SURVEY.md §8f row 4.  Follows /root/reference/sagan/layers.py:75-135,159-194 -- the TF-Addons `WeightNormalization`
wrapper that the `sagan/` tree ships under the name `SpectralNormalization` -- and the record reader's preprocessing,
/root/reference/sagan/dataset.py:27-40.
"""
import numpy as np

L2N_EPS = 1e-12      # tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), epsilon))


def kernel_from_vg(v, g):
    """sagan/layers.py:124: kernel = l2_normalize(v, axis=all but the last) * g."""
    axes = tuple(range(v.ndim - 1))                                       # sagan/layers.py:73 kernel_norm_axes
    ss = np.sum(v * v, axis=axes, keepdims=True)
    return v / np.sqrt(np.maximum(ss, L2N_EPS)) * g


def backward(dw, v, g):
    """Gradients of kernel_from_vg wrt v and g (tf.GradientTape in the reference)."""
    axes = tuple(range(v.ndim - 1))
    ss = np.maximum(np.sum(v * v, axis=axes, keepdims=True), L2N_EPS)
    inv = 1.0 / np.sqrt(ss)
    dot = np.sum(dw * v, axis=axes, keepdims=True)
    dv = g * inv * (dw - v * dot / ss)
    dg = (dot * inv).reshape(-1)
    return dv, dg


def init_norm(v):
    """sagan/layers.py:152-157 (data_init=False): g = ||v|| per filter."""
    return np.sqrt(np.sum(v.reshape(-1, v.shape[-1]) ** 2, axis=0))


def data_dep_init(x_init, g, bias=None):
    """sagan/layers.py:159-194 (data_init=True): x_init = the wrapped layer (raw kernel, no activation) applied to the
    first batch; scale = 1 / sqrt(var + 1e-10); g <- g * scale, bias <- -mean * scale."""
    axes = tuple(range(x_init.ndim - 1))
    m, var = x_init.mean(axis=axes), x_init.var(axis=axes)                # tf.nn.moments (biased variance)
    scale = 1.0 / np.sqrt(var + 1e-10)
    return g * scale, (None if bias is None else -m * scale)


def decode_records(raw_u8):
    """sagan/dataset.py:31-34: image = cast(uint8, float32) * (2. / 255) - 1.  (fp32 arithmetic, two rounded operations)."""
    return raw_u8.astype(np.float32) * np.float32(2.0 / 255) - np.float32(1.0)
