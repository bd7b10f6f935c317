"""Drop-in for the reference's top-level `layers` module.

/root/reference/sagan/models/generator.py:4 and discriminator.py:4 do
    from layers import SpectralNormalization, SNConv2D, SNDense, AttentionLayer
(the reference's `sagan/` directory is the script directory, so `layers` is a top-level module); the
legacy tree imports `Attention_Layer` (/root/reference/layers.py:71).  Put this directory on
sys.path in place of the reference's and those imports resolve to the B200 implementation.
"""
from sagan_b200.nn import (  # noqa: F401
    Attention_Layer,
    AttentionLayer,
    BatchNormalization,
    Conv2D,
    Conv2DTranspose,
    Dense,
    Embedding,
    LeakyReLU,
    ReLU,
    SNConv2D,
    SNDense,
    SpectralNormalization,
    WeightNormalization,
)


def l2normalize(v, eps=1e-12):
    """layers.py:4-5 (host-side helper; the kernels normalise on the device)."""
    return v / (v.norm() + eps)
