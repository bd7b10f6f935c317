"""Keras-style layer surface of the reference, hosted on torch.nn.Module (device memory + autograd only).

Mirrors what the reference's model builders consume:
  SpectralNormalization(module, name="weights", Ip=1, factor=None)   /root/reference/layers.py:7-68
  Attention_Layer() / AttentionLayer()                               /root/reference/layers.py:71-120,
                                                                     sagan/models/generator.py:4,34
  SNConv2D / SNDense (imported by sagan/models/discriminator.py:4, never defined there)
plus the stock Keras layers those builders wrap (Conv2D, Conv2DTranspose, Dense,
BatchNormalization, LeakyReLU).  Layers follow the Keras protocol: weights are created by
`build(input_shape)` on first call, `call(inputs[, training])` does the work, activations are NHWC,
kernels use the Keras layouts.  All arithmetic is in libsagan_b200.so (see functional.py).
"""
import math

import torch

from . import functional as F
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, MATH_BF16_TC, MATH_FP32_STRICT  # noqa: F401

_DEFAULT_MATH = [MATH_FP32_STRICT]


def set_default_math_mode(mode):
    """MATH_FP32_STRICT (1e-5 tier) or MATH_BF16_TC (2e-3 tier) for layers created afterwards."""
    _DEFAULT_MATH[0] = mode


def _device():
    if not torch.cuda.is_available():
        raise F._lib.SaganError("sagan_b200 layers need a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _glorot_uniform(shape, fan_in, fan_out, gen=None):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * lim


_LEARNING_PHASE = [True]


def set_learning_phase(value):
    """tf.keras.backend.set_learning_phase (sagan/main.py:239): the default `training` of BatchNormalization calls that
    do not pass one.  nets.Network sets it from its own `training` argument for the duration of a forward."""
    _LEARNING_PHASE[0] = bool(value)


def learning_phase():
    return _LEARNING_PHASE[0]


class Layer(torch.nn.Module):
    """Keras Layer protocol: build(input_shape) once, then call()."""

    def __init__(self):
        super().__init__()
        self.built = False

    def build(self, input_shape):
        self.built = True

    def forward(self, x, *args, **kwargs):
        if not self.built:
            self.build(tuple(x.shape))
            self.built = True
        return self.call(x, *args, **kwargs)

    def _p(self, name):
        """The parameter `name`, or its proxy for the current network forward (nets.Network hands out proxies whose
        gradients are accumulated into the flat bucket by one launch, functional.param_proxies)."""
        px = self.__dict__.get("_px")
        if px is not None and name in px:
            return px[name]
        return getattr(self, name)

    @property
    def weights(self):
        """Keras order: kernel first (SpectralNormalization reads module.weights[0], layers.py:31,55)."""
        return [p for p in (getattr(self, "kernel", None), getattr(self, "bias", None)) if p is not None]


_ACTS = {None: ACT_NONE, "linear": ACT_NONE, "tanh": ACT_TANH}


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None,
                 math_mode=None, leaky_slope=None):
        super().__init__()
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        self.filters, self.kernel_size = int(filters), (int(ks[0]), int(ks[1]))
        self.strides = int(strides[0] if isinstance(strides, (tuple, list)) else strides)
        self.padding, self.use_bias = padding, use_bias
        if activation not in _ACTS:
            raise ValueError(f"unsupported activation {activation!r}")
        self.activation = _ACTS[activation]
        # leaky_slope fuses the LeakyReLU(alpha) that follows the conv at sagan/models/discriminator.py:10
        self.leaky_slope = 0.0
        if leaky_slope is not None:
            if activation is not None:
                raise ValueError("give either activation or leaky_slope")
            self.activation, self.leaky_slope = ACT_LRELU, float(leaky_slope)
        self.math_mode = _DEFAULT_MATH[0] if math_mode is None else math_mode
        self.kernel = self.bias = None

    def build(self, input_shape):
        cin = int(input_shape[-1])
        kh, kw = self.kernel_size
        k = _glorot_uniform((kh, kw, cin, self.filters), cin * kh * kw, self.filters * kh * kw)
        self.kernel = torch.nn.Parameter(k.to(_device()))
        if self.use_bias:
            self.bias = torch.nn.Parameter(torch.zeros(self.filters, device=_device()))
        self.built = True

    def call_with_kernel(self, x, kernel):
        return F.conv2d(x, kernel, self._p("bias"), self.strides, self.padding, self.activation, self.leaky_slope,
                        self.math_mode)

    def call(self, x):
        return self.call_with_kernel(x, self._p("kernel"))


class Conv2DTranspose(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None,
                 math_mode=None):
        super().__init__()
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        self.filters, self.kernel_size = int(filters), (int(ks[0]), int(ks[1]))
        self.strides = int(strides[0] if isinstance(strides, (tuple, list)) else strides)
        self.padding = padding
        if activation is not None:
            raise ValueError("Conv2DTranspose: only activation=None is built "
                             "(the forms used at sagan/models/generator.py:8 and models/generator.py:11,18)")
        self.use_bias = use_bias
        self.math_mode = _DEFAULT_MATH[0] if math_mode is None else math_mode
        self.kernel = self.bias = None

    def build(self, input_shape):
        cin = int(input_shape[-1])
        kh, kw = self.kernel_size
        # Keras kernel layout [kh, kw, cout, cin]
        k = _glorot_uniform((kh, kw, self.filters, cin), cin * kh * kw, self.filters * kh * kw)
        self.kernel = torch.nn.Parameter(k.to(_device()))
        if self.use_bias:
            self.bias = torch.nn.Parameter(torch.zeros(self.filters, device=_device()))
        self.built = True

    def call_with_kernel(self, x, kernel):
        y = F.conv2d_transpose(x, kernel, self.strides, self.padding, self.math_mode)
        return F.bias_add(y, self._p("bias")) if self.bias is not None else y

    def call(self, x):
        return self.call_with_kernel(x, self._p("kernel"))


class Dense(Layer):
    def __init__(self, units, use_bias=True, math_mode=None):
        super().__init__()
        self.units, self.use_bias = int(units), use_bias
        self.math_mode = _DEFAULT_MATH[0] if math_mode is None else math_mode
        self.kernel = self.bias = None

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = torch.nn.Parameter(_glorot_uniform((cin, self.units), cin, self.units).to(_device()))
        if self.use_bias:
            self.bias = torch.nn.Parameter(torch.zeros(self.units, device=_device()))
        self.built = True

    def call_with_kernel(self, x, kernel):
        return F.dense(x, kernel, self._p("bias"), self.math_mode)

    def call(self, x):
        return self.call_with_kernel(x, self._p("kernel"))


class BatchNormalization(Layer):
    """Keras BatchNormalization (eps 1e-3, momentum 0.99), training-mode batch statistics per replica.
    `leaky_slope` fuses the LeakyReLU that follows it at sagan/models/generator.py:10-11 (1.0 = none)."""

    def __init__(self, momentum=0.99, epsilon=1e-3, leaky_slope=1.0):
        super().__init__()
        self.momentum, self.epsilon, self.leaky_slope = momentum, epsilon, leaky_slope

    def build(self, input_shape):
        c = int(input_shape[-1])
        dev = _device()
        self.gamma = torch.nn.Parameter(torch.ones(c, device=dev))
        self.beta = torch.nn.Parameter(torch.zeros(c, device=dev))
        self.register_buffer("moving_mean", torch.zeros(c, device=dev))
        self.register_buffer("moving_var", torch.ones(c, device=dev))
        self.built = True

    def call(self, x, training=None):
        if training is None:
            training = learning_phase()
        if not training:
            return F.batchnorm_lrelu_infer(x, self.gamma, self.beta, self.moving_mean, self.moving_var, self.epsilon,
                                           self.leaky_slope)
        return F.batchnorm_lrelu(x, self._p("gamma"), self._p("beta"), self.moving_mean, self.moving_var, self.epsilon,
                                 self.momentum, self.leaky_slope)


class SpectralNormalization(Layer):
    """layers.py:7-68.  `SpectralNormalization(layer)(x)` runs `Ip` power iterations on the wrapped
    layer's kernel (raw-reshaped to [last axis, -1], layers.py:56) and calls the layer with W / sigma.

    Readings fixed where the literal reference is ill-formed: W/sigma is
    what the wrapped layer uses; the iteration runs on every training call; u / v persist.
    """

    def __init__(self, module, name="weights", Ip=1, factor=None, **kwargs):
        super().__init__()
        if not Ip >= 1:
            # layers.py:17-18
            raise ValueError("The number of power iterations should be positive integer")
        self.module = module
        self.weight_name = name
        self.Ip = int(Ip)
        self.factor = factor
        self._group = None       # SpectralNormGroup that owns u / v / sigma / W_bar
        self._index = 0
        self._pending = None     # W_bar handed down by an enclosing group for the current forward

    # -- Keras-visible state ------------------------------------------------------------------
    @property
    def u(self):
        return self._group.u(self._index).view(1, -1) if self._group is not None else None

    @property
    def v(self):
        return self._group.v(self._index).view(1, -1) if self._group is not None else None

    @property
    def sigma(self):
        return self._group.sigma(self._index) if self._group is not None else None

    def _kernel(self):
        return getattr(self.module, self.weight_name)[0]     # layers.py:31,55

    def _make_param(self, u_init=None):
        # layers.py:30-38: u ~ N(0,1) [1, rows], l2-normalised (v is recomputed from u before its first use)
        w = self._kernel()
        rows = w.shape[-1]
        if u_init is None:
            u_init = torch.randn(1, rows)
            u_init = u_init / (u_init.norm() + 1e-12)
        self._group = F.SpectralNormGroup([w], [u_init], self.Ip, [self.factor])
        self._index = 0
        # layers.py:36,38: v ~ N(0,1) [1, cols], l2-normalised.  Every training call recomputes v from u before using
        # it; a training=False call before the first training call reads this one.
        # (drawn from a private generator so that the default stream -- and with it every initial kernel that follows --
        # is what it was before v was materialised here)
        v_init = torch.randn(1, w.numel() // rows, generator=torch.Generator().manual_seed(0x5EED + w.numel()))
        self._group.v(0).copy_((v_init / (v_init.norm() + 1e-12)).reshape(-1).to(self._group.out.device))

    def build(self, input_shape):
        if not self.module.built:
            self.module.build(input_shape)                   # layers.py:41
            self.module.built = True
        if self._group is None:
            self._make_param()                               # layers.py:42-43
        self.built = True

    def adopt(self, group, index):
        """Move this wrapper's state into a multi-tensor group (one launch for a whole network)."""
        self._group, self._index = group, index

    def update_uv(self):
        """layers.py:50-68 for this kernel alone; returns W / sigma."""
        if self._group is None:
            raise RuntimeError("SpectralNormalization.update_uv called before build()")
        if len(self._group.weights) == 1:
            return self._group.normalized(update=True)[0]
        solo = F.SpectralNormGroup([self._kernel()], [self.u], self.Ip, [self.factor])
        out = solo.normalized(update=True)[0]
        self._group.u(self._index).copy_(solo.u(0))
        return out

    def call(self, x, training=None):
        if self._pending is not None:                        # an enclosing network already normalised us
            w_bar, self._pending = self._pending, None
        else:
            training = self.training if training is None else training
            if training:
                w_bar = self.update_uv()
            else:
                w_bar = self._group.normalized(update=False)[self._index]
        return self.module.call_with_kernel(x, w_bar)


class WeightNormalization(Layer):
    """sagan/layers.py:6-211: the weight-normalisation wrapper the `sagan/` tree ships (there under the NAME
    `SpectralNormalization`): kernel = l2_normalize(v, all axes but the last) * g, with `v` the wrapped layer's kernel
    and a trainable per-filter `g`.  On the first call g (and the layer's bias) are initialised from the data
    (`data_init=True`, sagan/layers.py:159-194) or from ||v|| (sagan/layers.py:152-157)."""

    def __init__(self, layer, data_init=True, **kwargs):
        super().__init__()
        if not isinstance(layer, Layer):
            raise ValueError("Please initialize `WeightNormalization` layer with a `Layer` instance")
        self.layer = layer
        self.data_init = data_init
        self._initialized = False

    def build(self, input_shape):
        if not self.layer.built:
            self.layer.build(input_shape)                    # sagan/layers.py:57-58
            self.layer.built = True
        if getattr(self.layer, "kernel", None) is None:      # sagan/layers.py:62-64
            raise ValueError("`WeightNormalization` must wrap a layer that contains a `kernel` for weights")
        self.v = self.layer.kernel                           # sagan/layers.py:83
        self.layer_depth = int(self.v.shape[-1])             # sagan/layers.py:72
        self.g = torch.nn.Parameter(torch.ones(self.layer_depth, device=self.v.device))   # sagan/layers.py:75-82
        self.built = True

    @torch.no_grad()
    def _initialize_weights(self, x):
        if self.data_init:                                   # sagan/layers.py:159-194
            act, slope = getattr(self.layer, "activation", ACT_NONE), getattr(self.layer, "leaky_slope", 0.0)
            if hasattr(self.layer, "activation"):
                self.layer.activation = ACT_NONE             # the naked clone has no activation (sagan/layers.py:104-105)
            x_init = self.layer.call_with_kernel(x, self.v)
            if hasattr(self.layer, "activation"):
                self.layer.activation, self.layer.leaky_slope = act, slope
            mean, scale = F.batch_moments(x_init, 1e-10)     # scale = 1 / sqrt(var + 1e-10)
            self.g.mul_(scale)
            if getattr(self.layer, "bias", None) is not None:
                self.layer.bias.copy_(-mean * scale)
        else:                                                # sagan/layers.py:152-157
            self.g.copy_(self.v.reshape(-1, self.layer_depth).norm(dim=0))
        self._initialized = True

    def call(self, x):
        if not self._initialized:                            # sagan/layers.py:109-121
            self._initialize_weights(x)
        kernel = F.weight_norm(self.v, self.g)               # sagan/layers.py:124
        return self.layer.call_with_kernel(x, kernel)

    def remove(self):
        """sagan/layers.py:201-211: bake the normalised kernel into the wrapped layer and return it."""
        with torch.no_grad():
            self.layer.kernel.copy_(F.weight_norm(self.v, self.g))
        return self.layer


def SNConv2D(filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None, **sn_kwargs):
    """Name imported at sagan/models/discriminator.py:4 (never defined in the reference):
    spectrally-normalised Conv2D."""
    return SpectralNormalization(Conv2D(filters, kernel_size, strides, padding, use_bias, activation), **sn_kwargs)


def SNDense(units, use_bias=True, **sn_kwargs):
    """Name imported at sagan/models/discriminator.py:4: spectrally-normalised Dense."""
    return SpectralNormalization(Dense(units, use_bias), **sn_kwargs)


class LeakyReLU(Layer):
    """Stand-alone LeakyReLU(alpha) (the pre-activation blocks of models/discriminator.py:22-34); inside the vanilla
    builders it is fused into the producing kernel instead."""

    def __init__(self, alpha=0.3):
        super().__init__()
        self.alpha = float(alpha)

    def call(self, x):
        return F.activation(x, ACT_LRELU, self.alpha)


class ReLU(LeakyReLU):
    def __init__(self):
        super().__init__(0.0)


def add(tensors):
    """keras `layers.add([a, b])` (models/generator.py:21, models/discriminator.py:17,38)."""
    a, b = tensors
    return F.add(a, b)


class Embedding(Layer):
    """keras Embedding(input_dim, output_dim): `kernel` = the [input_dim, output_dim] table (weights[0], what
    SpectralNormalization normalises at models/discriminator.py:53-54); the lookup itself is a gather."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.kernel = self.bias = None

    def build(self, input_shape=None):
        self.kernel = torch.nn.Parameter((torch.rand(self.input_dim, self.output_dim) * 0.1 - 0.05).to(_device()))
        self.built = True

    def forward(self, labels, *args, **kwargs):
        if not self.built:
            self.build()
        return self.call(labels, *args, **kwargs)

    def call_with_kernel(self, labels, kernel):
        return kernel.index_select(0, labels.long())

    def call(self, labels):
        return self.call_with_kernel(labels, self._p("kernel"))


class Attention_Layer(Layer):
    """layers.py:71-120 (paper form).  One fused kernel:
    Y = X + sigma * ((softmax((X Wtheta + b)(X Wphi + b)^T) (X Wg + b)) Wo + bo)."""

    def __init__(self, math_mode=None, pool=None):
        """pool=None: paper form, every token is a key (the oracle's fixed reading).
        pool='2x2s2': phi and g are max-pooled 2x2 / stride 2 before the attention map, the down-sampling
        layers.py:96,100,113 reaches for ("downsampled attn layer", example_configs/church64_attn.py:3)."""
        super().__init__()
        self.math_mode = _DEFAULT_MATH[0] if math_mode is None else math_mode
        if pool not in (None, "2x2s2"):
            raise ValueError(f"pool must be None or '2x2s2', got {pool!r}")
        self.pool = pool

    def build(self, input_shape):
        b, w, h, c = [int(s) if s is not None else None for s in input_shape]
        if c % 8 != 0:
            raise ValueError(f"Attention_Layer needs channels divisible by 8, got {c}")
        # layers.py:76-79: scalar `sigma` (gamma), zero-initialised, trainable
        self.sigma = torch.nn.Parameter(torch.zeros((), device=_device()))
        # layers.py:81-85: phi, theta (c//8), g (c//2), out (c) -- 1x1 convs with bias, all SN-wrapped
        self.SN_conv = torch.nn.ModuleList([
            SpectralNormalization(Conv2D(c // 8, 1, 1)),
            SpectralNormalization(Conv2D(c // 8, 1, 1)),
            SpectralNormalization(Conv2D(c // 2, 1, 1)),
            SpectralNormalization(Conv2D(c, 1, 1)),
        ])
        for i in range(3):
            self.SN_conv[i].build(input_shape)               # layers.py:87-88
        self.SN_conv[3].build((b, w, h, c // 2))             # layers.py:90
        self.built = True

    def call(self, x, training=None):
        B, H, W, Cc = x.shape
        kernels = []
        for sn in self.SN_conv:
            if sn._pending is not None:
                wb, sn._pending = sn._pending, None
            else:
                tr = self.training if training is None else training
                wb = sn.update_uv() if tr else sn._group.normalized(update=False)[sn._index]
            kernels.append(wb.reshape(wb.shape[2], wb.shape[3]))
        phi, theta, g, o = self.SN_conv
        y = F.attention(x.reshape(B, H * W, Cc),
                        kernels[1], theta.module._p("bias"), # queries: theta, layers.py:104-105
                        kernels[0], phi.module._p("bias"),   # keys:    phi,   layers.py:99
                        kernels[2], g.module._p("bias"),     # values:  g,     layers.py:112
                        kernels[3], o.module._p("bias"),     # output conv,    layers.py:119
                        self._p("sigma"), self.math_mode,
                        (H, W) if self.pool else None)       # layers.py:100,113 (MaxPool2D on phi and g)
        return y.reshape(B, H, W, Cc)


AttentionLayer = Attention_Layer     # name used by sagan/models/generator.py:4,34
