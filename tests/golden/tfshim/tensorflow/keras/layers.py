"""tf.keras.layers subset, float64 numpy, eager.  Semantics follow the Keras documentation of each layer:
HWIO conv kernels, [kh, kw, out, in] transposed-conv kernels, TF `SAME` padding (extra pixel AFTER), kernel-then-bias
weight order, `build` creating fresh variables every time it is called (which is what makes layers.py:87-90's double
build observable).  The convolution arithmetic is written with numpy window views / scatter-adds, on purpose not with
the torch functions the oracle uses, so the two are independent restatements of the same documented operators."""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from .._core import Tensor, TensorShape, Variable, raw

created = []            # every layer instantiated since the last keras.feed(), in creation order
_init_rng = [np.random.Generator(np.random.PCG64(0))]


def seed_initializers(seed):
    _init_rng[0] = np.random.Generator(np.random.PCG64(seed))


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def _same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


_depth = [0]            # > 0 while some layer's build() is running
_replay = [None]        # a list of already-built layers to hand out again (see replay()), or None


class replay:
    """with replay(recorded_layers): ... -- while active, constructing a layer returns the next object of the recorded
    sequence (same class, asserted) with its variables and `built` flag intact, so a builder function can be run a
    second time over layers whose kernels a script has set, without editing the builder."""

    def __init__(self, recorded):
        self.rec = [l for l in recorded if not l._inner]      # inner layers are not constructed again (owner is built)

    def __enter__(self):
        _replay[0] = self.rec
        return self

    def __exit__(self, *exc):
        _replay[0] = None
        return False

    def done(self):
        return not self.rec


class Layer:
    def __new__(cls, *args, **kwargs):
        if _replay[0] is None:
            return object.__new__(cls)
        obj = _replay[0].pop(0)
        assert type(obj) is cls, (type(obj), cls)
        obj._replayed = True
        return obj

    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        if getattr(self, "_replayed", False):
            return
        self.built = False
        self._weights = []
        self.name = name
        self._inner = _depth[0] > 0         # constructed inside another layer's build (Attention_Layer's four convs)
        created.append(self)

    @property
    def weights(self):
        return list(self._weights)

    def add_weight(self, name=None, shape=(), initializer="glorot_uniform", trainable=True, dtype=None, **_distribution):
        shape = () if shape is None else tuple(int(s) for s in shape)
        if dtype == "bool":
            a = np.zeros(shape, dtype=bool)
        elif initializer in ("zero", "zeros"):
            a = np.zeros(shape)
        elif initializer in ("one", "ones"):
            a = np.ones(shape)
        else:
            assert initializer in ("glorot_uniform", "uniform"), initializer
            if initializer == "uniform":
                a = _init_rng[0].uniform(-0.05, 0.05, shape)
            else:
                rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
                fan_in, fan_out = (shape[-2] * rf, shape[-1] * rf) if len(shape) >= 2 else (shape[0], shape[0])
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                a = _init_rng[0].uniform(-lim, lim, shape)
        v = Variable(a, name=name, trainable=trainable)
        self._weights.append(v)
        return v

    def build(self, input_shape):
        pass

    def call(self, inputs, *args, **kwargs):
        return inputs

    def _track_trackable(self, obj, name=None):
        return obj

    def get_weights(self):
        return [np.array(w.numpy()) for w in self._weights]

    def set_weights(self, arrays):
        assert len(arrays) == len(self._weights)
        for w, a in zip(self._weights, arrays):
            w.assign(a)

    def get_config(self):
        return dict(getattr(self, "_config", {}), trainable=True)

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            shp = [t.shape for t in inputs] if isinstance(inputs, (list, tuple)) else inputs.shape
            _depth[0] += 1
            try:
                self.build(shp)
            finally:
                _depth[0] -= 1
            self.built = True
        return self.call(inputs, *args, **kwargs)


def _activation(name):
    if name is None:
        return lambda a: a
    if name == "tanh":
        return np.tanh
    if name == "relu":
        return lambda a: np.maximum(a, 0.0)
    raise ValueError(name)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None, **kw):
        super().__init__(**kw)
        self._config = dict(filters=filters, kernel_size=kernel_size, strides=strides, padding=padding,
                            use_bias=use_bias, activation=activation)
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.use_bias, self.activation = padding.lower(), use_bias, activation

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", self.kernel_size + (cin, self.filters))
        self.bias = self.add_weight("bias", (self.filters,), "zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x, w = raw(inputs), raw(self.kernel)
        (kh, kw), (sh, sw) = self.kernel_size, self.strides
        if self.padding == "same":
            ph, pw = _same_pads(x.shape[1], kh, sh), _same_pads(x.shape[2], kw, sw)
            x = np.pad(x, ((0, 0), ph, pw, (0, 0)))
        win = sliding_window_view(x, (kh, kw), axis=(1, 2))[:, ::sh, ::sw]      # [B, Ho, Wo, Cin, kh, kw]
        y = np.einsum("bhwcij,ijco->bhwo", win, w, optimize=True)
        if self.bias is not None:
            y = y + raw(self.bias)
        return Tensor(_activation(self.activation)(y))


class Conv2DTranspose(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None, **kw):
        super().__init__(**kw)
        self._config = dict(filters=filters, kernel_size=kernel_size, strides=strides, padding=padding,
                            use_bias=use_bias, activation=activation)
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.use_bias, self.activation = padding.lower(), use_bias, activation

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", self.kernel_size + (self.filters, cin))
        self.bias = self.add_weight("bias", (self.filters,), "zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        """The gradient of Conv2D w.r.t. its input (what TF's conv2d_transpose is): every input pixel scatters
        kernel[i, j] into the stride-dilated output; 'same' keeps in * stride pixels, cropping the forward conv's
        leading pad."""
        x, w = raw(inputs), raw(self.kernel)
        (kh, kw), (sh, sw) = self.kernel_size, self.strides
        B, H, W, _ = x.shape
        full = np.zeros((B, (H - 1) * sh + kh, (W - 1) * sw + kw, self.filters))
        for i in range(kh):
            for j in range(kw):
                full[:, i:i + H * sh:sh, j:j + W * sw:sw, :] += np.einsum("bhwc,oc->bhwo", x, w[i, j], optimize=True)
        if self.padding == "same":
            oh, ow = H * sh, W * sw
            pl_h, pl_w = _same_pads(oh, kh, sh)[0], _same_pads(ow, kw, sw)[0]
            full = full[:, pl_h:pl_h + oh, pl_w:pl_w + ow, :]
        if self.bias is not None:
            full = full + raw(self.bias)
        return Tensor(_activation(self.activation)(full))


class Dense(Layer):
    def __init__(self, units, use_bias=True, activation=None, **kw):
        super().__init__(**kw)
        self._config = dict(units=units, use_bias=use_bias, activation=activation)
        self.units, self.use_bias, self.activation = int(units), use_bias, activation

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units))
        self.bias = self.add_weight("bias", (self.units,), "zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        y = raw(inputs) @ raw(self.kernel)
        if self.bias is not None:
            y = y + raw(self.bias)
        return Tensor(_activation(self.activation)(y))


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **kw):
        super().__init__(**kw)
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)

    def build(self, input_shape):
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim), "uniform")
        self.built = True

    def call(self, inputs):
        return Tensor(raw(self.embeddings)[raw(inputs).astype(np.int64)])


class BatchNormalization(Layer):
    """Training-mode statistics (biased batch variance over B, H, W), epsilon 1e-3, momentum 0.99 -- the Keras
    defaults the reference relies on (sagan/models/generator.py:10)."""

    def __init__(self, momentum=0.99, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", (c,), "ones")
        self.beta = self.add_weight("beta", (c,), "zeros")
        self.moving_mean = self.add_weight("moving_mean", (c,), "zeros", trainable=False)
        self.moving_variance = self.add_weight("moving_variance", (c,), "ones", trainable=False)
        self.built = True

    def call(self, inputs, training=True):
        x = raw(inputs)
        if not training:
            return Tensor((x - raw(self.moving_mean)) / np.sqrt(raw(self.moving_variance) + self.epsilon)
                          * raw(self.gamma) + raw(self.beta))
        ax = tuple(range(x.ndim - 1))
        mean, var = x.mean(axis=ax), x.var(axis=ax)
        self.moving_mean.assign(raw(self.moving_mean) * self.momentum + mean * (1 - self.momentum))
        self.moving_variance.assign(raw(self.moving_variance) * self.momentum + var * (1 - self.momentum))
        return Tensor((x - mean) / np.sqrt(var + self.epsilon) * raw(self.gamma) + raw(self.beta))


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kw):
        super().__init__(**kw)
        self.alpha = float(alpha)

    def call(self, inputs):
        a = raw(inputs)
        return Tensor(np.where(a >= 0, a, self.alpha * a))


class ReLU(Layer):
    def call(self, inputs):
        return Tensor(np.maximum(raw(inputs), 0.0))


class MaxPool2D(Layer):
    def __init__(self, pool_size=2, strides=None, padding="valid", **kw):
        super().__init__(**kw)
        self.pool_size = _pair(pool_size)
        self.strides = self.pool_size if strides is None else _pair(strides)
        assert padding == "valid"

    def call(self, inputs):
        x = raw(inputs)
        win = sliding_window_view(x, self.pool_size, axis=(1, 2))[:, ::self.strides[0], ::self.strides[1]]
        return Tensor(win.max(axis=(-2, -1)))


MaxPooling2D = MaxPool2D


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, inputs):
        return Tensor(np.concatenate([raw(t) for t in inputs], axis=self.axis))


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super().__init__(**kw)
        self.target_shape = tuple(target_shape)

    def call(self, inputs):
        x = raw(inputs)
        return Tensor(x.reshape((x.shape[0],) + self.target_shape))


class Wrapper(Layer):
    """tf.keras.layers.Wrapper: holds `layer`."""

    def __init__(self, layer, **kw):
        super().__init__(**kw)
        self.layer = layer


class RNN(Layer):
    pass


class InputSpec:
    def __init__(self, shape=None, **kw):
        self.shape = shape


def serialize(layer):
    return {"class_name": type(layer).__name__, "config": layer.get_config()}


def deserialize(config):
    cfg = dict(config["config"])
    cfg.pop("trainable", None)
    return globals()[config["class_name"]](**cfg)


def add(inputs):
    out = raw(inputs[0])
    for t in inputs[1:]:
        out = out + raw(t)
    return Tensor(out)


__all__ = ["Layer", "ReLU", "Conv2D", "Conv2DTranspose", "Dense", "Embedding", "BatchNormalization", "LeakyReLU", "MaxPool2D",
           "MaxPooling2D", "Concatenate", "Reshape", "add", "TensorShape"]
