"""Oracle: vanilla SAGAN generator / discriminator (torch-CPU, un-fused).  TEST INFRASTRUCTURE.

Follows
  /root/reference/sagan/models/generator.py:7-37      (Block, get_generator)
  /root/reference/sagan/models/discriminator.py:7-36  (Block, get_discriminator)
with the layer math of /root/reference/layers.py:4-120 (oracle.sn / oracle.attention
restate it in numpy; here the same ops are written with torch so that autograd plays
the role of tf.GradientTape, sagan/main.py:180,197).

Keras/TF semantics restated from knowledge (TensorFlow is not in /root/reference):
  * NHWC activations, Conv2D kernel HWIO [kh,kw,cin,cout], Conv2DTranspose kernel
    [kh,kw,cout,cin], Dense kernel [in,out]
  * padding='same': out = ceil(in/s), pad_total = max((out-1)*s + k - in, 0),
    pad_before = pad_total // 2 (so k=4,s=1 pads 1 before / 2 after)
  * BatchNormalization: eps 1e-3, momentum 0.99, training mode uses biased batch variance
  * LeakyReLU(alpha=0.1)

Parameters live in a flat dict name -> tensor (Keras layouts), spectral-norm `u`
vectors in a second dict (persistent, updated in place by every training forward).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
LRELU = 0.1
SN_EPS = 1e-12


# ----------------------------------------------------------------------------- layers
def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2, out


def conv2d_same(x, w_hwio, bias, stride):
    """x NHWC, w [kh,kw,cin,cout]; TF padding='same'."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb, _ = same_pad(x.shape[1], kh, stride)
    pl, pr, _ = same_pad(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn.contiguous(), w_hwio.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, w_hwoi, stride):
    """x NHWC [B,H,W,cin], w [kh,kw,cout,cin]; TF padding='same' => out = in*stride.

    Keras Conv2DTranspose is the gradient of a 'same' Conv2D mapping the output
    grid back to the input grid: full transposed conv, then crop pad_before.
    """
    kh, kw = w_hwoi.shape[0], w_hwoi.shape[1]
    H, W = x.shape[1], x.shape[2]
    pt, _, _ = same_pad(H * stride, kh, stride)
    pl, _, _ = same_pad(W * stride, kw, stride)
    # torch conv_transpose2d weight: [cin, cout, kh, kw]
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2).contiguous(), w_hwoi.permute(3, 2, 0, 1).contiguous(), stride=stride)
    need_h, need_w = pt + H * stride, pl + W * stride
    if y.shape[2] < need_h or y.shape[3] < need_w:
        y = F.pad(y, (0, max(need_w - y.shape[3], 0), 0, max(need_h - y.shape[2], 0)))
    y = y[:, :, pt:pt + H * stride, pl:pl + W * stride]
    return y.permute(0, 2, 3, 1)


def l2n(v):
    return v / (torch.linalg.vector_norm(v) + SN_EPS)


def spectral_norm(W, sn_state, key, training=True, Ip=1, factor=None):
    """layers.py:45-68 in the reading of SURVEY.md §8c: returns W / sigma.

    u, v are constants in the backward; sigma = sum((u W_mat) * v) stays a
    differentiable function of W (the sngan_projection semantics cited at layers.py:8-9).
    """
    Wm = W.reshape(W.shape[-1], -1)                       # layers.py:56 (raw reshape)
    u = sn_state[key]
    with torch.no_grad():
        if training:
            for _ in range(Ip):
                v = l2n(u @ Wm)                           # layers.py:59
                u = l2n(v @ Wm.t())                       # layers.py:60
            sn_state[key] = u
            sn_state[key + ":v"] = v
        else:
            v = sn_state.get(key + ":v")
            if v is None:
                v = l2n(u @ Wm)
    sigma = torch.sum((u @ Wm) * v)                       # layers.py:62
    if factor:
        sigma = sigma / factor
    sn_state[key + ":sigma"] = sigma.detach()
    return W / sigma                                      # layers.py:68


def batchnorm_train(x, gamma, beta, stats=None, key=None):
    """Keras BatchNormalization(training=True) over NHWC (per-replica batch stats)."""
    mean = x.mean(dim=(0, 1, 2))
    var = x.var(dim=(0, 1, 2), unbiased=False)
    if stats is not None:
        with torch.no_grad():
            stats[key + ".moving_mean"] = BN_MOMENTUM * stats[key + ".moving_mean"] + (1 - BN_MOMENTUM) * mean
            stats[key + ".moving_var"] = BN_MOMENTUM * stats[key + ".moving_var"] + (1 - BN_MOMENTUM) * var
    return (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta


def batchnorm_infer(x, gamma, beta, stats, key):
    """Keras BatchNormalization(training=False): the moving statistics (generator(..., training=False), the sample
    dumps of sagan/main.py:333)."""
    return (x - stats[key + ".moving_mean"]) * torch.rsqrt(stats[key + ".moving_var"] + BN_EPS) * gamma + beta


def attention(x, p, sn_state, prefix, training=True, downsample=False):
    """layers.py:93-120 (paper form).  x NHWC.  downsample=True: phi and g max-pooled 2x2 / stride 2 before the
    attention map (layers.py:100,113 in the well-formed reading of oracle.attention.forward_pooled)."""
    B, H, W, C = x.shape
    X = x.reshape(B, H * W, C)

    def proj(name, inp):
        k = spectral_norm(p[f"{prefix}.{name}.kernel"], sn_state, f"{prefix}.{name}.u", training)
        return inp @ k.reshape(k.shape[2], k.shape[3]) + p[f"{prefix}.{name}.bias"]

    def pool(t):                                           # layers.py:100,113
        c = t.shape[-1]
        t = torch.nn.functional.max_pool2d(t.reshape(B, H, W, c).permute(0, 3, 1, 2), 2, 2)
        return t.permute(0, 2, 3, 1).reshape(B, (H // 2) * (W // 2), c)

    phi = proj("phi", X)                                   # layers.py:99
    theta = proj("theta", X)                               # layers.py:104-105
    g = proj("g", X)                                       # layers.py:112-114
    if downsample:
        phi, g = pool(phi), pool(g)
    S = theta @ phi.transpose(1, 2)                        # layers.py:108
    P = torch.softmax(S, dim=-1)                           # layers.py:109
    A = P @ g                                              # layers.py:116
    O = proj("o", A)                                       # layers.py:119
    return (X + p[f"{prefix}.sigma"] * O).reshape(B, H, W, C)   # layers.py:120


# ----------------------------------------------------------------------------- specs
def _power(cfg):
    return int(np.log2(cfg["img_size"] / 4))              # generator.py:28, discriminator.py:20


def _attn_spec(prefix, C):
    d, dv = C // 8, C // 2
    return [
        (f"{prefix}.phi.kernel", (1, 1, C, d)), (f"{prefix}.phi.bias", (d,)),
        (f"{prefix}.theta.kernel", (1, 1, C, d)), (f"{prefix}.theta.bias", (d,)),
        (f"{prefix}.g.kernel", (1, 1, C, dv)), (f"{prefix}.g.bias", (dv,)),
        (f"{prefix}.o.kernel", (1, 1, dv, C)), (f"{prefix}.o.bias", (C,)),
        (f"{prefix}.sigma", ()),
    ]


def generator_spec(cfg):
    """Ordered (name, shape) list of trainable G parameters (generator.py:14-37)."""
    gf = cfg["gf_dim"]
    zin = cfg["z_dim"] + (cfg["num_classes"] if cfg.get("use_label") else 0)
    spec = [("dense.kernel", (zin, 4 * 4 * gf * 16)), ("dense.bias", (4 * 4 * gf * 16,))]
    cin, size = gf * 16, 4
    for i, p in enumerate(reversed(range(_power(cfg)))):
        cout = gf * (2 ** p)
        spec += [(f"block{i}.deconv.kernel", (4, 4, cout, cin)),
                 (f"block{i}.bn.gamma", (cout,)), (f"block{i}.bn.beta", (cout,))]
        size *= 2
        if cfg.get("use_attention") and size in cfg["attn_dim_G"]:
            spec += _attn_spec(f"block{i}.attn", cout)
        cin = cout
    spec += [("head.kernel", (4, 4, cin, 3))]
    return spec


def discriminator_spec(cfg):
    """Ordered (name, shape) list of trainable D parameters (discriminator.py:13-36)."""
    df = cfg["df_dim"]
    spec = []
    cin, size = 3, cfg["img_size"]
    for i, p in enumerate(range(_power(cfg))):
        cout = df * 2 ** p
        spec += [(f"block{i}.conv.kernel", (4, 4, cin, cout)), (f"block{i}.conv.bias", (cout,))]
        size //= 2
        # discriminator.py:23 reads attn_dim_G (sic), attn_dim_D is ignored
        if cfg.get("use_attention") and size in cfg["attn_dim_G"]:
            spec += _attn_spec(f"block{i}.attn", cout)
        cin = cout
    if cfg.get("use_label"):
        spec += [("head.dense.kernel", (cin, 1)), ("head.dense.bias", (1,)),
                 ("head.embedding", (cfg["num_classes"], cin))]
    else:
        spec += [("head.kernel", (4, 4, cin, 1)), ("head.bias", (1,))]
    return spec


def sn_keys(spec):
    """Names of spectrally-normalised kernels -> their `u` key and [R, K] matrix shape."""
    out = OrderedDict()
    for name, shape in spec:
        if name.endswith(".kernel") and not name.startswith("head"):
            R = shape[-1]
            K = int(np.prod(shape)) // R
            out[name[:-len("kernel")] + "u"] = (R, K)
    return out


def init_params(spec, seed=0, dtype=torch.float64, attn_sigma=0.0, bias_scale=0.0):
    """Glorot-uniform kernels (Keras default), zero biases (or small random ones when
    bias_scale > 0 so that bias paths are exercised), BN gamma=1 / beta=0, attention
    `sigma` = attn_sigma (Keras zero-init; tests also use 0.37)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = OrderedDict()
    for name, shape in spec:
        if name.endswith("sigma"):
            a = np.full(shape, attn_sigma)
        elif name.endswith("bn.gamma"):
            a = np.ones(shape) + bias_scale * rng.standard_normal(shape)
        elif name.endswith("bias") or name.endswith("bn.beta"):
            a = bias_scale * rng.standard_normal(shape)
        elif name.endswith("embedding"):
            a = rng.uniform(-0.05, 0.05, shape)
        else:
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            if "deconv" in name:                           # Keras Conv2DTranspose: [kh,kw,cout,cin]
                fan_in, fan_out = shape[-1] * rf, shape[-2] * rf
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            a = rng.uniform(-lim, lim, shape)
        p[name] = torch.tensor(np.asarray(a), dtype=dtype)
    return p


def init_sn_state(spec, seed=1, dtype=torch.float64):
    """layers.py:30-38: u ~ N(0,1)[1,R] l2-normalised (v is recomputed from u before use)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    st = OrderedDict()
    for key, (R, K) in sn_keys(spec).items():
        u = torch.tensor(rng.standard_normal((1, R)), dtype=dtype)
        st[key] = l2n(u)
    return st


def init_bn_stats(spec, dtype=torch.float64):
    st = {}
    for name, shape in spec:
        if name.endswith("bn.gamma"):
            k = name[:-len(".gamma")]
            st[k + ".moving_mean"] = torch.zeros(shape, dtype=dtype)
            st[k + ".moving_var"] = torch.ones(shape, dtype=dtype)
    return st


# ----------------------------------------------------------------------------- models
def generator_forward(p, sn_state, z, cfg, labels=None, training=True, bn_stats=None):
    """generator.py:14-37.  z [B, z_dim] -> images NHWC [B, S, S, 3] in (-1, 1)."""
    gf = cfg["gf_dim"]
    x = z
    if cfg.get("use_label"):
        # generator.py:19-21 (with the `x` -> `z` slip at :21 fixed)
        x = torch.cat([z, F.one_hot(labels.long(), cfg["num_classes"]).to(z.dtype)], dim=1)
    W = spectral_norm(p["dense.kernel"], sn_state, "dense.u", training)        # generator.py:25
    x = x @ W + p["dense.bias"]
    x = x.reshape(-1, 4, 4, gf * 16)                                           # generator.py:26
    for i, _ in enumerate(reversed(range(_power(cfg)))):
        W = spectral_norm(p[f"block{i}.deconv.kernel"], sn_state, f"block{i}.deconv.u", training)
        x = conv2d_transpose_same(x, W, 2)                                     # generator.py:8-9
        if training or bn_stats is None:
            x = batchnorm_train(x, p[f"block{i}.bn.gamma"], p[f"block{i}.bn.beta"], bn_stats, f"block{i}.bn")
        else:
            x = batchnorm_infer(x, p[f"block{i}.bn.gamma"], p[f"block{i}.bn.beta"], bn_stats, f"block{i}.bn")
        x = F.leaky_relu(x, LRELU)                                             # generator.py:11
        if cfg.get("use_attention") and x.shape[1] in cfg["attn_dim_G"]:       # generator.py:33-34
            x = attention(x, p, sn_state, f"block{i}.attn", training, bool(cfg.get("attn_downsample")))
    x = conv2d_same(x, p["head.kernel"], None, 1)                              # generator.py:36
    return torch.tanh(x)


def discriminator_forward(p, sn_state, img, cfg, labels=None, training=True):
    """discriminator.py:13-36.  img NHWC -> patch logits [B,4,4,1] (or [B,1] with labels)."""
    x = img
    n = _power(cfg)
    for i in range(n):
        W = spectral_norm(p[f"block{i}.conv.kernel"], sn_state, f"block{i}.conv.u", training)
        x = conv2d_same(x, W, p[f"block{i}.conv.bias"], 2)                     # discriminator.py:8-9
        x = F.leaky_relu(x, LRELU)                                             # discriminator.py:10
        if cfg.get("use_attention") and x.shape[1] in cfg["attn_dim_G"]:       # discriminator.py:23-24
            x = attention(x, p, sn_state, f"block{i}.attn", training, bool(cfg.get("attn_downsample")))
    if cfg.get("use_label"):
        h = x.sum(dim=(1, 2))                                                  # discriminator.py:27
        out = h @ p["head.dense.kernel"] + p["head.dense.bias"]               # discriminator.py:28
        emb = p["head.embedding"][labels.long()]                              # discriminator.py:31
        return out + torch.sum(h * emb, dim=1, keepdim=True)                  # discriminator.py:32
    return conv2d_same(x, p["head.kernel"], p["head.bias"], 1)                 # discriminator.py:35
