"""Oracle: the residual SAGAN generator / discriminator (torch-CPU, un-fused).  TEST INFRASTRUCTURE.

SURVEY.md §8f row 3.  Follows the reference's legacy top-level models, the only well-formed residual topologies it holds
(the `sagan/models` res variants are disabled and reference undefined names):
  /root/reference/models/generator.py:6-43       (Block, get_generator: 128x128 class-conditional, attention at 32x32)
  /root/reference/models/discriminator.py:6-57   (Optimized_Block, Block, get_discriminator: projection head with a
                                                  spectrally-normalised Embedding)
with the layer math of /root/reference/layers.py (oracle.nets.spectral_norm / attention).  The block count follows
log2(img_size / 4) so that small grids can be tested; img_size = 128 reproduces the reference exactly.
Keras defaults restated: Conv2D / Conv2DTranspose / Dense have a bias; BatchNormalization eps 1e-3; ReLU.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import nets
from .nets import attention, batchnorm_train, conv2d_same, conv2d_transpose_same, spectral_norm


def _power(cfg):
    return int(np.log2(cfg["img_size"] / 4))


def _attn_at(cfg):
    return cfg.get("attn_dim_G", [32])                     # models/generator.py:34, models/discriminator.py:42: 32x32


# ----------------------------------------------------------------------------- specs
def res_generator_spec(cfg):
    gf, P = cfg["gf_dim"], _power(cfg)
    cin = gf * 2 ** (P - 1)                                # gf * 16 at 128x128 (models/generator.py:28)
    spec = [("dense.kernel", (cfg["z_dim"] + cfg["num_classes"], 4 * 4 * cin)), ("dense.bias", (4 * 4 * cin,))]
    size = 4
    for i in range(P):
        c = gf * 2 ** (P - 1 - i)                          # 16, 8, 4, 2, 1 x gf (models/generator.py:31-37)
        spec += [(f"block{i}.pre.bn.gamma", (cin,)), (f"block{i}.pre.bn.beta", (cin,)),
                 (f"block{i}.deconv1.kernel", (3, 3, c, cin)), (f"block{i}.deconv1.bias", (c,)),
                 (f"block{i}.mid.bn.gamma", (c,)), (f"block{i}.mid.bn.beta", (c,)),
                 (f"block{i}.conv2.kernel", (3, 3, c, c)), (f"block{i}.conv2.bias", (c,)),
                 (f"block{i}.deconv_sc.kernel", (3, 3, c, cin)), (f"block{i}.deconv_sc.bias", (c,))]
        size *= 2
        if size in _attn_at(cfg):
            spec += nets._attn_spec(f"block{i}.attn", c)
        cin = c
    spec += [("final.bn.gamma", (cin,)), ("final.bn.beta", (cin,)), ("final.conv.kernel", (3, 3, cin, 3)),
             ("final.conv.bias", (3,))]
    return spec


def res_discriminator_spec(cfg):
    df, P = cfg["df_dim"], _power(cfg)
    spec = [("opt.conv1.kernel", (3, 3, 3, df)), ("opt.conv1.bias", (df,)), ("opt.conv2.kernel", (3, 3, df, df)),
            ("opt.conv2.bias", (df,)), ("opt.conv_sc.kernel", (3, 3, 3, df)), ("opt.conv_sc.bias", (df,))]
    cin, size = df, cfg["img_size"] // 2
    chans = [df * 2 ** p for p in range(1, P)] + [df * 2 ** (P - 1)]      # 2, 4, 8, 16, 16 x df (models/discriminator.py:41-47)
    for i, c in enumerate(chans):
        spec += [(f"block{i}.conv1.kernel", (3, 3, cin, c)), (f"block{i}.conv1.bias", (c,)),
                 (f"block{i}.conv2.kernel", (3, 3, c, c)), (f"block{i}.conv2.bias", (c,)),
                 (f"block{i}.conv_sc.kernel", (3, 3, cin, c)), (f"block{i}.conv_sc.bias", (c,))]
        if i < len(chans) - 1:
            size //= 2
            if size in _attn_at(cfg):
                spec += nets._attn_spec(f"block{i}.attn", c)
        cin = c
    spec += [("head.dense.kernel", (cin, 1)), ("head.dense.bias", (1,)), ("head.embedding", (cfg["num_classes"], cin))]
    return spec


def res_sn_keys(spec):
    """Every wrapped kernel is spectrally normalised here, the output conv, the head's Dense and the Embedding included
    (models/generator.py:41-42, models/discriminator.py:52-54)."""
    out = OrderedDict()
    for name, shape in spec:
        if name.endswith(".kernel") or name.endswith("embedding"):
            R = shape[-1]
            out[name.rsplit(".", 1)[0] + ".u" if name.endswith(".kernel") else name + ".u"] = (R, int(np.prod(shape)) // R)
    return out


def init_res_sn_state(spec, seed=1, dtype=torch.float64):
    rng = np.random.Generator(np.random.PCG64(seed))
    st = OrderedDict()
    for key, (R, K) in res_sn_keys(spec).items():
        st[key] = nets.l2n(torch.tensor(rng.standard_normal((1, R)), dtype=dtype))
    return st


# ----------------------------------------------------------------------------- models
def _sn(p, sn, name, training):
    return spectral_norm(p[name + ".kernel"], sn, name + ".u", training)


def res_generator_forward(p, sn, z, labels, cfg, training=True, bn_stats=None):
    """models/generator.py:23-43."""
    P = _power(cfg)
    x = torch.cat([z, F.one_hot(labels.long(), cfg["num_classes"]).to(z.dtype)], dim=1)     # :26-27
    x = x @ _sn(p, sn, "dense", training) + p["dense.bias"]                                 # :28
    x = x.reshape(-1, 4, 4, x.shape[1] // 16)                                               # :29
    for i in range(P):
        inp = x
        h = F.relu(batchnorm_train(inp, p[f"block{i}.pre.bn.gamma"], p[f"block{i}.pre.bn.beta"], bn_stats, f"block{i}.pre.bn"))  # :7-8
        h = conv2d_transpose_same(h, _sn(p, sn, f"block{i}.deconv1", training), 2) + p[f"block{i}.deconv1.bias"]            # :11-12
        h = F.relu(batchnorm_train(h, p[f"block{i}.mid.bn.gamma"], p[f"block{i}.mid.bn.beta"], bn_stats, f"block{i}.mid.bn"))   # :13-14
        h = conv2d_same(h, _sn(p, sn, f"block{i}.conv2", training), p[f"block{i}.conv2.bias"], 1)                           # :15-16
        sc = conv2d_transpose_same(inp, _sn(p, sn, f"block{i}.deconv_sc", training), 2) + p[f"block{i}.deconv_sc.bias"]     # :18-19
        x = sc + h                                                                                                          # :21
        if f"block{i}.attn.sigma" in p:
            x = attention(x, p, sn, f"block{i}.attn", training, bool(cfg.get("attn_downsample")))                           # :34
    x = F.relu(batchnorm_train(x, p["final.bn.gamma"], p["final.bn.beta"], bn_stats, "final.bn"))                           # :39-40
    return torch.tanh(conv2d_same(x, _sn(p, sn, "final.conv", training), p["final.conv.bias"], 1))                          # :41-42


def res_discriminator_forward(p, sn, img, labels, cfg, training=True):
    """models/discriminator.py:40-57."""
    # Optimized_Block, :6-17
    h = F.relu(conv2d_same(img, _sn(p, sn, "opt.conv1", training), p["opt.conv1.bias"], 1))
    h = conv2d_same(h, _sn(p, sn, "opt.conv2", training), p["opt.conv2.bias"], 2)
    x = conv2d_same(img, _sn(p, sn, "opt.conv_sc", training), p["opt.conv_sc.bias"], 2) + h
    nb = _power(cfg)
    for i in range(nb):
        stride = 2 if i < nb - 1 else 1                                                     # :47 downsample=False
        a = F.relu(x)                                                                       # :22 / :32 (the same tensor)
        h = F.relu(conv2d_same(a, _sn(p, sn, f"block{i}.conv1", training), p[f"block{i}.conv1.bias"], 1))   # :23-26
        h = conv2d_same(h, _sn(p, sn, f"block{i}.conv2", training), p[f"block{i}.conv2.bias"], stride)      # :27-28
        sc = conv2d_same(a, _sn(p, sn, f"block{i}.conv_sc", training), p[f"block{i}.conv_sc.bias"], stride)  # :32-34
        x = sc + h                                                                          # :36
        if f"block{i}.attn.sigma" in p:
            x = attention(x, p, sn, f"block{i}.attn", training, bool(cfg.get("attn_downsample")))           # :42
    h = F.relu(x).sum(dim=(1, 2))                                                           # :49-50
    out = h @ _sn(p, sn, "head.dense", training) + p["head.dense.bias"]                     # :52
    emb = spectral_norm(p["head.embedding"], sn, "head.embedding.u", training)[labels.long()]   # :53-54
    return out + torch.sum(h * emb, dim=1, keepdim=True)                                    # :55
