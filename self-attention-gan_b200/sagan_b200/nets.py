"""Vanilla SAGAN generator / discriminator builders, same config dict as the reference.

Mirrors /root/reference/sagan/models/generator.py:7-37 (Block, get_generator) and
/root/reference/sagan/models/discriminator.py:7-36 (Block, get_discriminator) on the layer
surface of nn.py.  Two things are done at network level because they pay on B200:
  * all trainable parameters (and their gradients) live in ONE flat fp32 buffer per network, so the
    data-parallel gradient exchange and the Adam update are single-bucket operations;
  * all spectrally-normalised kernels of a network are normalised by ONE cooperative launch
    (functional.SpectralNormGroup) at the start of each training forward.
"""
import numpy as np
import torch

from . import functional as F
from . import nn
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH

LRELU = 0.1

# allocator of the flat parameter / gradient buffers: fn(numel, device) -> zeroed 1-D fp32 tensor.  The data-parallel
# trainer swaps in a symmetric-memory allocator so that peers can address the buffers over NVLink (parallel.PeerAdam).
_FLAT_ALLOC = [None]


def set_flat_allocator(fn):
    _FLAT_ALLOC[0] = fn


def _attn_pool(config):
    """Optional config key `attn_downsample` (not in the reference's dict; default off = the oracle's fixed reading):
    True selects the down-sampled attention layers.py:96,100,113 reaches for ("downsampled attn layer",
    example_configs/church64_attn.py:3): keys / values max-pooled 2x2 / stride 2."""
    return "2x2s2" if config.get("attn_downsample") else None


class GBlock(torch.nn.Module):
    """generator.py:7-12: SN(Conv2DTranspose(c,4,2,'same',no bias)) -> BN -> LeakyReLU(0.1)."""

    def __init__(self, output_channels):
        super().__init__()
        self.deconv = nn.SpectralNormalization(nn.Conv2DTranspose(output_channels, 4, 2, padding="same", use_bias=False))
        self.bn = nn.BatchNormalization(leaky_slope=LRELU)

    def forward(self, x):
        return self.bn(self.deconv(x))


class DBlock(torch.nn.Module):
    """discriminator.py:7-11: SN(Conv2D(c,4,2,'same')) -> LeakyReLU(0.1) (fused in the conv epilogue)."""

    def __init__(self, output_channels):
        super().__init__()
        self.conv = nn.SpectralNormalization(nn.Conv2D(output_channels, 4, 2, padding="same", leaky_slope=LRELU))

    def forward(self, x):
        return self.conv(x)


class Network(torch.nn.Module):
    """Common plumbing: lazy build on first call, flat parameter bucket, grouped spectral norm."""

    def __init__(self, config):
        super().__init__()
        self.config = dict(config)
        self.finalized = False
        self.flat_params = self.flat_grads = None
        self.sn_group = None
        self._sn_layers = []

    def __call__(self, inputs, training=True):
        """model(inputs, training=...) as Keras models are called (sagan/main.py:178,181-182,198-199).  `training` is
        also the learning phase of the BatchNormalization layers for the duration of the call: training=False is the
        sample-dump forward (moving statistics, no power iteration)."""
        prev = nn.learning_phase()
        nn.set_learning_phase(training)
        try:
            return super().__call__(inputs, training=training)
        finally:
            nn.set_learning_phase(prev)

    def finalize(self):
        """Re-home every parameter into one flat buffer and every SN wrapper into one group."""
        params = [p for p in self.parameters()]
        pad = lambda k: (k + 63) // 64 * 64
        total = sum(pad(p.numel()) for p in params)
        dev = params[0].device
        alloc = _FLAT_ALLOC[0] or (lambda n, d: torch.zeros(n, device=d))
        self.flat_params = alloc(total, dev)
        self.flat_grads = alloc(total, dev)
        self.param_slices = []
        off = 0
        for p in params:
            n = p.numel()
            self.flat_params[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_params[off:off + n].view(p.shape)
            p.grad = self.flat_grads[off:off + n].view(p.shape)
            self.param_slices.append((off, n))
            off += pad(n)
        self._sn_layers = [m for m in self.modules() if isinstance(m, nn.SpectralNormalization)]
        if self._sn_layers:
            ws = [m._kernel() for m in self._sn_layers]
            us = [m.u for m in self._sn_layers]
            self.sn_group = F.SpectralNormGroup(ws, us, [m.Ip for m in self._sn_layers],
                                                [m.factor for m in self._sn_layers])
            for i, m in enumerate(self._sn_layers):
                self.sn_group.v(i).copy_(m.v.reshape(-1))     # v travels with u into the network's group
                m.adopt(self.sn_group, i)
        # parameters that are not spectrally-normalised kernels (biases, BatchNorm gamma / beta, attention gamma,
        # un-normalised head kernels): read through proxies during a forward, see _normalise_all
        sn_kernels = {id(m._kernel()) for m in self._sn_layers}
        self._plain = []
        for mod in self.modules():
            if isinstance(mod, nn.Layer):
                for name in ("kernel", "bias", "gamma", "beta", "sigma"):
                    p_ = mod.__dict__.get("_parameters", {}).get(name)
                    if isinstance(p_, torch.nn.Parameter) and id(p_) not in sn_kernels:
                        self._plain.append((mod, name, p_))
        self.finalized = True

    def named_flat_parameters(self):
        return list(self.named_parameters())

    def zero_grad_flat(self):
        self.flat_grads.zero_()

    def _normalise_all(self, training):
        """One launch for all spectrally-normalised kernels; hands W_bar to each wrapper."""
        if self._plain:
            proxies = F.param_proxies([p_ for _, _, p_ in self._plain])
            for (mod, name, _), px in zip(self._plain, proxies):
                mod.__dict__.setdefault("_px", {})[name] = px
        if not self._sn_layers:
            return
        wbars = self.sn_group.normalized(update=training)
        for m, wb in zip(self._sn_layers, wbars):
            m._pending = wb

    def load_keras_weights(self, named_arrays, sn_u=None):
        """Set parameters (Keras layouts) and spectral-norm `u` vectors by oracle-style names."""
        own = dict(self.named_parameters_by_oracle_name())
        with torch.no_grad():
            for k, a in named_arrays.items():
                t = torch.as_tensor(np.asarray(a), dtype=torch.float32)
                if k not in own:
                    raise KeyError(f"unknown parameter {k}; have {sorted(own)}")
                if tuple(own[k].shape) != tuple(t.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(own[k].shape)}")
                own[k].copy_(t.to(own[k].device))
            if sn_u:
                sn = dict(self.sn_by_oracle_name())
                for k, a in sn_u.items():
                    m = sn[k]
                    m._group.u(m._index).copy_(torch.as_tensor(np.asarray(a), dtype=torch.float32).reshape(-1).to(
                        m._group.out.device))


def _attn_names(prefix, layer):
    phi, theta, g, o = layer.SN_conv
    out = []
    for nm, sn in (("phi", phi), ("theta", theta), ("g", g), ("o", o)):
        out += [(f"{prefix}.{nm}.kernel", sn.module.kernel), (f"{prefix}.{nm}.bias", sn.module.bias)]
    out.append((f"{prefix}.sigma", layer.sigma))
    return out


def _attn_sn(prefix, layer):
    return [(f"{prefix}.{nm}.u", sn) for nm, sn in zip(("phi", "theta", "g", "o"), layer.SN_conv)]


class Generator(Network):
    """generator.py:14-37."""

    def __init__(self, config):
        super().__init__(config)
        gf = config["gf_dim"]
        self.dense = nn.SpectralNormalization(nn.Dense(4 * 4 * gf * 16))                  # generator.py:25
        self.power = int(np.log2(config["img_size"] / 4))                                  # generator.py:28
        self.blocks = torch.nn.ModuleList()
        self.attn = torch.nn.ModuleDict()
        size = 4
        for i, p in enumerate(reversed(range(self.power))):
            self.blocks.append(GBlock(gf * (2 ** p)))                                      # generator.py:32
            size *= 2
            if config.get("use_attention") and size in config["attn_dim_G"]:              # generator.py:33-34
                self.attn[str(i)] = nn.AttentionLayer(pool=_attn_pool(config))
        self.head = nn.Conv2D(3, 4, 1, padding="same", use_bias=False, activation="tanh")  # generator.py:36

    def forward(self, inputs, training=True):
        z, labels = inputs if isinstance(inputs, (tuple, list)) else (inputs, None)
        cfg = self.config
        x = z
        if cfg.get("use_label"):
            # generator.py:19-21 (with the `x` -> `z` slip at :21 fixed)
            onehot = torch.nn.functional.one_hot(labels.long(), cfg["num_classes"]).to(z.dtype)
            x = torch.cat([z, onehot], dim=1).contiguous()
        if self.finalized:
            self._normalise_all(training)
        x = self.dense(x)
        x = x.reshape(-1, 4, 4, cfg["gf_dim"] * 16)                                        # generator.py:26
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if str(i) in self.attn:
                x = self.attn[str(i)](x)
        out = self.head(x)
        if not self.finalized:
            self.finalize()
        return out

    def named_parameters_by_oracle_name(self):
        out = [("dense.kernel", self.dense.module.kernel), ("dense.bias", self.dense.module.bias)]
        for i, blk in enumerate(self.blocks):
            out += [(f"block{i}.deconv.kernel", blk.deconv.module.kernel), (f"block{i}.bn.gamma", blk.bn.gamma),
                    (f"block{i}.bn.beta", blk.bn.beta)]
            if str(i) in self.attn:
                out += _attn_names(f"block{i}.attn", self.attn[str(i)])
        out.append(("head.kernel", self.head.kernel))
        return out

    def sn_by_oracle_name(self):
        out = [("dense.u", self.dense)]
        for i, blk in enumerate(self.blocks):
            out.append((f"block{i}.deconv.u", blk.deconv))
            if str(i) in self.attn:
                out += _attn_sn(f"block{i}.attn", self.attn[str(i)])
        return out


class Discriminator(Network):
    """discriminator.py:13-36."""

    def __init__(self, config):
        super().__init__(config)
        df = config["df_dim"]
        self.power = int(np.log2(config["img_size"] / 4))                                  # discriminator.py:20
        self.blocks = torch.nn.ModuleList()
        self.attn = torch.nn.ModuleDict()
        size = config["img_size"]
        for i, p in enumerate(range(self.power)):
            self.blocks.append(DBlock(df * 2 ** p))                                        # discriminator.py:22
            size //= 2
            # discriminator.py:23 reads attn_dim_G (sic); attn_dim_D is ignored, kept for drop-in parity
            if config.get("use_attention") and size in config["attn_dim_G"]:
                self.attn[str(i)] = nn.AttentionLayer(pool=_attn_pool(config))
        if config.get("use_label"):
            self.head_dense = nn.Dense(1)                                                  # discriminator.py:28
            self.embedding = None
        else:
            self.head = nn.Conv2D(1, 4, 1, padding="same")                                 # discriminator.py:35

    def forward(self, inputs, training=True):
        img, labels = inputs if isinstance(inputs, (tuple, list)) else (inputs, None)
        cfg = self.config
        if self.finalized:
            self._normalise_all(training)
        x = img
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if str(i) in self.attn:
                x = self.attn[str(i)](x)
        if cfg.get("use_label"):
            if self.embedding is None:
                c = x.shape[-1]
                self.embedding = torch.nn.Parameter(
                    (torch.rand(cfg["num_classes"], c) * 0.1 - 0.05).to(x.device))       # Keras Embedding: U(-0.05,0.05)
            h = x.sum(dim=(1, 2))                                                          # discriminator.py:27
            out = self.head_dense(h)                                                       # discriminator.py:28
            out = out + torch.sum(h * self.embedding[labels.long()], dim=1, keepdim=True)  # discriminator.py:31-32
        else:
            out = self.head(x)
        if not self.finalized:
            self.finalize()
        return out

    def named_parameters_by_oracle_name(self):
        out = []
        for i, blk in enumerate(self.blocks):
            out += [(f"block{i}.conv.kernel", blk.conv.module.kernel), (f"block{i}.conv.bias", blk.conv.module.bias)]
            if str(i) in self.attn:
                out += _attn_names(f"block{i}.attn", self.attn[str(i)])
        if self.config.get("use_label"):
            out += [("head.dense.kernel", self.head_dense.kernel), ("head.dense.bias", self.head_dense.bias),
                    ("head.embedding", self.embedding)]
        else:
            out += [("head.kernel", self.head.kernel), ("head.bias", self.head.bias)]
        return out

    def sn_by_oracle_name(self):
        out = []
        for i, blk in enumerate(self.blocks):
            out.append((f"block{i}.conv.u", blk.conv))
            if str(i) in self.attn:
                out += _attn_sn(f"block{i}.attn", self.attn[str(i)])
        return out


# ------------------------------------------------------------------------------------------------ residual topologies
class ResGBlock(torch.nn.Module):
    """/root/reference/models/generator.py:6-21: BN -> ReLU -> SN(Conv2DTranspose(c,3,2,'same')) -> BN -> ReLU ->
    SN(Conv2D(c,3,1,'same')), plus the shortcut SN(Conv2DTranspose(c,3,2,'same')) of the block INPUT."""

    def __init__(self, c):
        super().__init__()
        self.pre_bn = nn.BatchNormalization(leaky_slope=0.0)                                        # :7-8 (BN + ReLU fused)
        self.deconv1 = nn.SpectralNormalization(nn.Conv2DTranspose(c, 3, 2, padding="same"))        # :11-12
        self.mid_bn = nn.BatchNormalization(leaky_slope=0.0)                                        # :13-14
        self.conv2 = nn.SpectralNormalization(nn.Conv2D(c, 3, 1, padding="same"))                   # :15-16
        self.deconv_sc = nn.SpectralNormalization(nn.Conv2DTranspose(c, 3, 2, padding="same"))      # :18-19

    def forward(self, x):
        h = self.conv2(self.mid_bn(self.deconv1(self.pre_bn(x))))
        return nn.add([self.deconv_sc(x), h])                                                       # :21


class ResGenerator(Network):
    """/root/reference/models/generator.py:23-43 (class-conditional; block count from img_size, 5 at 128x128)."""

    def __init__(self, config):
        super().__init__(config)
        gf = config["gf_dim"]
        self.power = int(np.log2(config["img_size"] / 4))
        self.dense = nn.SpectralNormalization(nn.Dense(4 * 4 * gf * 2 ** (self.power - 1)))         # :28
        self.blocks = torch.nn.ModuleList()
        self.attn = torch.nn.ModuleDict()
        size = 4
        for i in range(self.power):
            self.blocks.append(ResGBlock(gf * 2 ** (self.power - 1 - i)))                           # :31-37
            size *= 2
            if size in config.get("attn_dim_G", [32]):                                              # :34 (32x32)
                self.attn[str(i)] = nn.Attention_Layer(pool=_attn_pool(config))
        self.final_bn = nn.BatchNormalization(leaky_slope=0.0)                                      # :39-40
        self.final_conv = nn.SpectralNormalization(nn.Conv2D(3, 3, 1, padding="same", activation="tanh"))   # :41-42

    def forward(self, inputs, training=True):
        z, labels = inputs
        cfg = self.config
        onehot = torch.nn.functional.one_hot(labels.long(), cfg["num_classes"]).to(z.dtype)         # :26
        x = torch.cat([z, onehot], dim=1).contiguous()                                              # :27
        if self.finalized:
            self._normalise_all(training)
        x = self.dense(x)
        x = x.reshape(-1, 4, 4, x.shape[1] // 16)                                                   # :29
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if str(i) in self.attn:
                x = self.attn[str(i)](x)
        out = self.final_conv(self.final_bn(x))
        if not self.finalized:
            self.finalize()
        return out

    def named_parameters_by_oracle_name(self):
        out = [("dense.kernel", self.dense.module.kernel), ("dense.bias", self.dense.module.bias)]
        for i, b in enumerate(self.blocks):
            out += [(f"block{i}.pre.bn.gamma", b.pre_bn.gamma), (f"block{i}.pre.bn.beta", b.pre_bn.beta),
                    (f"block{i}.deconv1.kernel", b.deconv1.module.kernel), (f"block{i}.deconv1.bias", b.deconv1.module.bias),
                    (f"block{i}.mid.bn.gamma", b.mid_bn.gamma), (f"block{i}.mid.bn.beta", b.mid_bn.beta),
                    (f"block{i}.conv2.kernel", b.conv2.module.kernel), (f"block{i}.conv2.bias", b.conv2.module.bias),
                    (f"block{i}.deconv_sc.kernel", b.deconv_sc.module.kernel), (f"block{i}.deconv_sc.bias", b.deconv_sc.module.bias)]
            if str(i) in self.attn:
                out += _attn_names(f"block{i}.attn", self.attn[str(i)])
        out += [("final.bn.gamma", self.final_bn.gamma), ("final.bn.beta", self.final_bn.beta),
                ("final.conv.kernel", self.final_conv.module.kernel), ("final.conv.bias", self.final_conv.module.bias)]
        return out

    def sn_by_oracle_name(self):
        out = [("dense.u", self.dense)]
        for i, b in enumerate(self.blocks):
            out += [(f"block{i}.deconv1.u", b.deconv1), (f"block{i}.conv2.u", b.conv2), (f"block{i}.deconv_sc.u", b.deconv_sc)]
            if str(i) in self.attn:
                out += _attn_sn(f"block{i}.attn", self.attn[str(i)])
        out.append(("final.conv.u", self.final_conv))
        return out


class ResDBlock(torch.nn.Module):
    """/root/reference/models/discriminator.py:19-38 (`first=True`: Optimized_Block, :6-17, no activation on the image)."""

    def __init__(self, c, downsample=True, first=False):
        super().__init__()
        s = 2 if downsample else 1
        self.first = first
        self.relu = nn.ReLU()
        self.conv1 = nn.SpectralNormalization(nn.Conv2D(c, 3, 1, padding="same", leaky_slope=0.0))   # conv + the ReLU that follows
        self.conv2 = nn.SpectralNormalization(nn.Conv2D(c, 3, s, padding="same"))
        self.conv_sc = nn.SpectralNormalization(nn.Conv2D(c, 3, s, padding="same"))

    def forward(self, x):
        a = x if self.first else self.relu(x)                                                       # :22, :32
        return nn.add([self.conv_sc(a), self.conv2(self.conv1(a))])                                 # :17, :36


class ResDiscriminator(Network):
    """/root/reference/models/discriminator.py:40-57: projection discriminator with a spectrally-normalised Embedding."""

    def __init__(self, config):
        super().__init__(config)
        df = config["df_dim"]
        self.power = int(np.log2(config["img_size"] / 4))
        self.opt = ResDBlock(df, first=True)                                                        # :44
        chans = [df * 2 ** p for p in range(1, self.power)] + [df * 2 ** (self.power - 1)]          # :45-51
        self.blocks = torch.nn.ModuleList()
        self.attn = torch.nn.ModuleDict()
        size = config["img_size"] // 2
        for i, c in enumerate(chans):
            last = i == len(chans) - 1
            self.blocks.append(ResDBlock(c, downsample=not last))
            if not last:
                size //= 2
                if size in config.get("attn_dim_G", [32]):                                          # :46 (32x32)
                    self.attn[str(i)] = nn.Attention_Layer(pool=_attn_pool(config))
        self.relu = nn.ReLU()
        self.head_dense = nn.SpectralNormalization(nn.Dense(1))                                     # :52
        self.embedding = nn.SpectralNormalization(nn.Embedding(config["num_classes"], chans[-1]))   # :53-54

    def forward(self, inputs, training=True):
        img, labels = inputs
        if self.finalized:
            self._normalise_all(training)
        x = self.opt(img)
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if str(i) in self.attn:
                x = self.attn[str(i)](x)
        h = self.relu(x).sum(dim=(1, 2))                                                            # :49-50
        out = self.head_dense(h) + torch.sum(h * self.embedding(labels), dim=1, keepdim=True)       # :52-55
        if not self.finalized:
            self.finalize()
        return out

    def named_parameters_by_oracle_name(self):
        out = []
        for nm, b in [("opt", self.opt)] + [(f"block{i}", b) for i, b in enumerate(self.blocks)]:
            out += [(f"{nm}.conv1.kernel", b.conv1.module.kernel), (f"{nm}.conv1.bias", b.conv1.module.bias),
                    (f"{nm}.conv2.kernel", b.conv2.module.kernel), (f"{nm}.conv2.bias", b.conv2.module.bias),
                    (f"{nm}.conv_sc.kernel", b.conv_sc.module.kernel), (f"{nm}.conv_sc.bias", b.conv_sc.module.bias)]
            if nm != "opt" and nm[5:] in self.attn:
                out += _attn_names(f"{nm}.attn", self.attn[nm[5:]])
        out += [("head.dense.kernel", self.head_dense.module.kernel), ("head.dense.bias", self.head_dense.module.bias),
                ("head.embedding", self.embedding.module.kernel)]
        return out

    def sn_by_oracle_name(self):
        out = []
        for nm, b in [("opt", self.opt)] + [(f"block{i}", b) for i, b in enumerate(self.blocks)]:
            out += [(f"{nm}.conv1.u", b.conv1), (f"{nm}.conv2.u", b.conv2), (f"{nm}.conv_sc.u", b.conv_sc)]
            if nm != "opt" and nm[5:] in self.attn:
                out += _attn_sn(f"{nm}.attn", self.attn[nm[5:]])
        out += [("head.dense.u", self.head_dense), ("head.embedding.u", self.embedding)]
        return out


def get_res_generator(config):
    """The residual generator of /root/reference/models/generator.py:23 on the config dict of the `sagan/` tree
    (`model: 'resnet'`, sagan/main.py:104-107 -- disabled there with "TODO: fix resnet model")."""
    return ResGenerator(config)


def get_res_discriminator(config):
    """/root/reference/models/discriminator.py:40."""
    return ResDiscriminator(config)


def get_generator(config):
    """generator.py:14.  Returns a callable model: model([z, labels], training=True) -> images NHWC."""
    return Generator(config)


def get_discriminator(config):
    """discriminator.py:13.  model([images, labels], training=True) -> patch logits [B,4,4,1] (or [B,1])."""
    return Discriminator(config)
