"""Parity against vectors produced by the REFERENCE'S OWN CODE (tests/golden/reference_layers.npz).

`tests/golden/make_reference_vectors.py` imports /root/reference/layers.py, sagan/models/generator.py,
sagan/models/discriminator.py and the hinge-loss definitions of sagan/main.py UNMODIFIED (with a float64 numpy stand-in
for the `tensorflow` module, since TensorFlow cannot be installed) and records what they compute.  Here:

* not-gpu tests hold the ORACLE (oracle/sn.py, attention.py, train.py, nets.py) to those numbers at fp64 round-off
  (1e-12): this is what pins the oracle to the reference rather than to our reading of it;
* gpu tests hold the CUDA kernels, called through the C ABI, to the same numbers at the north_star tolerances
  (sigma / u / v 1e-5 relative; attention and model outputs 1e-5 relative L2 in the strict mode).

Nothing here reads /root/reference or the stand-in at run time: only the committed fixture.
"""
import os

import numpy as np
import pytest
import torch

from oracle import attention as oattn
from oracle import nets as onets
from oracle import sn as osn
from oracle import train as otrain

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_layers.npz")
FP64_TOL = 1e-12
STRICT_TOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLD)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def _factor(ref, tag):
    f = float(ref[f"sn_{tag}_factor"])
    return f if f else None


# ------------------------------------------------------------------------------------------------ oracle == reference
def test_oracle_l2normalize_matches_reference(ref):
    for i in range(3):
        assert rel_l2(osn.l2normalize(ref[f"l2n_{i}_x"]), ref[f"l2n_{i}_y"]) < FP64_TOL
    assert np.array_equal(osn.l2normalize(np.zeros((1, 4))), ref["l2n_zero_y"])


def test_oracle_power_iteration_matches_reference_update_uv(ref):
    """layers.py:30-38,50-68 executed by the reference on Dense / Conv2D / Conv2DTranspose / 1x1 kernels in their Keras
    layouts: u, v after one and two calls, sigma, W / sigma, the [R, K] the kernel is reshaped to."""
    for tag in ref["sn_tags"]:
        p = f"sn_{tag}_"
        W, u0, Ip, factor = ref[p + "W"], ref[p + "u0"], int(ref[p + "Ip"]), _factor(ref, tag)
        assert tuple(osn.matricize(W).shape) == tuple(ref[p + "Wmat_shape"]), tag
        R, K = ref[p + "Wmat_shape"]
        assert R == W.shape[-1] and u0.shape == (1, R) and ref[p + "v0"].shape == (1, K), tag   # _make_param shapes
        assert abs(np.linalg.norm(u0) - 1) < 1e-12 and abs(np.linalg.norm(ref[p + "v0"]) - 1) < 1e-12
        u1, v1, s1, Wb1 = osn.power_iteration(W, u0, Ip, factor)
        assert rel_l2(u1, ref[p + "u1"]) < FP64_TOL and rel_l2(v1, ref[p + "v1"]) < FP64_TOL, tag
        assert abs(s1 - ref[p + "sigma1"]) < FP64_TOL * abs(s1), tag
        assert rel_l2(Wb1, ref[p + "Wbar1"]) < FP64_TOL and Wb1.shape == W.shape, tag
        u2, v2, s2, _ = osn.power_iteration(W, u1, Ip, factor)       # u persists into the next call
        assert rel_l2(u2, ref[p + "u2"]) < FP64_TOL and rel_l2(v2, ref[p + "v2"]) < FP64_TOL, tag
        assert abs(s2 - ref[p + "sigma2"]) < FP64_TOL * abs(s2), tag
    with pytest.raises(ValueError) as e:
        osn.power_iteration(np.ones((2, 2)), np.ones((1, 2)), Ip=0)
    assert str(e.value) == str(ref["sn_ip0_error"])


def test_oracle_nets_spectral_norm_matches_reference(ref):
    """The torch restatement inside oracle/nets.py (what the model-level oracle uses) against the same vectors."""
    for tag in ref["sn_tags"]:
        p = f"sn_{tag}_"
        st = {"k": torch.tensor(ref[p + "u0"])}
        Wb = onets.spectral_norm(torch.tensor(ref[p + "W"]), st, "k", True, int(ref[p + "Ip"]), _factor(ref, tag))
        assert rel_l2(Wb.numpy(), ref[p + "Wbar1"]) < FP64_TOL and rel_l2(st["k"].numpy(), ref[p + "u1"]) < FP64_TOL


def _attn_args(ref, p):
    return dict(Wphi=ref[p + "Wphi"], bphi=ref[p + "bphi"], Wtheta=ref[p + "Wtheta"], btheta=ref[p + "btheta"],
                Wg=ref[p + "Wg"], bg=ref[p + "bg"], Wo=ref[p + "Wo"], bo=ref[p + "bo"], gamma=float(ref[p + "gamma"]))


def test_oracle_attention_matches_reference_call(ref):
    """Attention_Layer.call (layers.py:93-120) run by the reference with the ill-formed MaxPool2D(2, 1) replaced by
    the identity, at C = 8 where its raw reshape of phi is a true transpose."""
    for C, d1, d2, dv, co in ref["attn_split"]:
        assert (d1, dv) == oattn.channel_split(int(C)) and d2 == d1 and co == C
    for i in range(int(ref["attn_cases"])):
        p = f"attn_{i}_"
        X = ref[p + "X"]
        B, H, W, C = X.shape
        Y = oattn.forward(X.reshape(B, H * W, C), **_attn_args(ref, p)).reshape(X.shape)
        assert rel_l2(Y, ref[p + "Y"]) < FP64_TOL
        if float(ref[p + "gamma"]) == 0.0:
            assert np.array_equal(ref[p + "Y"], X)                    # zero-initialised gamma: identity block


def test_reference_literal_pooled_call_is_not_attention(ref):
    """Documents SURVEY.md appendix A.5 with the reference's own numbers: at the one shape family where the literal
    MaxPool2D(2, 1) + reshape goes through (2x2 map, B = 4), the four samples' pooled keys / values are folded into ONE
    [d, 4] / [4, dv] matrix shared (broadcast) by every sample.  Reproduced here in numpy from the recorded weights."""
    p = "attnlit_"
    X, a = ref[p + "X"], _attn_args(ref, p)
    B, H, W, C = X.shape
    Xf = X.reshape(B, H * W, C)
    phi = (Xf @ a["Wphi"] + a["bphi"]).max(axis=1)                    # 2x2 window, stride 1 on a 2x2 map: one value
    g = (Xf @ a["Wg"] + a["bg"]).max(axis=1)
    theta = Xf @ a["Wtheta"] + a["btheta"]
    phi_m = phi.reshape(1, C // 8, H * W)                              # the batch axis is folded into the tokens
    g_m = g.reshape(1, H * W, C // 2)
    S = theta @ phi_m
    P = np.exp(S - S.max(-1, keepdims=True))
    P /= P.sum(-1, keepdims=True)
    Y = Xf + a["gamma"] * ((P @ g_m) @ a["Wo"] + a["bo"])
    assert rel_l2(Y.reshape(X.shape), ref[p + "Y"]) < FP64_TOL
    assert rel_l2(oattn.forward(Xf, **a).reshape(X.shape), ref[p + "Y"]) > 1e-3     # and it is not the paper form


def test_oracle_hinge_losses_match_reference(ref):
    real, fake = torch.tensor(ref["hinge_real"]), torch.tensor(ref["hinge_fake"])
    assert np.array_equal(otrain.hinge_loss_d(real, fake).numpy(), ref["hinge_d"])
    assert np.array_equal(otrain.hinge_loss_g(fake).numpy(), ref["hinge_g"])


# the configuration the builders were run with (make_reference_vectors.section_builders)
BUILDER_CFG = dict(z_dim=32, gf_dim=4, df_dim=4, img_size=64, use_attention=False, attn_dim_G=[32, 64],
                   attn_dim_D=[8, 4], use_label=False, batch_size=2, num_classes=1)


def _gen_params(ref):
    """Maps the creation-ordered layers the reference's get_generator built onto the oracle's parameter names, checking
    on the way that each layer is what generator.py:7-37 says (kind, width, 4x4 / stride 2 / same, bias flag, slope)."""
    desc = list(ref["gen_layers"])
    gf, n = BUILDER_CFG["gf_dim"], 4
    assert desc[0] == f"Dense sn units={4 * 4 * gf * 16} use_bias=True activation=None"
    p = {"dense.kernel": ref["gen_L0_kernel"], "dense.bias": ref["gen_L0_bias"]}
    u = {"dense.u": ref["gen_L0_u"]}
    for i in range(n):
        cout = gf * 2 ** (n - 1 - i)
        j = 1 + 3 * i
        assert desc[j] == (f"Conv2DTranspose sn filters={cout} kernel_size=(4, 4) strides=(2, 2) padding=same "
                           "use_bias=False activation=None")
        assert desc[j + 1] == "BatchNormalization plain epsilon=0.001 momentum=0.99"
        assert desc[j + 2] == "LeakyReLU plain alpha=0.1"
        p[f"block{i}.deconv.kernel"] = ref[f"gen_L{j}_kernel"]
        p[f"block{i}.bn.gamma"], p[f"block{i}.bn.beta"] = ref[f"gen_L{j + 1}_gamma"], ref[f"gen_L{j + 1}_beta"]
        u[f"block{i}.deconv.u"] = ref[f"gen_L{j}_u"]
    assert desc[13] == "Conv2D plain filters=3 kernel_size=(4, 4) strides=(1, 1) padding=same use_bias=False activation=tanh"
    assert len(desc) == 14
    p["head.kernel"] = ref["gen_L13_kernel"]
    return p, u


def _dis_params(ref, prefix, cfg):
    desc = list(ref[prefix + "layers"])
    df, n = cfg["df_dim"], 4
    p, u = {}, {}
    for i in range(n):
        j = 2 * i
        assert desc[j] == (f"Conv2D sn filters={df * 2 ** i} kernel_size=(4, 4) strides=(2, 2) padding=same "
                           "use_bias=True activation=None")
        assert desc[j + 1] == "LeakyReLU plain alpha=0.1"
        p[f"block{i}.conv.kernel"], p[f"block{i}.conv.bias"] = ref[f"{prefix}L{j}_kernel"], ref[f"{prefix}L{j}_bias"]
        u[f"block{i}.conv.u"] = ref[f"{prefix}L{j}_u"]
    if cfg["use_label"]:
        assert desc[8:] == ["Dense plain units=1 use_bias=True activation=None", "Embedding plain"]
        p["head.dense.kernel"], p["head.dense.bias"] = ref[prefix + "L8_kernel"], ref[prefix + "L8_bias"]
        p["head.embedding"] = ref[prefix + "L9_embeddings"]
    else:
        assert desc[8:] == ["Conv2D plain filters=1 kernel_size=(4, 4) strides=(1, 1) padding=same use_bias=True activation=None"]
        p["head.kernel"], p["head.bias"] = ref[prefix + "L8_kernel"], ref[prefix + "L8_bias"]
    return p, u


def _t64(d):
    return {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in d.items()}


def test_oracle_generator_matches_reference_builder(ref):
    """get_generator (sagan/models/generator.py:14-37) run by the reference on a concrete z: same layers in the same
    order as oracle.nets.generator_spec, same image out of oracle.nets.generator_forward."""
    p, u = _gen_params(ref)
    assert {k: tuple(v.shape) for k, v in p.items()} == dict(onets.generator_spec(BUILDER_CFG))
    assert list(u) == list(onets.sn_keys(onets.generator_spec(BUILDER_CFG)))
    st = _t64(u)
    img = onets.generator_forward(_t64(p), st, torch.tensor(ref["gen_in"]), BUILDER_CFG)
    assert rel_l2(img.numpy(), ref["gen_out"]) < 1e-11
    for k in u:                                                       # kernels were pre-scaled to sigma = 1
        assert abs(float(st[k + ":sigma"]) - 1) < 1e-10          # (the 1e-12 eps of l2normalize is not scale-free)


@pytest.mark.parametrize("prefix", ["dis_", "disc_"])
def test_oracle_discriminator_matches_reference_builder(ref, prefix):
    """get_discriminator (sagan/models/discriminator.py:13-36): the patch-logit head and the projection head."""
    cfg = dict(BUILDER_CFG, use_label=True, num_classes=10) if prefix == "disc_" else BUILDER_CFG
    p, u = _dis_params(ref, prefix, cfg)
    assert {k: tuple(v.shape) for k, v in p.items()} == dict(onets.discriminator_spec(cfg))
    assert list(u) == list(onets.sn_keys(onets.discriminator_spec(cfg)))
    labels = torch.tensor(ref[prefix + "labels"].astype(np.int64))
    out = onets.discriminator_forward(_t64(p), _t64(u), torch.tensor(ref[prefix + "in"]), cfg, labels=labels)
    assert out.shape == ref[prefix + "out"].shape
    assert rel_l2(out.numpy(), ref[prefix + "out"]) < 1e-11


RES_CFG = dict(model="resnet", z_dim=128, gf_dim=2, df_dim=4, img_size=128, num_classes=5, use_label=True,
               use_attention=True, attn_dim_G=[32], batch_size=2)


class _Cursor:
    """Walks the creation-ordered layer list the reference's builder produced, asserting each entry on the way."""

    def __init__(self, ref, prefix):
        self.ref, self.prefix, self.desc, self.i = ref, prefix, list(ref[prefix + "layers"]), 0

    def take(self, expect):
        assert self.desc[self.i] == expect, (self.i, self.desc[self.i], expect)
        self.i += 1
        return self.i - 1

    def arr(self, idx, name):
        return self.ref[f"{self.prefix}L{idx}_{name}"]

    def conv(self, kind, filters, k, s, name, p, u, inner=False, act=None, padding="same"):
        idx = self.take(f"{kind} sn {'inner ' if inner else ''}filters={filters} kernel_size=({k}, {k}) strides=({s}, {s}) "
                        f"padding={padding} use_bias=True activation={act}")
        p[name + ".kernel"], p[name + ".bias"], u[name + ".u"] = self.arr(idx, "kernel"), self.arr(idx, "bias"), self.arr(idx, "u")

    def bn(self, name, p):
        idx = self.take("BatchNormalization plain epsilon=0.001 momentum=0.99")
        p[name + ".gamma"], p[name + ".beta"] = self.arr(idx, "gamma"), self.arr(idx, "beta")

    def attention(self, prefix, C, p, u):
        idx = self.take("Attention_Layer plain")
        p[prefix + ".sigma"] = self.arr(idx, "sigma")
        for nm, f in (("phi", C // 8), ("theta", C // 8), ("g", C // 2), ("o", C)):       # layers.py:82-85 order
            self.conv("Conv2D", f, 1, 1, f"{prefix}.{nm}", p, u, inner=True, padding="valid")

    def done(self):
        return self.i == len(self.desc)


def _res_gen_params(ref):
    """models/generator.py:6-43 as the reference built it -> oracle.resnets parameter names."""
    c, p, u = _Cursor(ref, "rgen_"), {}, {}
    gf = RES_CFG["gf_dim"]
    c.take("Concatenate plain")                                                              # :27
    idx = c.take(f"Dense sn units={4 * 4 * gf * 16} use_bias=True activation=None")         # :28
    p["dense.kernel"], p["dense.bias"], u["dense.u"] = c.arr(idx, "kernel"), c.arr(idx, "bias"), c.arr(idx, "u")
    for i, mult in enumerate((16, 8, 4, 2, 1)):                                              # :31-37
        ch = gf * mult
        c.bn(f"block{i}.pre.bn", p), c.take("ReLU plain")                                    # :7-8
        c.conv("Conv2DTranspose", ch, 3, 2, f"block{i}.deconv1", p, u)                       # :11-12
        c.bn(f"block{i}.mid.bn", p), c.take("ReLU plain")                                    # :13-14
        c.conv("Conv2D", ch, 3, 1, f"block{i}.conv2", p, u)                                  # :15-16
        c.conv("Conv2DTranspose", ch, 3, 2, f"block{i}.deconv_sc", p, u)                     # :18-19
        if i == 2:
            c.attention(f"block{i}.attn", ch, p, u)                                          # :34 (after the 32x32 block)
    c.bn("final.bn", p), c.take("ReLU plain")                                                # :39-40
    c.conv("Conv2D", 3, 3, 1, "final.conv", p, u, act="tanh")                                # :41-42
    assert c.done()
    return p, u


def _res_dis_params(ref):
    """models/discriminator.py:6-57 -> oracle.resnets parameter names."""
    c, p, u = _Cursor(ref, "rdis_"), {}, {}
    df = RES_CFG["df_dim"]
    c.conv("Conv2D", df, 3, 1, "opt.conv1", p, u), c.take("ReLU plain")                      # :7-9
    c.conv("Conv2D", df, 3, 2, "opt.conv2", p, u)                                            # :11-12
    c.conv("Conv2D", df, 3, 2, "opt.conv_sc", p, u)                                          # :14-15
    for i, (mult, stride) in enumerate(((2, 2), (4, 2), (8, 2), (16, 2), (16, 1))):          # :41-47
        ch = df * mult
        c.take("ReLU plain"), c.conv("Conv2D", ch, 3, 1, f"block{i}.conv1", p, u)            # :22-24
        c.take("ReLU plain"), c.conv("Conv2D", ch, 3, stride, f"block{i}.conv2", p, u)       # :26-28
        c.take("ReLU plain"), c.conv("Conv2D", ch, 3, stride, f"block{i}.conv_sc", p, u)     # :30-32
        if i == 0:
            c.attention(f"block{i}.attn", ch, p, u)                                          # :42
    c.take("ReLU plain")                                                                     # :49
    idx = c.take("Dense sn units=1 use_bias=True activation=None")                           # :52
    p["head.dense.kernel"], p["head.dense.bias"], u["head.dense.u"] = c.arr(idx, "kernel"), c.arr(idx, "bias"), c.arr(idx, "u")
    idx = c.take("Embedding sn")                                                             # :53-54
    p["head.embedding"], u["head.embedding.u"] = c.arr(idx, "embeddings"), c.arr(idx, "u")
    assert c.done()
    return p, u


def test_oracle_residual_generator_matches_reference_builder(ref):
    """models/generator.py:23-43 run by the reference (attention at 32x32 with C = 8 included): same layers in the
    same order as oracle.resnets.res_generator_spec, same image."""
    from oracle import resnets as ores
    p, u = _res_gen_params(ref)
    assert {k: tuple(v.shape) for k, v in p.items()} == dict(ores.res_generator_spec(RES_CFG))
    assert sorted(u) == sorted(ores.res_sn_keys(ores.res_generator_spec(RES_CFG)))
    labels = torch.tensor(ref["rgen_labels"].astype(np.int64))
    img = ores.res_generator_forward(_t64(p), _t64(u), torch.tensor(ref["rgen_in"]), labels, RES_CFG)
    assert img.shape == ref["rgen_out"].shape == (2, 128, 128, 3)                            # test/test_generator.py:26
    assert rel_l2(img.numpy(), ref["rgen_out"]) < 1e-10


def test_oracle_residual_discriminator_matches_reference_builder(ref):
    from oracle import resnets as ores
    p, u = _res_dis_params(ref)
    assert {k: tuple(v.shape) for k, v in p.items()} == dict(ores.res_discriminator_spec(RES_CFG))
    assert sorted(u) == sorted(ores.res_sn_keys(ores.res_discriminator_spec(RES_CFG)))
    labels = torch.tensor(ref["rdis_labels"].astype(np.int64))
    out = ores.res_discriminator_forward(_t64(p), _t64(u), torch.tensor(ref["rdis_in"]), labels, RES_CFG)
    assert out.shape == ref["rdis_out"].shape == (2, 1)                                      # test/test_discriminator.py:28
    assert rel_l2(out.numpy(), ref["rdis_out"]) < 1e-10


# tag -> (layer kind, kernel size, stride) of make_reference_vectors.WN_CASES
WN_LAYERS = {"conv_data": ("conv", 3, 1), "conv_norm": ("conv", 3, 1), "conv_s2_data": ("conv", 4, 2),
             "dense_data": ("dense", 0, 0), "dense_norm": ("dense", 0, 0), "deconv_norm": ("deconv", 4, 2)}


def _wn_apply(kind, stride, x, kernel, bias):
    """The wrapped Keras layer (no activation) through the oracle's layer functions, fp64."""
    x, kernel = torch.tensor(x), torch.tensor(kernel)
    b = torch.tensor(bias) if bias.size else None
    if kind == "dense":
        y = x @ kernel
        return (y + b if b is not None else y).numpy()
    if kind == "conv":
        return onets.conv2d_same(x, kernel, b, stride).numpy()
    y = onets.conv2d_transpose_same(x, kernel, stride)
    return (y + b if b is not None else y).numpy()


def test_oracle_weightnorm_matches_reference_wrapper(ref):
    """sagan/layers.py:6-211 executed by the reference around Conv2D / Dense / Conv2DTranspose: g and the bias after the
    data-dependent (:159-194) or norm (:152-157) initialisation, the kernel l2_normalize(v) * g the layer then runs with
    (:124), the output of the first and of the second call."""
    from oracle import weightnorm as own
    assert sorted(ref["wn_tags"]) == sorted(WN_LAYERS)
    for tag, (kind, k, stride) in WN_LAYERS.items():
        p = f"wn_{tag}_"
        x, v, b0 = ref[p + "x"], ref[p + "v"], ref[p + "bias0"]
        assert list(ref[p + "norm_axes"]) == list(range(v.ndim - 1)), tag          # every axis but the last (:73)
        if int(ref[p + "data_init"]):
            g, b1 = own.data_dep_init(_wn_apply(kind, stride, x, v, b0), np.ones(v.shape[-1]), b0 if b0.size else None)
        else:
            g, b1 = own.init_norm(v), (b0 if b0.size else None)
        assert rel_l2(g, ref[p + "g"]) < FP64_TOL, tag
        if b0.size:
            assert rel_l2(b1, ref[p + "bias1"]) < 1e-11, tag
        kern = own.kernel_from_vg(v, g)
        assert rel_l2(kern, ref[p + "kernel1"]) < FP64_TOL, tag
        y = _wn_apply(kind, stride, x, kern, ref[p + "bias1"])
        assert rel_l2(y, ref[p + "y1"]) < 1e-11, tag
        assert np.array_equal(ref[p + "y1"], ref[p + "y2"]), tag                   # no second initialisation


def test_oracle_record_decode_matches_reference_reader(ref):
    """get_dataset_from_tfrecord (sagan/dataset.py:12-40) run by the reference over 7 stand-in records, batch 3: the
    float32 images BIT FOR BIT, the int64 labels, and the dropped remainder."""
    from oracle import weightnorm as own
    raw, batch = ref["rec_raw"], int(ref["rec_batch"])
    keep = raw.shape[0] // batch * batch
    got = own.decode_records(raw[:keep])
    assert ref["rec_images"].dtype == np.float32 and ref["rec_images"].shape == got.shape == (keep,) + raw.shape[1:]
    assert np.array_equal(got.view(np.uint32), ref["rec_images"].view(np.uint32))
    assert ref["rec_labels"].dtype == np.int64 and np.array_equal(ref["rec_labels"], ref["rec_labels_in"][:keep])
    assert set(np.unique(raw)) == set(range(256))                                  # every byte value is covered


def test_reference_train_step_schedule_and_loss_scaling(ref):
    """Trainer.train_step / distributed_train_step (sagan/main.py:171-236) run by the reference on stub models, with
    update_ratio = 2: the schedule our trainers follow, the scalar that is differentiated, the reported losses."""
    B, ur, gbs = (int(v) for v in ref["step_cfg"])
    d_phase = ["('G', (4, 16), True, 'outside_tape')",        # :178 fakes for D are generated OUTSIDE the tape, training=True
               "('tape_open',)",
               "('D', 'real', True, 'in_tape')",              # :181
               "('D', 'fake', True, 'in_tape')",              # :182
               "('tape_close',)",
               None,                                          # :188 gradient of mean(L) / global_batch wrt D's variables
               "('apply', 'D', ('D/w0',))"]                   # :190 one Adam step per D iteration
    g_phase = ["('tape_open',)",
               "('G', (4, 16), True, 'in_tape')",             # :198
               "('D', 'fake', True, 'in_tape')",              # :199 D in training mode (its u advances), updated D weights
               "('tape_close',)",
               None,                                          # :203
               "('apply', 'G', ('G/w0', 'G/w1'))"]            # :205
    want = d_phase * ur + g_phase
    got = list(ref["step_log"])
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g.startswith("('gradient',") if w is None else g == w), (g, w)
    r, f = torch.tensor(ref["step_d_real"]), torch.tensor(ref["step_d_fake"])
    scal = [float(otrain.hinge_loss_d(r[i], f[i]).mean() / gbs) for i in range(ur)] + [float(otrain.hinge_loss_g(f[ur]).mean() / gbs)]
    assert np.allclose(ref["step_grad_scalars"], scal, rtol=1e-13, atol=0)       # :184,201
    # reported losses (:186,192,216-229): D accumulated over update_ratio and averaged, summed over the batch, divided by
    # the GLOBAL batch, then the Keras Mean over the remaining [4,4,1]
    accu = sum(otrain.hinge_loss_d(r[i], f[i]) for i in range(ur)) / ur
    assert rel_l2((accu.sum(0) / gbs).numpy(), ref["step_mean_D"]) < FP64_TOL
    assert rel_l2((otrain.hinge_loss_g(f[ur]).sum(0) / gbs).numpy(), ref["step_mean_G"]) < FP64_TOL
    assert np.allclose(ref["step_reported"], [ref["step_mean_D"].mean(), ref["step_mean_G"].mean()], rtol=1e-13)


def test_oracle_train_step_reports_what_the_reference_reports(ref):
    """oracle.train.OracleTrainer.train_step with its gradient evaluations replaced by the fixture's logits: the same
    accumulation over update_ratio and the same reported D / G losses as the reference's distributed_train_step."""
    B, ur, gbs = (int(v) for v in ref["step_cfg"])
    cfg = dict(z_dim=8, gf_dim=4, df_dim=4, img_size=64, use_attention=False, attn_dim_G=[], use_label=False, batch_size=B,
               lr_g=2e-4, lr_d=7e-4, decay_rate=0.99, update_ratio=ur)
    orc = otrain.OracleTrainer(cfg, torch.float64, global_batch_size=gbs)
    r, f = torch.tensor(ref["step_d_real"]), torch.tensor(ref["step_d_fake"])
    calls = {"d": 0, "applied": []}

    def d_grads(images, noise, labels=None, fake_labels=None):
        i = calls["d"]
        calls["d"] += 1
        return {}, otrain.hinge_loss_d(r[i], f[i])

    orc.d_grads = d_grads
    orc.g_grads = lambda noise, fake_labels=None: ({}, otrain.hinge_loss_g(f[ur]))
    orc.opt_D.apply_gradients = lambda params, grads: calls["applied"].append("D")
    orc.opt_G.apply_gradients = lambda params, grads: calls["applied"].append("G")
    rep = orc.train_step(None, [None] * ur, None)
    assert calls["applied"] == ["D"] * ur + ["G"]
    assert abs(rep["D_loss"] - ref["step_reported"][0]) < 1e-13 and abs(rep["G_loss"] - ref["step_reported"][1]) < 1e-13


# keys of the reference's config dicts that belong to its data pipeline, logging and checkpointing (SURVEY.md section 2:
# out of scope), i.e. that the layer / train-step path never reads
CONFIG_KEYS_OUTSIDE_THE_PATH = {"_description", "gpu", "dataset", "data_path", "data_size", "use_image_generator", "epoch",
                                "num_sample", "summary_step_freq", "log_dir", "ckpt_dir", "img_dir"}


def test_reference_example_configs_are_consumed_unchanged(ref):
    """example_configs/*.py as shipped: every key is either read by the product's Trainer / model builders or belongs to
    the reference's IO side; the model keys give the reference's topology (1 227 638 G / 175 438 D parameters at the
    church64 widths; discriminator.py:23 reads attn_dim_G, so attn_dim_D is accepted and ignored)."""
    import json
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "self-attention-gan_b200", "sagan_b200")
    src = "".join(open(os.path.join(pkg, f)).read() for f in ("trainer.py", "nets.py"))
    read = set(re.findall(r"""(?:cfg|config)(?:\.get\(|\[)["']([A-Za-z_]+)["']""", src))
    assert list(ref["config_names"]) == ["church64_attn", "test"]
    for name in ref["config_names"]:
        cfg = json.loads(str(ref["config_" + name]))
        unknown = set(cfg) - read - CONFIG_KEYS_OUTSIDE_THE_PATH - {"attn_dim_D"}
        assert not unknown, (name, unknown)
        assert cfg["loss"] == "hinge_loss" and cfg["model"] == "vanilla"          # the path that is built
        for k in ("z_dim", "gf_dim", "df_dim", "use_attention", "attn_dim_G", "use_label", "batch_size", "lr_g", "lr_d",
                  "decay_rate", "update_ratio", "model"):
            assert k in read, k
        cfg.setdefault("img_size", 64)                                            # sagan/main.py derives it from the dataset
        n_g = sum(int(np.prod(sh)) for _, sh in onets.generator_spec(cfg))
        n_d = sum(int(np.prod(sh)) for _, sh in onets.discriminator_spec(cfg))
        if (cfg["gf_dim"], cfg["df_dim"], cfg["z_dim"], cfg["use_attention"]) == (16, 16, 128, True):
            assert (n_g, n_d) == (1227638, 175438)


# ------------------------------------------------------------------------------------------------- CUDA == reference
def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.fixture(scope="module")
def F():
    import sagan_b200.functional as F
    from sagan_b200 import _lib
    _lib.load()
    return F


@pytest.mark.gpu
def test_cuda_power_iteration_matches_reference_update_uv(F, ref):
    """sagan_sn_plan_run on the reference's kernels (Keras layouts, raw [last dim, -1] matricisation): u, v, sigma,
    W / sigma of the first and second call within 1e-5 of what the reference's update_uv computed."""
    tags = list(ref["sn_tags"])
    by_ip = {}
    for t in tags:
        by_ip.setdefault(int(ref[f"sn_{t}_Ip"]), []).append(t)
    for Ip, group_tags in by_ip.items():
        Ws = [cu(ref[f"sn_{t}_W"]) for t in group_tags]
        us = [cu(ref[f"sn_{t}_u0"]) for t in group_tags]
        g = F.SpectralNormGroup(Ws, us, Ip, [_factor(ref, t) for t in group_tags])
        for call in (1, 2):
            g.run()
            torch.cuda.synchronize()
            for i, t in enumerate(group_tags):
                p = f"sn_{t}_"
                assert rel_l2(g.u(i).cpu().numpy(), ref[p + f"u{call}"]) < STRICT_TOL, (t, call)
                assert rel_l2(g.v(i).cpu().numpy(), ref[p + f"v{call}"]) < STRICT_TOL, (t, call)
                s = float(ref[p + f"sigma{call}"])
                assert abs(float(g.sigma(i).cpu()) - s) < STRICT_TOL * abs(s), (t, call)
                if call == 1:
                    wb = g.w_bar(i).cpu().numpy()
                    assert rel_l2(wb, ref[p + "Wbar1"]) < STRICT_TOL and wb.shape == ref[p + "W"].shape, t


@pytest.mark.gpu
def test_cuda_attention_matches_reference_call(F, ref):
    """sagan_attn_fwd (strict mode; C = 8 has no tensor-core form) against Attention_Layer.call's own output."""
    for i in range(int(ref["attn_cases"])):
        p = f"attn_{i}_"
        X, a = ref[p + "X"], _attn_args(ref, p)
        B, H, W, C = X.shape
        y = F.attention(cu(X.reshape(B, H * W, C)), cu(a["Wtheta"]), cu(a["btheta"]), cu(a["Wphi"]), cu(a["bphi"]),
                        cu(a["Wg"]), cu(a["bg"]), cu(a["Wo"]), cu(a["bo"]), cu(np.float32(a["gamma"])),
                        F.MATH_FP32_STRICT)
        torch.cuda.synchronize()
        assert rel_l2(y.cpu().numpy().reshape(X.shape), ref[p + "Y"]) < STRICT_TOL, i


@pytest.mark.gpu
def test_cuda_hinge_matches_reference(F, ref):
    real, fake = cu(ref["hinge_real"]), cu(ref["hinge_fake"])
    loss = torch.zeros(1, device="cuda")
    n = real.shape[0]
    g_real, g_fake = F.hinge_d_grads(real, fake, n, loss)
    torch.cuda.synchronize()
    # the differentiated scalar is mean(L) / global_batch (sagan/main.py:184): loss accumulates sum(L)
    assert abs(float(loss) - ref["hinge_d"].sum()) < 1e-5 * abs(ref["hinge_d"].sum())
    scale = 1.0 / (ref["hinge_d"].size * n)
    assert np.allclose(g_real.cpu().numpy(), np.where(1 - ref["hinge_real"] > 0, -scale, 0.0), rtol=1e-6, atol=0)
    assert np.allclose(g_fake.cpu().numpy(), np.where(1 + ref["hinge_fake"] > 0, scale, 0.0), rtol=1e-6, atol=0)


def _gpu_net(kind, cfg):
    from sagan_b200 import nets
    torch.manual_seed(0)
    return nets.get_generator(cfg) if kind == "G" else nets.get_discriminator(cfg)


@pytest.mark.gpu
def test_cuda_generator_matches_reference_builder(F, ref):
    """The host-side Generator (nets.py) over the CUDA kernels, loaded with the kernels the reference's get_generator
    built, reproduces the image the reference computed (strict mode, 1e-5)."""
    p, u = _gen_params(ref)
    net = _gpu_net("G", dict(BUILDER_CFG))
    z = cu(ref["gen_in"])
    with torch.no_grad():
        net([torch.zeros_like(z), None])                              # build pass (creates the parameters)
        net.load_keras_weights(p, u)
        img = net([z, None], training=True)
    torch.cuda.synchronize()
    assert rel_l2(img.cpu().numpy(), ref["gen_out"]) < STRICT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("prefix", ["dis_", "disc_"])
def test_cuda_discriminator_matches_reference_builder(F, ref, prefix):
    cfg = dict(BUILDER_CFG, use_label=True, num_classes=10) if prefix == "disc_" else dict(BUILDER_CFG)
    p, u = _dis_params(ref, prefix, cfg)
    net = _gpu_net("D", cfg)
    img = cu(ref[prefix + "in"])
    labels = torch.as_tensor(ref[prefix + "labels"].astype(np.int64)).cuda() if cfg["use_label"] else None
    with torch.no_grad():
        net([torch.zeros_like(img), labels])                          # build pass
        net.load_keras_weights(p, u)
        out = net([img, labels], training=True)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu().numpy(), ref[prefix + "out"]) < STRICT_TOL


@pytest.mark.gpu
def test_cuda_residual_nets_match_reference_builders(F, ref):
    """The residual generator / discriminator (nets.ResGenerator / ResDiscriminator over the CUDA layers, strict mode)
    loaded with the kernels the reference's legacy builders created: image and logits within 1e-5 / 2e-5."""
    from sagan_b200 import nets
    cfg = dict(RES_CFG)
    pg, ug = _res_gen_params(ref)
    pd, ud = _res_dis_params(ref)
    torch.manual_seed(0)
    G, D = nets.get_res_generator(cfg), nets.get_res_discriminator(cfg)
    z, img = cu(ref["rgen_in"]), cu(ref["rdis_in"])
    lg = torch.as_tensor(ref["rgen_labels"].astype(np.int64)).cuda()
    ld = torch.as_tensor(ref["rdis_labels"].astype(np.int64)).cuda()
    with torch.no_grad():
        G([torch.zeros_like(z), lg]), D([torch.zeros_like(img), ld])  # build pass
        G.load_keras_weights(pg, ug), D.load_keras_weights(pd, ud)
        fake, logit = G([z, lg], training=True), D([img, ld], training=True)
    torch.cuda.synchronize()
    assert rel_l2(fake.cpu().numpy(), ref["rgen_out"]) < STRICT_TOL
    assert rel_l2(logit.cpu().numpy(), ref["rdis_out"]) < 2e-5        # 19 fp32 conv layers deep, summed over 16 pixels


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(WN_LAYERS))
def test_cuda_weightnorm_matches_reference_wrapper(F, ref, tag):
    """nn.WeightNormalization (sagan_wn_fwd + the wrapped layer's kernels) given the reference's v / bias / first batch:
    the same g, bias and outputs as the reference's wrapper computed."""
    from sagan_b200 import nn as snn
    kind, k, stride = WN_LAYERS[tag]
    p = f"wn_{tag}_"
    x, v, b0 = ref[p + "x"], ref[p + "v"], ref[p + "bias0"]
    if kind == "dense":
        inner = snn.Dense(v.shape[-1])
    elif kind == "conv":
        inner = snn.Conv2D(v.shape[-1], k, stride, padding="same")
    else:
        inner = snn.Conv2DTranspose(v.shape[-2], k, stride, padding="same", use_bias=False)
    layer = snn.WeightNormalization(inner, data_init=bool(int(ref[p + "data_init"])))
    layer.build(tuple(x.shape))
    with torch.no_grad():
        layer.v.copy_(cu(v))
        if b0.size:
            inner.bias.copy_(cu(b0))
    tx = cu(x)
    y1 = layer(tx)
    y2 = layer(tx)
    torch.cuda.synchronize()
    assert rel_l2(layer.g.detach().cpu().numpy(), ref[p + "g"]) < STRICT_TOL
    if b0.size:
        assert rel_l2(inner.bias.detach().cpu().numpy(), ref[p + "bias1"]) < 2e-5
    assert rel_l2(y1.detach().cpu().numpy(), ref[p + "y1"]) < 2e-5
    assert rel_l2(y2.detach().cpu().numpy(), ref[p + "y2"]) < 2e-5


@pytest.mark.gpu
def test_cuda_record_decode_matches_reference_reader(F, ref):
    """sagan_u8_to_f32 on the raw record bytes: the reference reader's float32 images bit for bit."""
    raw, batch = ref["rec_raw"], int(ref["rec_batch"])
    keep = raw.shape[0] // batch * batch
    got = F.decode_records(torch.tensor(raw[:keep]).cuda()).cpu().numpy()
    assert got.dtype == np.float32
    assert np.array_equal(got.reshape(-1).view(np.uint32), ref["rec_images"].reshape(-1).view(np.uint32))
