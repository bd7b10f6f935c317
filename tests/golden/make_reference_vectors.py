"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON (run in the build container only).

The reference holds no golden vectors and TensorFlow is not installable here, so `make_golden.py` can only freeze the
oracle's own outputs.  This script closes the loop from the other side: it imports the UNMODIFIED reference files from
/root/reference with `tests/golden/tfshim/tensorflow` (a float64 numpy stand-in for the few `tf.*` / Keras symbols
those files use) on the import path, drives them with seeded inputs and writes what THEIR code computes to
`reference_layers.npz`.  `tests/test_reference_vectors.py` then holds the oracle (CPU) and the CUDA kernels (GPU)
to these numbers.  What is pinned is the reference's algorithm as written -- which matrix is reshaped how, what is
normalised by what, the order of the u / v updates, the channel split, where biases and the residual enter, which
layers the builders wire in which order with which kernel sizes / strides / padding / bias flags; NOT TensorFlow's
kernels (matmul, conv, softmax come from the stand-in, with their documented semantics).

    python tests/golden/make_reference_vectors.py        # needs /root/reference; writes tests/golden/reference_layers.npz

Sections of the fixture
  l2n_*    `l2normalize`                                      layers.py:4-5
  sn*_     `SpectralNormalization.build / _make_param / update_uv` on real Keras-layout kernels (Dense, Conv2D,
           Conv2DTranspose, 1x1 Conv2D), Ip in {1, 2, 3}, factor in {None, 2}; sigma and W / sigma are locals of
           `update_uv` (layers.py:62-68, never returned: SURVEY.md appendix A.1), read from its frame at return.
           Two consecutive calls, so the persistence of u is pinned too.
  attn_*   `Attention_Layer.build / call` (layers.py:71-120).  The literal call max-pools keys and values with
           MaxPool2D(2, 1) and then reshapes to the un-pooled token count (layers.py:100-101,113-114), which is
           ill-formed at every real shape; it is executed here with that one layer class
           replaced by the identity, at C = 8 where d = C // 8 = 1 makes the raw reshape of layers.py:101 a true
           transpose -- there the reference's literal dataflow IS the paper form the oracle implements.  The channel
           split (c//8, c//8, c//2, c) is read off the built layer for several C.  A second block records the literal
           (pooled, broadcasting) result at the one degenerate shape where it runs (B=4, 2x2 map), for the record.
  hinge_*  `hinge_loss_g / hinge_loss_d`, the two function definitions compiled from sagan/main.py:21-27 (the module
           itself cannot be imported: SURVEY.md appendix A.7).
  gen_* / dis_*   `get_generator / get_discriminator` of sagan/models/*.py run on concrete arrays (eager `Input`):
           every layer the builder creates, in order, with its kernel; the image / logits that come out.  The literal
           wrapper's forward uses the RAW kernel (appendix A.1), so the kernels are pre-scaled to sigma = 1 under the
           stored u, which makes the literal forward equal to the normalised forward the oracle computes.
           Attention is off in these two runs (see attn_* for why) and BatchNormalization runs on batch statistics.
  wn_*     the weight-normalisation wrapper of sagan/layers.py:6-211 around Conv2D / Dense /
           Conv2DTranspose: g and bias after the data-dependent / norm initialisation, the effective kernel, two calls.
  rec_*    get_dataset_from_tfrecord of sagan/dataset.py:12-40: the uint8 record decode in float32 (bit-exact contract),
           labels, batching with drop_remainder.
  step_*   Trainer.train_step / distributed_train_step of sagan/main.py:171-236 on stub models: the call schedule, the
           scalars that are differentiated, the reported losses (no gradients: the tape only records).
  config_* the shipped example_configs/*.py dicts as JSON.
  rgen_* / rdis_*  the legacy residual builders models/generator.py:23-43, models/discriminator.py:40-57 the same way,
           at widths where their Attention_Layer sees C = 8, so attention stays IN (identity pool).
"""
import ast
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "reference_layers.npz")


def _import_reference():
    assert os.path.isdir(REF), "the reference tree is only present in the build container"
    sys.path.insert(0, os.path.join(HERE, "tfshim"))
    sys.path.insert(0, REF)
    import tensorflow as tf            # the stand-in
    assert tf.__version__.endswith("numpy-shim")
    import layers as ref_layers        # /root/reference/layers.py, unmodified
    assert os.path.samefile(ref_layers.__file__, os.path.join(REF, "layers.py"))
    return tf, ref_layers


def _locals_at_return(fn, code_name):
    """Runs fn() and returns (result, f_locals of the frame named `code_name` at its return)."""
    box = {}

    def tracer(frame, event, arg):
        if event == "call" and frame.f_code.co_name == code_name:
            def local(frame, event, arg):
                if event == "return":
                    box.update(frame.f_locals)
                return local
            return local
        return None

    sys.settrace(tracer)
    try:
        res = fn()
    finally:
        sys.settrace(None)
    return res, box


def rng_of(seed):
    return np.random.Generator(np.random.PCG64(seed))


# ---------------------------------------------------------------------------------------------------------------------
def section_l2n(tf, ref, out):
    r = rng_of(11)
    for i, shape in enumerate([(1, 7), (1, 4096), (3, 5)]):
        x = r.standard_normal(shape)
        out["l2n_%d_x" % i] = x
        out["l2n_%d_y" % i] = ref.l2normalize(tf.Tensor(x)).numpy()
    out["l2n_zero_y"] = ref.l2normalize(tf.Tensor(np.zeros((1, 4)))).numpy()      # eps keeps 0 / 0 finite


SN_CASES = [
    # (tag, layer ctor, input shape, Ip, factor)
    ("dense", lambda L: L.Dense(512), [4, 96], 1, None),                                      # kernel [96, 512]
    ("conv_d0", lambda L: L.Conv2D(16, 4, 2, padding="same"), [4, 64, 64, 3], 1, None),        # [4,4,3,16]
    ("conv_d1", lambda L: L.Conv2D(32, 4, 2, padding="same"), [4, 32, 32, 16], 2, None),       # [4,4,16,32]
    ("deconv_g0", lambda L: L.Conv2DTranspose(32, 4, 2, padding="same", use_bias=False), [4, 4, 4, 64], 1, None),
    ("deconv_g3", lambda L: L.Conv2DTranspose(16, 4, 2, padding="same", use_bias=False), [4, 32, 32, 32], 3, None),
    ("conv1x1_q", lambda L: L.Conv2D(4, 1, 1), [4, 8, 8, 32], 1, None),                        # [1,1,32,4]
    ("conv1x1_o", lambda L: L.Conv2D(32, 1, 1), [4, 8, 8, 16], 1, 2.0),                        # factor
    ("conv3x3", lambda L: L.Conv2D(24, 3, 1, padding="same"), [2, 8, 8, 10], 2, 0.5),
]


def section_sn(tf, ref, out):
    L = tf.keras.layers
    tags = []
    for ci, (tag, ctor, in_shape, Ip, factor) in enumerate(SN_CASES):
        tf.random.set_seed(100 + ci)
        L.seed_initializers(200 + ci)
        module = ctor(L)
        sn = ref.SpectralNormalization(module, Ip=Ip, factor=factor)
        sn.build(tf.TensorShape(in_shape))                       # layers.py:40-43: builds the module, makes u / v
        W = module.weights[0]
        # a kernel with a non-trivial spectrum (glorot init would do; this has larger dynamic range)
        W.assign(rng_of(300 + ci).standard_normal(W.shape) * 0.05)
        p = "sn_%s_" % tag
        out[p + "W"], out[p + "u0"], out[p + "v0"] = W.numpy().copy(), sn.u.numpy().copy(), sn.v.numpy().copy()
        out[p + "Ip"], out[p + "factor"] = np.int64(Ip), np.float64(factor if factor else 0.0)
        for call in (1, 2):
            _, loc = _locals_at_return(sn.update_uv, "update_uv")
            out[p + "u%d" % call], out[p + "v%d" % call] = sn.u.numpy().copy(), sn.v.numpy().copy()
            out[p + "sigma%d" % call] = np.float64(loc["sigma"].numpy())
            if call == 1:
                out[p + "Wbar1"] = loc["W"].numpy().copy()
            out[p + "Wmat_shape"] = np.asarray(loc["W_mat"].shape, dtype=np.int64)
        tags.append(tag)
    out["sn_tags"] = np.asarray(tags)
    # the constructor's argument check (layers.py:17-18)
    try:
        ref.SpectralNormalization(L.Dense(4), Ip=0)
        raised = ""
    except ValueError as e:
        raised = str(e)
    out["sn_ip0_error"] = np.asarray(raised)


class _IdentityPool:
    """Stands in for keras.layers.MaxPool2D inside Attention_Layer.call for the well-formed run (see module docstring)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, x):
        return x


def _set_attention_weights(att, seed, gamma):
    """Seeded kernels / biases for the four 1x1 convs.  layers.py:87-90 builds each wrapped conv once by hand and
    Keras builds it again on first call, so a conv owns two kernels (weights[0] and .kernel): both get the same values."""
    r = rng_of(seed)
    got = {}
    for name, wrap in zip(("phi", "theta", "g", "o"), att.SN_conv):
        m = wrap.module
        shape = m.kernel.shape
        k = r.standard_normal(shape) * (0.6 if name in ("phi", "theta") else 0.3)
        b = r.standard_normal(shape[-1:]) * 0.1
        for w in m.weights:
            w.assign(k if len(w.shape) == 4 else b)
        got["W" + name], got["b" + name] = k.reshape(shape[2], shape[3]).copy(), b.copy()
    att.sigma.assign(np.float64(gamma))
    got["gamma"] = np.float64(gamma)
    return got


def section_attention(tf, ref, out):
    L = tf.keras.layers
    # channel split of build() for several C
    splits = []
    for C in (8, 16, 32, 64, 128):
        att = ref.Attention_Layer()
        att.build(tf.TensorShape([2, 4, 4, C]))
        splits.append([C] + [int(w.module.filters) for w in att.SN_conv])
        assert att.sigma.shape == () and float(att.sigma.numpy()) == 0.0          # zero-initialised scalar
    out["attn_split"] = np.asarray(splits, dtype=np.int64)

    # well-formed run: identity instead of MaxPool2D(2, 1), C = 8 (d = 1)
    real_pool = L.MaxPool2D
    cases = [(2, 4, 4, 8, 0.37), (3, 8, 8, 8, -1.25), (1, 16, 16, 8, 0.0)]
    for i, (B, H, Wd, C, gamma) in enumerate(cases):
        tf.random.set_seed(400 + i)
        L.seed_initializers(500 + i)
        X = rng_of(600 + i).standard_normal((B, H, Wd, C))
        att = ref.Attention_Layer()
        L.MaxPool2D = _IdentityPool
        try:
            att(tf.Tensor(X))                                     # builds (twice, as Keras would)
            w = _set_attention_weights(att, 700 + i, gamma)
            Y = att(tf.Tensor(X)).numpy()
        finally:
            L.MaxPool2D = real_pool
        p = "attn_%d_" % i
        out[p + "X"], out[p + "Y"] = X, Y
        for k, v in w.items():
            out[p + k] = v
    out["attn_cases"] = np.int64(len(cases))

    # literal run (real MaxPool2D(2, 1)) at the one shape family where the reshapes go through: a 2x2 map, B = 4
    tf.random.set_seed(450)
    L.seed_initializers(550)
    X = rng_of(650).standard_normal((4, 2, 2, 16))
    att = ref.Attention_Layer()
    att(tf.Tensor(X))
    w = _set_attention_weights(att, 750, 0.5)
    out["attnlit_X"], out["attnlit_Y"] = X, att(tf.Tensor(X)).numpy()
    for k, v in w.items():
        out["attnlit_" + k] = v


def section_hinge(tf, out):
    src = open(os.path.join(REF, "sagan", "main.py")).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("hinge_loss_g", "hinge_loss_d")]
    assert len(fns) == 2
    ns = {"tf": tf}
    exec(compile(ast.Module(body=fns, type_ignores=[]), os.path.join(REF, "sagan", "main.py"), "exec"), ns)
    r = rng_of(21)
    real, fake = r.standard_normal((6, 4, 4, 1)) * 2, r.standard_normal((6, 4, 4, 1)) * 2
    out["hinge_real"], out["hinge_fake"] = real, fake
    out["hinge_d"] = ns["hinge_loss_d"](tf.Tensor(real), tf.Tensor(fake)).numpy()
    out["hinge_g"] = ns["hinge_loss_g"](tf.Tensor(fake)).numpy()


def _import_builder(tf, ref_layers, which, tree="sagan/models"):
    """sagan/models/<which>.py does `from layers import SpectralNormalization, AttentionLayer[, SNConv2D, SNDense]`
    (names the top-level layers.py spells differently or lacks: SURVEY.md appendix A.7).  The module object named
    `layers` it sees is the reference's layers.py plus those names; the builder file itself is read unmodified."""
    import importlib.util
    import types
    facade = types.ModuleType("layers")
    facade.__dict__.update(ref_layers.__dict__)
    facade.AttentionLayer = ref_layers.Attention_Layer
    facade.SNConv2D = facade.SNDense = None                      # imported by discriminator.py, unused by the vanilla builder
    saved = sys.modules.get("layers")
    sys.modules["layers"] = facade
    try:
        path = os.path.join(REF, *tree.split("/"), which + ".py")
        spec = importlib.util.spec_from_file_location("ref_" + tree.replace("/", "_") + "_" + which, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["layers"] = saved
    return mod


def _prescale_to_unit_sigma(ref_layers, model_layers, seed, attn_gamma=0.37):
    """Seeded kernels for every layer; a wrapped module's kernel is divided by the sigma the reference's own update_uv
    computes for it from the wrapper's current u, so that sigma(kernel; u) = 1 and W / sigma = W.  A conv built twice
    (Attention_Layer's, layers.py:87-90 + Keras' own build) owns two kernels -- weights[0], which update_uv reads, and
    .kernel, which call uses: both get the same values."""
    r = rng_of(seed)
    wrapped = {id(l.module): l for l in model_layers if isinstance(l, ref_layers.SpectralNormalization)}
    for l in model_layers:
        if isinstance(l, ref_layers.SpectralNormalization) or not l.weights:
            continue
        vals = {}
        for w in l.weights:
            if w.name in ("kernel", "embeddings", "bias"):
                if w.name not in vals:
                    vals[w.name] = r.standard_normal(w.shape) * (0.05 if w.name == "bias" else 0.08)
                w.assign(vals[w.name])
            elif w.name == "sigma":
                w.assign(np.float64(attn_gamma))
        if id(l) in wrapped:
            sn = wrapped[id(l)]
            u_keep, v_keep = sn.u, sn.v
            _, loc = _locals_at_return(sn.update_uv, "update_uv")
            sig = float(loc["sigma"].numpy())
            for w in l.weights:
                if w.name in ("kernel", "embeddings"):
                    w.assign(w.numpy() / sig)
            sn.u, sn.v = u_keep, v_keep                          # the probe must not advance the stored u


def _dump_layers(ref_layers, model_layers, prefix, out):
    """Creation-ordered description of what the builder wired: kind, hyper-parameters, kernels, and for wrapped
    modules the u the wrapper holds."""
    kinds = []
    wrapped = {id(l.module): l for l in model_layers if isinstance(l, ref_layers.SpectralNormalization)}
    n = 0
    for l in model_layers:
        if isinstance(l, ref_layers.SpectralNormalization):
            continue
        kind = type(l).__name__
        desc = [kind, "sn" if id(l) in wrapped else "plain"] + (["inner"] if l._inner else [])
        for attr in ("filters", "units", "kernel_size", "strides", "padding", "use_bias", "activation", "alpha",
                     "epsilon", "momentum"):
            if hasattr(l, attr):
                desc.append("%s=%s" % (attr, getattr(l, attr)))
        kinds.append(" ".join(desc))
        for w in l.weights:
            out["%sL%d_%s" % (prefix, n, w.name)] = w.numpy().copy()
        if id(l) in wrapped:
            out["%sL%d_u" % (prefix, n)] = wrapped[id(l)].u.numpy().copy()
        n += 1
    out[prefix + "layers"] = np.asarray(kinds)


def section_builders(tf, ref_layers, out):
    K, L = tf.keras, tf.keras.layers
    cfg = dict(z_dim=32, gf_dim=4, df_dim=4, img_size=64, use_attention=False, attn_dim_G=[32, 64],
               attn_dim_D=[8, 4], use_label=False, batch_size=2, num_classes=1)
    B = cfg["batch_size"]
    # generator: the first pass builds the layers, then kernels are seeded and the builder's own layers re-applied
    gen = _import_builder(tf, ref_layers, "generator")
    dis = _import_builder(tf, ref_layers, "discriminator")
    cfg_label = dict(cfg, use_label=True, num_classes=10)        # projection head (discriminator.py:26-33)
    for which, mod, fn, data, cfg in (
            ("gen_", gen, "get_generator", rng_of(31).standard_normal((B, cfg["z_dim"])), cfg),
            ("dis_", dis, "get_discriminator", rng_of(32).uniform(-1, 1, (B, 64, 64, 3)), cfg),
            ("disc_", dis, "get_discriminator", rng_of(33).uniform(-1, 1, (B, 64, 64, 3)), cfg_label)):
        tf.random.set_seed(40)
        L.seed_initializers(41)
        labels = np.asarray([7, 2], dtype=np.int32) if cfg["use_label"] else np.zeros((B,), dtype=np.int32)
        out[which + "labels"] = labels
        K.feed(data, labels)
        model = getattr(mod, fn)(cfg)                            # pass 1: creates + builds every layer
        layers1 = list(model.layers)
        _prescale_to_unit_sigma(ref_layers, layers1, 50)
        _dump_layers(ref_layers, layers1, which, out)            # kernels and the u each wrapper holds BEFORE pass 2
        # pass 2: same layer OBJECTS, replayed in creation order by handing them back to the builder's constructors
        replay = L.replay(layers1)
        with replay:
            K.feed(data, labels)
            model2 = getattr(mod, fn)(cfg)
        assert replay.done(), "builder created a different layer sequence on the second pass"
        out[which + "in"] = data
        out[which + "out"] = model2.outputs.numpy()


def section_res_builders(tf, ref_layers, out):
    """The legacy residual builders, models/generator.py:23-43 and models/discriminator.py:40-57, at
    widths where their Attention_Layer sees C = 8 (gf_dim 2, df_dim 4), with the identity pool (see attn_*), called with
    training=True (batch-statistics BatchNormalization; the wrapper then skips update_uv, layers.py:46)."""
    K, L = tf.keras, tf.keras.layers
    B, ncls = 2, 5
    gen = _import_builder(tf, ref_layers, "generator", "models")
    dis = _import_builder(tf, ref_layers, "discriminator", "models")
    labels = np.asarray([3, 1], dtype=np.int32)
    real_pool = L.MaxPool2D
    L.MaxPool2D = _IdentityPool
    try:
        for which, build, data in (
                ("rgen_", lambda: gen.get_generator(ncls, gf_dim=2, training=True), rng_of(61).standard_normal((B, 128))),
                ("rdis_", lambda: dis.get_discriminator(ncls, df_dim=4, training=True),
                 rng_of(62).uniform(-1, 1, (B, 128, 128, 3)))):
            tf.random.set_seed(70)
            L.seed_initializers(71)
            K.feed(data, labels)
            model = build()
            layers1 = list(model.layers)
            _prescale_to_unit_sigma(ref_layers, layers1, 80)
            _dump_layers(ref_layers, layers1, which, out)
            replay = L.replay(layers1)
            with replay:
                K.feed(data, labels)
                model2 = build()
            assert replay.done()
            out[which + "in"], out[which + "labels"], out[which + "out"] = data, labels, model2.outputs.numpy()
    finally:
        L.MaxPool2D = real_pool


WN_CASES = [
    # (tag, layer ctor, input shape, data_init)
    ("conv_data", lambda L: L.Conv2D(24, 3, 1, padding="same"), (4, 16, 16, 8), True),
    ("conv_norm", lambda L: L.Conv2D(24, 3, 1, padding="same"), (4, 16, 16, 8), False),
    ("conv_s2_data", lambda L: L.Conv2D(16, 4, 2, padding="same"), (3, 12, 12, 5), True),
    ("dense_data", lambda L: L.Dense(40), (16, 24), True),
    ("dense_norm", lambda L: L.Dense(40), (16, 24), False),
    # Conv2DTranspose: the kernel's last axis is the INPUT channel count, so g has cin entries; with data_init=True the
    # reference multiplies that g by a cout-sized scale (sagan/layers.py:189) and fails unless cin == cout -- norm init only
    ("deconv_norm", lambda L: L.Conv2DTranspose(8, 4, 2, padding="same", use_bias=False), (2, 6, 6, 12), False),
]


def section_weightnorm(tf, out):
    """sagan/layers.py (the weight-normalisation wrapper the `sagan/` tree ships under the name SpectralNormalization,
    SURVEY.md 8f-4), imported unmodified: build, the first call with its data-dependent (sagan/layers.py:159-194) or
    norm (:152-157) initialisation of g and of the wrapped layer's bias, and a second call."""
    import importlib.util
    L = tf.keras.layers
    spec = importlib.util.spec_from_file_location("ref_sagan_layers", os.path.join(REF, "sagan", "layers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tags = []
    for ci, (tag, ctor, in_shape, data_init) in enumerate(WN_CASES):
        L.seed_initializers(900 + ci)
        r = rng_of(910 + ci)
        x = r.standard_normal(in_shape) * 1.3 + 0.2
        layer = ctor(L)
        wrap = mod.SpectralNormalization(layer, data_init=data_init)
        wrap.build(tf.TensorShape(in_shape))
        wrap.built = True
        v0 = r.standard_normal(layer.kernel.shape) * 0.2
        layer.kernel.assign(v0)                                   # == wrap.v (sagan/layers.py:85)
        if layer.bias is not None:
            layer.bias.assign(r.standard_normal(layer.bias.shape) * 0.1)
        if data_init:                                             # the clone made in build() copies the weights (:101)
            wrap._naked_clone_layer.set_weights(layer.get_weights())
        p = "wn_%s_" % tag
        out[p + "x"], out[p + "v"] = x, v0.copy()
        out[p + "bias0"] = layer.bias.numpy().copy() if layer.bias is not None else np.zeros(0)
        out[p + "data_init"] = np.int64(data_init)
        y1 = wrap(tf.Tensor(x)).numpy()
        out[p + "g"] = wrap.g.numpy().copy()
        out[p + "bias1"] = layer.bias.numpy().copy() if layer.bias is not None else np.zeros(0)
        out[p + "kernel1"] = tf._core.raw(layer.kernel).copy()    # l2_normalize(v) * g, what the layer ran with (:124-130)
        out[p + "y1"] = y1
        assert bool(wrap._initialized.numpy())
        out[p + "y2"] = wrap(tf.Tensor(x)).numpy()                # second call: no re-initialisation (:113-114)
        assert np.array_equal(out[p + "g"], wrap.g.numpy())
        out[p + "norm_axes"] = np.asarray(wrap.kernel_norm_axes, dtype=np.int64)
        tags.append(tag)
    out["wn_tags"] = np.asarray(tags)


def section_records(tf, out):
    """get_dataset_from_tfrecord of sagan/dataset.py:12-40, imported unmodified and run over stand-in records (a dict per
    record instead of a serialized tf.train.Example, see tfshim/tensorflow/io.py): which features are read, the uint8
    HWC decode, `cast(float32) * (2. / 255) - 1.` in float32, int64 labels, batches of global_batch_size with the
    remainder dropped.  Includes every byte value 0..255."""
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location("ref_sagan_dataset", os.path.join(REF, "sagan", "dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    size, n, batch = 8, 7, 3
    r = rng_of(95)
    raw = r.integers(0, 256, (n, size, size, 3), dtype=np.uint8)
    raw.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)                      # every byte value once
    labels = r.integers(0, 1000, n).astype(np.int64)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "part-0.tfrecords")
        open(path, "wb").close()                                                # the reference globs for *.tfrecords
        tf.data.records[path] = [{"label": int(labels[i]), "image_raw": raw[i].tobytes(),
                                  "height": size, "width": size, "depth": 3} for i in range(n)]
        ds = mod.get_dataset_from_tfrecord({"data_path": d, "img_size": size, "data_size": -1, "global_batch_size": batch})
        batches = list(ds)
    assert len(batches) == n // batch                                           # drop_remainder
    imgs = np.concatenate([b[0].numpy() for b in batches])
    labs = np.concatenate([b[1].numpy() for b in batches])
    assert imgs.dtype == np.float32 and labs.dtype == np.int64
    out["rec_raw"], out["rec_labels_in"] = raw, labels
    out["rec_images"], out["rec_labels"] = imgs, labs
    out["rec_batch"] = np.int64(batch)


def section_train_step(tf, out):
    """Trainer.train_step and Trainer.distributed_train_step of sagan/main.py:171-236, the two method definitions
    compiled from the file (the module does not import, appendix A.7) and run on a stand-in `self`: stub generator /
    discriminator that log their calls and return fixed logits, a GradientTape that records what is differentiated, stub
    optimisers, a one-replica strategy.  Pins (a) the schedule -- which model is called on what, in which order, with
    which `training` flag, inside or outside the tape, when each optimiser is applied; (b) the scalar handed to
    tape.gradient, mean(L) / global_batch_size; (c) the returned per-example losses and the reported means of
    :216-229.  update_ratio = 2 so that the accumulation of :186,192 shows."""
    import types
    src = open(os.path.join(REF, "sagan", "main.py")).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Trainer"][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("train_step", "distributed_train_step")]
    assert len(fns) == 2
    hinge = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("hinge_loss_g", "hinge_loss_d")]
    ns = {"tf": tf}
    exec(compile(ast.Module(body=hinge + fns, type_ignores=[]), os.path.join(REF, "sagan", "main.py"), "exec"), ns)

    B, ur, gbs = 4, 2, 8                     # per-replica batch 4, global batch 8 (two replicas' worth), update_ratio 2
    r = rng_of(77)
    d_real = [r.standard_normal((B, 4, 4, 1)) * 2 for _ in range(ur)]
    d_fake = [r.standard_normal((B, 4, 4, 1)) * 2 for _ in range(ur + 1)]       # ur D-phase fakes + the G-phase one
    log = tf.GradientTape.log
    log.clear()
    counters = {"G": 0, "Dreal": 0, "Dfake": 0}

    class Var:
        def __init__(self, name):
            self.name = name

    class Gen:
        trainable_variables = [Var("G/w0"), Var("G/w1")]
        variables = []

        def __call__(self, inputs, training=None):
            noise, labels = inputs
            counters["G"] += 1
            log.append(("G", tuple(noise.shape), training, "in_tape" if tf.GradientTape.depth else "outside_tape"))
            t = tf.Tensor(np.full((noise.shape[0], 2, 2, 3), float(counters["G"])))
            t.is_fake = True
            return t

    class Dis:
        trainable_variables = [Var("D/w0")]

        def __call__(self, inputs, training=None):
            images, labels = inputs
            fake = getattr(images, "is_fake", False)
            k = "Dfake" if fake else "Dreal"
            log.append(("D", "fake" if fake else "real", training, "in_tape" if tf.GradientTape.depth else "outside_tape"))
            val = (d_fake if fake else d_real)[counters[k]]
            counters[k] += 1
            return tf.Tensor(val)

    class Opt:
        def __init__(self, name):
            self.name = name

        def apply_gradients(self, grads_and_vars):
            log.append(("apply", self.name, tuple(v.name for _, v in grads_and_vars)))

    class Strategy:
        def experimental_run_v2(self, fn, args=()):
            return fn(*args)

        def reduce(self, op, value, axis=None):
            assert op == tf.distribute.ReduceOp.SUM
            if isinstance(value, tuple):                     # a gradient token (main.py:222-226 reduces every G gradient)
                return value
            return tf.reduce_sum(value, axis=axis) if axis is not None else value

    reported = {}
    me = types.SimpleNamespace(
        config={"update_ratio": ur, "z_dim": 16, "num_classes": 1, "global_batch_size": gbs},
        generator=Gen(), discriminator=Dis(), optimizer_D=Opt("D"), optimizer_G=Opt("G"),
        dloss_fn=ns["hinge_loss_d"], gloss_fn=ns["hinge_loss_g"], strategy=Strategy(),
        metrics={"G_loss": lambda v: reported.__setitem__("G_loss", float(np.mean(v.numpy()))),
                 "D_loss": lambda v: reported.__setitem__("D_loss", float(np.mean(v.numpy())))})
    me.train_step = lambda inputs: ns["train_step"](me, inputs)
    tf.random.set_seed(5)
    images = tf.Tensor(r.uniform(-1, 1, (B, 8, 8, 3)))
    labels = tf.Tensor(np.zeros((B,), dtype=np.int32))
    mean_loss = ns["distributed_train_step"](me, (images, labels))
    out["step_d_real"], out["step_d_fake"] = np.stack(d_real), np.stack(d_fake)
    out["step_cfg"] = np.asarray([B, ur, gbs], dtype=np.int64)
    out["step_log"] = np.asarray([repr(e) for e in log])
    out["step_grad_scalars"] = np.asarray([e[1] for e in log if e[0] == "gradient"])
    out["step_mean_D"], out["step_mean_G"] = mean_loss["D_loss"].numpy(), mean_loss["G_loss"].numpy()
    out["step_reported"] = np.asarray([reported["D_loss"], reported["G_loss"]])


def section_configs(out):
    """The reference's shipped configuration dicts (example_configs/*.py, what sagan/main.py:352-355 loads), evaluated and
    stored as JSON: the key set and the values our Trainer has to accept unchanged."""
    import json
    import runpy
    names = sorted(f for f in os.listdir(os.path.join(REF, "example_configs")) if f.endswith(".py"))
    for f in names:
        ns = runpy.run_path(os.path.join(REF, "example_configs", f))
        out["config_" + f[:-3]] = np.asarray(json.dumps(ns["config"], sort_keys=True))
    out["config_names"] = np.asarray([f[:-3] for f in names])


def main():
    tf, ref = _import_reference()
    out = {}
    section_l2n(tf, ref, out)
    section_sn(tf, ref, out)
    section_attention(tf, ref, out)
    section_hinge(tf, out)
    section_builders(tf, ref, out)
    section_res_builders(tf, ref, out)
    section_weightnorm(tf, out)
    section_records(tf, out)
    section_train_step(tf, out)
    section_configs(out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "(%d arrays, %.0f kB)" % (len(out), os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
