"""GPU parity tests at network / training-step level: sagan_b200.nets + sagan_b200.trainer against the
oracle (oracle.nets / oracle.train) and the committed golden fixtures (tests/golden/nets.npz,
trajectory.npz), on identical weights, spectral-norm `u` vectors and injected noise.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402

from oracle import nets as onets  # noqa: E402
from oracle import train as otrain  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


MODES = ["fp32_strict", "bf16_tc"]


def _math(mode_name):
    from sagan_b200 import MATH_BF16_TC, MATH_FP32_STRICT
    return {"fp32_strict": MATH_FP32_STRICT, "bf16_tc": MATH_BF16_TC}[mode_name]


def make_trainer(cfg, mode_name="fp32_strict", **kw):
    """GPU trainer whose layers run in the given math mode (bf16_tc is the mode bench.py times)."""
    from sagan_b200 import nn as snn
    from sagan_b200.trainer import Trainer
    snn.set_default_math_mode(_math(mode_name))
    try:
        return Trainer(cfg, **kw)
    finally:
        snn.set_default_math_mode(_math("fp32_strict"))


def make_pair(cfg, attn_sigma, bias_scale, dtype=torch.float64, steps_per_epoch=1000, mode_name="fp32_strict"):
    """Oracle trainer and GPU trainer holding identical weights and spectral-norm state."""
    orc = otrain.OracleTrainer(cfg, dtype, seed=0, attn_sigma=attn_sigma, bias_scale=bias_scale,
                               global_batch_size=cfg["batch_size"], steps_per_epoch=steps_per_epoch)
    tr = make_trainer(cfg, mode_name, global_batch_size=cfg["batch_size"], steps_per_epoch=steps_per_epoch)
    tr.G.load_keras_weights({k: v.numpy() for k, v in orc.G.items()},
                            {k: v.numpy() for k, v in orc.G_sn.items() if k.endswith(".u")})
    tr.D.load_keras_weights({k: v.numpy() for k, v in orc.D.items()},
                            {k: v.numpy() for k, v in orc.D_sn.items() if k.endswith(".u")})
    return orc, tr


def test_param_inventory_matches_oracle_spec():
    """Same parameter names / shapes / counts as the reference topology (1 227 638 G, 175 438 D at church64)."""
    from sagan_b200.trainer import Trainer
    cfg = dict(mg.TEST_CFG)
    tr = Trainer(cfg)
    g = dict(tr.G.named_parameters_by_oracle_name())
    d = dict(tr.D.named_parameters_by_oracle_name())
    assert {k: tuple(v.shape) for k, v in g.items()} == dict(onets.generator_spec(cfg))
    assert {k: tuple(v.shape) for k, v in d.items()} == dict(onets.discriminator_spec(cfg))
    assert sum(v.numel() for v in g.values()) == 1227638 and sum(v.numel() for v in d.values()) == 175438
    # same set of spectrally-normalised matrices (the group orders them by module registration)
    assert sorted(tuple(s) for s in tr.G.sn_group.shapes) == sorted(onets.sn_keys(onets.generator_spec(cfg)).values())
    assert sorted(tuple(s) for s in tr.D.sn_group.shapes) == sorted(onets.sn_keys(onets.discriminator_spec(cfg)).values())
    assert [k for k, _ in tr.G.sn_by_oracle_name()] == list(onets.sn_keys(onets.generator_spec(cfg)))
    assert [k for k, _ in tr.D.sn_by_oracle_name()] == list(onets.sn_keys(onets.discriminator_spec(cfg)))


def test_forward_and_gradients_vs_oracle_and_golden():
    """One D-phase and one G-phase gradient evaluation at the example_configs/test.py model (B = 4)."""
    import sagan_b200.functional as F
    cfg = dict(mg.TEST_CFG)
    gold = np.load(os.path.join(GOLD, "nets.npz"))
    orc, tr = make_pair(cfg, attn_sigma=0.37, bias_scale=0.05)
    img, nd, ng = mg.step_inputs(cfg, 0)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)

    # ---- D phase (sagan/main.py:176-189)
    dgr, dl = orc.d_grads(t64(img), t64(nd))
    assert rel_l2(dl.numpy(), gold["D_loss_elems"]) < 1e-12           # oracle == golden
    with torch.no_grad():
        fake = tr.G([cu(nd), None], training=True)
    tr.D.zero_grad_flat()
    d_real = tr.D([cu(img), None], training=True)
    d_fake = tr.D([fake, None], training=True)
    loss = torch.zeros(1, device="cuda")
    g_real, g_fake = F.hinge_d_grads(d_real, d_fake, cfg["batch_size"], loss)
    torch.autograd.backward([d_real, d_fake], [g_real, g_fake])
    torch.cuda.synchronize()
    le = (torch.relu(1 - d_real) + torch.relu(1 + d_fake)).detach().cpu().numpy()
    assert rel_l2(le, gold["D_loss_elems"]) < 1e-5
    worst = 0.0
    for k, p in tr.D.named_parameters_by_oracle_name():
        if k.endswith("phi.bias"):   # mathematically zero gradient (softmax shift invariance)
            continue
        e = rel_l2(p.grad.cpu().numpy(), dgr[k].numpy())
        worst = max(worst, e)
        assert e < 1e-4, (k, e)
        assert rel_l2(mg.summarize(p.grad.cpu().numpy()), gold["Dgrad." + k]) < 1e-4, k
    print("D-phase worst per-parameter gradient rel-L2:", worst)

    # ---- G phase (sagan/main.py:194-204)
    ggr, gl = orc.g_grads(t64(ng))
    tr.G.zero_grad_flat()
    for p in tr.D.parameters():
        p.requires_grad_(False)
    fake = tr.G([cu(ng), None], training=True)
    d_fake = tr.D([fake, None], training=True)
    g = F.hinge_g_grads(d_fake, cfg["batch_size"], loss)
    d_fake.backward(g)
    for p in tr.D.parameters():
        p.requires_grad_(True)
    torch.cuda.synchronize()
    assert rel_l2((-d_fake).detach().cpu().numpy(), gold["G_loss_elems"]) < 1e-5
    worst = 0.0
    for k, p in tr.G.named_parameters_by_oracle_name():
        if k.endswith("phi.bias"):
            continue
        e = rel_l2(p.grad.cpu().numpy(), ggr[k].numpy())
        worst = max(worst, e)
        assert e < 1e-4, (k, e)
        assert rel_l2(mg.summarize(p.grad.cpu().numpy()), gold["Ggrad." + k]) < 1e-4, k
    print("G-phase worst per-parameter gradient rel-L2:", worst)
    # spectral-norm state advanced identically (G: 2 forwards, D: 3 forwards)
    for k, m in tr.G.sn_by_oracle_name() + tr.D.sn_by_oracle_name():
        ref = (orc.G_sn if k in orc.G_sn else orc.D_sn)[k]
        assert rel_l2(m.u.cpu().numpy(), ref.numpy()) < 1e-5, k


def _phase_grads(tr, cfg, img, nd, ng, labels=None, fake_labels=None):
    """D-phase and G-phase gradients of the GPU trainer (sagan/main.py:176-204) -> (D loss elems, G loss elems)."""
    import sagan_b200.functional as F
    lab = None if labels is None else torch.as_tensor(labels).cuda()
    flab = None if fake_labels is None else torch.as_tensor(fake_labels).cuda()
    with torch.no_grad():
        fake = tr.G([cu(nd), flab], training=True)
    tr.D.zero_grad_flat()
    d_real = tr.D([cu(img), lab], training=True)
    d_fake = tr.D([fake, flab], training=True)
    loss = torch.zeros(1, device="cuda")
    g_real, g_fake = F.hinge_d_grads(d_real, d_fake, cfg["batch_size"], loss)
    torch.autograd.backward([d_real, d_fake], [g_real, g_fake])
    le_d = (torch.relu(1 - d_real) + torch.relu(1 + d_fake)).detach().cpu().numpy()
    tr.G.zero_grad_flat()
    for p in tr.D.parameters():
        p.requires_grad_(False)
    fake = tr.G([cu(ng), flab], training=True)
    d_fake = tr.D([fake, flab], training=True)
    g = F.hinge_g_grads(d_fake, cfg["batch_size"], loss)
    d_fake.backward(g)
    for p in tr.D.parameters():
        p.requires_grad_(True)
    torch.cuda.synchronize()
    return le_d, (-d_fake).detach().cpu().numpy()


def test_forward_and_gradients_bf16_tc_mode():
    """The same D-phase / G-phase evaluation with every conv / deconv / dense / attention layer on the tcgen05 path
    (BF16_TC mode: tf32 conv forward / backward-data, bf16 backward-filter, bf16 attention with consistent rounding).
    Forward outputs of both networks and both loss tensors against the fp64 oracle: BASELINE.json tolerance 2e-3.
    Parameter gradients are reported and bounded, not held to 2e-3: a forward error of 1e-3 moves ~0.1 % of the
    LeakyReLU(0.1) inputs across zero, each flip changes dz by 0.9 dy, i.e. ~3 % rel-L2 per layer on the gradient --
    a property of ANY reduced-precision forward through this topology (the FP32_STRICT test above holds 1e-4)."""
    cfg = dict(mg.TEST_CFG)
    orc, tr = make_pair(cfg, attn_sigma=0.37, bias_scale=0.05, mode_name="bf16_tc")
    img, nd, ng = mg.step_inputs(cfg, 0)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    # forward outputs of each network on its own (generated images, logits on real images), from the INITIAL
    # spectral-norm state on both sides
    from oracle import nets as onets
    with torch.no_grad():
        ref_img = onets.generator_forward(orc.G, dict(orc.G_sn), t64(nd), cfg, None, True, None)
        ref_logit = onets.discriminator_forward(orc.D, dict(orc.D_sn), t64(img), cfg, None, True)
    dgr, dl = orc.d_grads(t64(img), t64(nd))
    # the oracle's G phase must see the same D state as ours: evaluate it before any update (no optimiser step here)
    ggr, gl = orc.g_grads(t64(ng))
    snap = (tr.G.sn_group.out.clone(), tr.D.sn_group.out.clone())
    with torch.no_grad():
        got_img = tr.G([cu(nd), None], training=True)
        got_logit = tr.D([cu(img), None], training=True)
    tr.G.sn_group.out.copy_(snap[0]); tr.D.sn_group.out.copy_(snap[1])      # undo the power-iteration advance
    e_img, e_logit = rel_l2(got_img.cpu().numpy(), ref_img.numpy()), rel_l2(got_logit.cpu().numpy(), ref_logit.numpy())
    le_d, le_g = _phase_grads(tr, cfg, img, nd, ng)
    e_d, e_g = rel_l2(le_d, dl.numpy()), rel_l2(le_g, gl.numpy())
    print("BF16_TC forward: G(z) %.2e  D(x) %.2e | loss elems: D %.2e G %.2e" % (e_img, e_logit, e_d, e_g))
    assert e_img < 2e-4 and e_logit < 2e-4 and e_d < 2e-4
    assert e_g < 2e-4          # -D(G(z)): both networks composed
    errs = {}
    for net, ref in ((tr.D, dgr), (tr.G, ggr)):
        for k, p in net.named_parameters_by_oracle_name():
            if k.endswith("phi.bias"):
                continue
            errs[("D." if net is tr.D else "G.") + k] = rel_l2(p.grad.cpu().numpy(), ref[k].numpy())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print("BF16_TC worst per-parameter gradient rel-L2:", [(k, "%.2e" % v) for k, v in worst])
    print("BF16_TC median per-parameter gradient rel-L2: %.2e" % float(np.median(list(errs.values()))))
    # round 1 (tf32 conv forward): worst 1.5e-1, median 5e-2.  With split-bf16 convs the forward is fp32-grade (G(z) 2e-5,
    # D(x) 8e-6), a handful of LeakyReLU flips remain (each flip among the n pre-activations of a layer costs
    # 0.9 / sqrt(n) on everything upstream): measured worst 1.0e-2, median 2.9e-3
    assert max(errs.values()) < 3e-2 and float(np.median(list(errs.values()))) < 8e-3


@pytest.mark.parametrize("mode_name", MODES)
def test_downsampled_attention_model_forward_and_gradients(mode_name):
    """SURVEY.md §8f row 2 at model level: church64 G / D with `attn_downsample` (keys / values max-pooled 2x2 / stride 2
    in all three attention layers) against the fp64 oracle: loss elements and every parameter gradient."""
    cfg = dict(mg.TEST_CFG, attn_downsample=True)
    orc, tr = make_pair(cfg, attn_sigma=0.37, bias_scale=0.05, mode_name=mode_name)
    img, nd, ng = mg.step_inputs(cfg, 0)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    dgr, dl = orc.d_grads(t64(img), t64(nd))
    ggr, gl = orc.g_grads(t64(ng))
    le_d, le_g = _phase_grads(tr, cfg, img, nd, ng)
    strict = mode_name == "fp32_strict"
    assert rel_l2(le_d, dl.numpy()) < (1e-5 if strict else 2e-4) and rel_l2(le_g, gl.numpy()) < (1e-5 if strict else 2e-4)
    errs = {}
    for net, ref in ((tr.D, dgr), (tr.G, ggr)):
        for k, p in net.named_parameters_by_oracle_name():
            if k.endswith("phi.bias"):
                continue
            errs[("D." if net is tr.D else "G.") + k] = rel_l2(p.grad.cpu().numpy(), ref[k].numpy())
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("down-sampled attention,", mode_name, "worst gradient rel-L2 %.2e at %s, median %.2e" % (
        worst[1], worst[0], float(np.median(list(errs.values())))))
    if strict:
        assert worst[1] < 2e-4, worst
    else:
        assert worst[1] < 3e-2 and float(np.median(list(errs.values()))) < 8e-3, worst


def test_conditional_128_forward_and_gradients():
    """BASELINE.json configs[3]: 128x128 class-conditional SAGAN (one-hot concat in G, projection head in D,
    attention at 32x32 and 64x64), B = 2, against the fp64 oracle (FP32_STRICT tier)."""
    cfg = dict(mg.TEST_CFG, img_size=128, use_label=True, num_classes=10, batch_size=2, attn_dim_G=[32, 64])
    orc, tr = make_pair(cfg, attn_sigma=0.3, bias_scale=0.05)
    # the projection-head embedding is created lazily by the GPU discriminator: give both sides the oracle's values
    rng = np.random.Generator(np.random.PCG64(5))
    img = rng.uniform(-1, 1, (2, 128, 128, 3)).astype(np.float32)
    nd = rng.standard_normal((2, cfg["z_dim"])).astype(np.float32)
    ng = rng.standard_normal((2, cfg["z_dim"])).astype(np.float32)
    labels = np.array([3, 7], dtype=np.int64)
    fake_labels = np.array([1, 9], dtype=np.int64)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    dgr, dl = orc.d_grads(t64(img), t64(nd), torch.tensor(labels), torch.tensor(fake_labels))
    ggr, gl = orc.g_grads(t64(ng), torch.tensor(fake_labels))
    le_d, le_g = _phase_grads(tr, cfg, img, nd, ng, labels, fake_labels)
    assert le_d.shape == (2, 1)                                  # discriminator.py:33 -> [B, 1]
    assert rel_l2(le_d, dl.numpy()) < 1e-5 and rel_l2(le_g, gl.numpy()) < 1e-5
    worst = 0.0
    for net, ref in ((tr.D, dgr), (tr.G, ggr)):
        for k, p in net.named_parameters_by_oracle_name():
            if k.endswith("phi.bias") or ref[k] is None:
                continue
            e = rel_l2(p.grad.cpu().numpy(), ref[k].numpy())
            worst = max(worst, e)
            assert e < 2e-4, (k, e)
    print("128x128 conditional: worst per-parameter gradient rel-L2 %.2e" % worst)


@pytest.mark.parametrize("mode_name", MODES)
def test_train_steps_match_oracle_fp32(mode_name):
    """5 full steps (both Adam updates, LR schedule evaluated on the device) against the fp32 oracle: losses and
    weights, in both math modes (bf16_tc is what bench.py times)."""
    cfg = dict(mg.TEST_CFG)
    orc, tr = make_pair(cfg, attn_sigma=0.2, bias_scale=0.02, dtype=torch.float32, steps_per_epoch=2, mode_name=mode_name)
    for s in range(5):
        img, nd, ng = mg.step_inputs(cfg, s)
        ref = orc.train_step(torch.tensor(img), [torch.tensor(nd)], torch.tensor(ng))
        tr.train_step(cu(img), None, [cu(nd)], cu(ng))
        got = tr.losses()
        # free-running: the reduced-precision mode follows the fp32 oracle to 1e-3 for the first two steps; after that
        # Adam(beta_1 = 0) has amplified the 1e-3-level attention-gradient differences (the teacher-forced 100-step test
        # below holds 1e-3 at EVERY step)
        tol = 1e-3 if (mode_name == "fp32_strict" or s < 2) else 2e-2
        assert abs(got["D_loss"] - ref["D_loss"]) < tol and abs(got["G_loss"] - ref["G_loss"]) < tol, (s, got, ref)
    assert tr.opt_G.iterations == 5 == orc.opt_G.iterations and tr.opt_D.iterations == 5
    assert abs(float(tr.opt_D.hyper[0]) - tr.opt_D.lr_t(4)) < 1e-6 * tr.opt_D.lr_t(4)    # staircase: 0.99^(4 // 2)
    worst = 0.0
    for net, ref_params in ((tr.G, orc.G), (tr.D, orc.D)):
        for k, p in net.named_parameters_by_oracle_name():
            e = rel_l2(p.detach().cpu().numpy(), ref_params[k].numpy())
            worst = max(worst, e)
            # Adam(beta_1 = 0) turns the sign of a rounding-sized gradient into a +-lr step: the tiny attention biases
            # / gamma can sit at a few lr after 5 steps in the reduced-precision mode
            assert e < (2e-3 if mode_name == "fp32_strict" else (2e-2 if p.numel() > 64 else 2e-1)), (k, e)
    print(mode_name, "worst weight rel-L2 after 5 steps: %.2e" % worst)


def _sync_from_oracle(tr, orc):
    """Copy the oracle's complete training state into the GPU trainer: parameters, spectral-norm u, Adam second
    moments and step counters (G's BatchNorm runs on batch statistics in training, so moving stats do not matter)."""
    tr.G.load_keras_weights({k: v.numpy() for k, v in orc.G.items()},
                            {k: v.numpy() for k, v in orc.G_sn.items() if k.endswith(".u")})
    tr.D.load_keras_weights({k: v.numpy() for k, v in orc.D.items()},
                            {k: v.numpy() for k, v in orc.D_sn.items() if k.endswith(".u")})
    for net, opt, oopt in ((tr.G, tr.opt_G, orc.opt_G), (tr.D, tr.opt_D, orc.opt_D)):
        base = net.flat_params.data_ptr()
        for k, p_ in net.named_parameters_by_oracle_name():
            off = (p_.data_ptr() - base) // 4
            opt.v[off:off + p_.numel()].copy_(oopt.v[k].reshape(-1).to(torch.float32))
        opt.iterations = oopt.iterations


@pytest.mark.parametrize("mode_name", MODES)
def test_loss_trajectory_100_steps_teacher_forced(mode_name):
    """Per-step G and D hinge losses within 1e-3 of the oracle's over 100 training steps (BASELINE.json tolerance).

    The comparison is made ALONG the oracle's trajectory: before every step the GPU trainer receives the oracle's
    state (weights, spectral-norm u, Adam moments, step counters), both take the same step on the same batch and
    noise, and the reported losses must agree.  A free-running comparison cannot hold this tolerance for ANY two
    fp32 implementations at this configuration: Adam with beta_1 = 0 is a sign-like update in its first steps, and
    the fp32 oracle run with 1 thread instead of 8 (summation order only) already differs from itself by 1.5e-3 at
    step 5 and by O(1) at step 30 ("Loss-trajectory parity"); see the free-running test below.
    Runs in both math modes: bf16_tc is the mode bench.py times."""
    cfg = dict(mg.TEST_CFG)
    orc, tr = make_pair(cfg, attn_sigma=0.0, bias_scale=0.0, dtype=torch.float32, steps_per_epoch=40, mode_name=mode_name)
    worst = worst_w = 0.0
    worst_k = None
    for s in range(100):
        img, nd, ng = mg.step_inputs(cfg, s)
        if s:
            _sync_from_oracle(tr, orc)
        ref = orc.train_step(torch.tensor(img), [torch.tensor(nd)], torch.tensor(ng))
        tr.train_step(cu(img), None, [cu(nd)], cu(ng))
        got = tr.losses()
        e = max(abs(got["D_loss"] - ref["D_loss"]), abs(got["G_loss"] - ref["G_loss"]))
        worst = max(worst, e)
        assert e < 1e-3, (s, got, ref)
        if s % 10 == 9:      # the post-step weights agree as well (Adam, LR schedule, SN state)
            for net, ref_params in ((tr.G, orc.G), (tr.D, orc.D)):
                for k, p_ in net.named_parameters_by_oracle_name():
                    if k.endswith("attn.phi.bias"):
                        # softmax is invariant to a shift of all keys, so d(phi bias) is exactly 0 in exact arithmetic:
                        # its computed gradient is rounding noise, which Adam(beta_1 = 0) turns into +-lr steps
                        continue
                    e_w = rel_l2(p_.detach().cpu().numpy(), ref_params[k].numpy())
                    if e_w > worst_w:
                        worst_w, worst_k = e_w, (s, k)
    print(mode_name, "worst |loss - oracle| over 100 teacher-forced steps: %.2e; worst post-step weight rel-L2: %.2e at %s" % (worst, worst_w, worst_k))
    # the worst parameter is always G's zero-initialised dense bias in the first steps (|b| ~ a few lr, so one sign-like
    # Adam(beta_1 = 0) step taken the other way on a few elements is a large RELATIVE error): measured 2.6e-4 in
    # FP32_STRICT and 1.5-2.1e-3 from run to run in BF16_TC (atomics order), every other parameter < 1e-3
    assert worst_w < (2e-3 if mode_name == "fp32_strict" else 6e-3)


@pytest.mark.parametrize("mode_name", MODES)
def test_loss_trajectory_free_running_vs_golden(mode_name):
    """Free-running (no state sync) against the committed golden trajectory: the first steps, before fp32
    summation-order noise is amplified, stay within 1e-3; afterwards the losses must stay finite and in range."""
    path = os.path.join(GOLD, "trajectory.npz")
    if not os.path.exists(path):
        pytest.skip("trajectory.npz not generated (python tests/golden/make_golden.py --traj)")
    gold = np.load(path)
    cfg = dict(mg.TEST_CFG)
    orc, tr = make_pair(cfg, attn_sigma=0.0, bias_scale=0.0, dtype=torch.float32, steps_per_epoch=40, mode_name=mode_name)
    devs = []
    for s in range(20):
        img, nd, ng = mg.step_inputs(cfg, s)
        tr.train_step(cu(img), None, [cu(nd)], cu(ng))
        got = tr.losses()
        devs.append(max(abs(got["D_loss"] - gold["D_loss"][s]), abs(got["G_loss"] - gold["G_loss"][s])))
        assert np.isfinite(got["D_loss"]) and np.isfinite(got["G_loss"])
    print(mode_name, "free-running |loss - golden| per step:", ["%.1e" % d for d in devs])
    assert max(devs[:3]) < 1e-3


def test_generator_inference_forward_matches_oracle():
    """generator(..., training=False) -- the sample dumps of sagan/main.py:333: BatchNormalization on its moving
    statistics, spectral normalisation of the CURRENT kernels with the stored u / v (sagan_sn_plan_refresh: no power
    iteration), nothing advanced."""
    cfg = dict(mg.TEST_CFG)
    orc, tr = make_pair(cfg, attn_sigma=0.37, bias_scale=0.05)
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    with torch.no_grad():
        for s in range(3):                                    # a few training forwards move the statistics and u
            _, _, ng = mg.step_inputs(cfg, s)
            onets.generator_forward(orc.G, orc.G_sn, t64(ng), cfg, None, True, orc.bn_stats)
            tr.G([cu(ng), None], training=True)
        # an optimiser step later: the kernels have moved, u / v are those of the last training forward
        rng = np.random.Generator(np.random.PCG64(99))
        for k in orc.G:
            orc.G[k] = orc.G[k] + torch.tensor(rng.standard_normal(tuple(orc.G[k].shape)) * 2e-3, dtype=torch.float64)
        tr.G.load_keras_weights({k: v.numpy() for k, v in orc.G.items()})
        _, _, ng = mg.step_inputs(cfg, 7)
        before = [b.clone() for b in _bn_buffers(tr)] + [tr.G.sn_group.u(i).clone() for i in range(len(tr.G.sn_group.shapes))] \
            + [tr.G.sn_group.v(i).clone() for i in range(len(tr.G.sn_group.shapes))]
        ref = onets.generator_forward(orc.G, orc.G_sn, t64(ng), cfg, None, False, orc.bn_stats)
        img = tr.G([cu(ng), None], training=False)
        again = tr.G([cu(ng), None], training=False)
    torch.cuda.synchronize()
    assert rel_l2(img.cpu().numpy(), ref.numpy()) < 2e-5
    assert torch.equal(img, again)
    n_sn = len(tr.G.sn_group.shapes)
    after = [b for b in _bn_buffers(tr)] + [tr.G.sn_group.u(i) for i in range(n_sn)] + [tr.G.sn_group.v(i) for i in range(n_sn)]
    for a, b in zip(before, after):
        assert torch.equal(a, b)                              # neither the moving statistics nor u / v moved
    # and it is not the training-mode forward
    with torch.no_grad():
        assert rel_l2(tr.G([cu(ng), None], training=True).cpu().numpy(), ref.numpy()) > 1e-3


def _bn_buffers(tr):
    return [t for m in tr.G.modules() if hasattr(m, "moving_mean") for t in (m.moving_mean, m.moving_var)]


@pytest.mark.parametrize("mode_name", MODES)
def test_cuda_graph_step_equals_eager_step(mode_name):
    """The captured whole-step graph (what bench.py times) IS the eager step: two trainers built from the same seed,
    one captured (`capture(static_noise=True)`: latent noise read from static buffers), fed the same images and noise.
    Capturing must not train (its warm-up runs on throw-away state).

    What "equal" can mean: the weight-gradient kernels accumulate fp32 partial sums with atomics, so two EAGER runs
    already differ in the last bits (measured 2e-6 rel-L2 on a gradient bucket).  After D's Adam update those bits can
    (i) turn a rounding-sized gradient into a +-lr step (beta_1 = 0: sign-like first steps) and (ii) once in a few
    runs move ONE LeakyReLU pre-activation of D's last 4x4 map across zero, which changes every upstream G-phase
    gradient by ~0.9 / sqrt(8192) = 1e-2 (tools/det2.py reproduces both between two eager trainers).  Hence:

      part 1 (real learning rates, one step): losses, D-phase gradient bucket, Adam second moments, schedule output,
             step counters, spectral-norm state and BatchNorm statistics agree to summation-order noise; the weights
             are identical except for (i): |delta| <= 2 lr on < 0.1 % of the elements; the G-phase bucket within (ii);
      part 2 (learning rates ~0, so both trainers keep the same weights; sagan_deterministic_forward(1), so no
             activation is accumulated with atomics and no mask can flip): four consecutive steps with fresh inputs --
             every replay of the graph reproduces the eager losses to 1e-6 and BOTH gradient buckets to 1e-5.
    """
    from sagan_b200 import _lib
    cfg = dict(mg.TEST_CFG)
    B = cfg["batch_size"]
    g = torch.Generator(device="cuda").manual_seed(99)

    def inputs():
        img = torch.rand(B, cfg["img_size"], cfg["img_size"], 3, device="cuda", generator=g) * 2 - 1
        return img, [torch.randn(B, cfg["z_dim"], device="cuda", generator=g)], torch.randn(B, cfg["z_dim"], device="cuda", generator=g)

    def rel(x, y):
        return float((x - y).norm() / (y.norm() + 1e-30))

    # ---- part 1
    a = make_trainer(cfg, mode_name, seed=3, steps_per_epoch=2)
    b = make_trainer(cfg, mode_name, seed=3, steps_per_epoch=2)
    before = [t.clone() for t in b._state_tensors()]
    b.capture(warmup=3, static_noise=True)
    for t0, t1 in zip(before, b._state_tensors()):
        assert torch.equal(t0, t1)                           # capture() left weights / optimiser / SN / BN state alone
    for t0, t1 in zip(a._state_tensors(), b._state_tensors()):
        assert torch.equal(t0, t1)                           # same seed -> same initial state
    img, nd, ng = inputs()
    a.train_step(img, noises_d=nd, noise_g=ng)
    b.graph_step(img, noises_d=nd, noise_g=ng)
    la, lb = a.losses(), b.losses()
    assert abs(la["D_loss"] - lb["D_loss"]) < 1e-6 * max(1, abs(la["D_loss"])), (la, lb)
    assert abs(la["G_loss"] - lb["G_loss"]) < 2e-5 * max(1, abs(la["G_loss"])), (la, lb)     # through the updated D
    assert a.opt_G.iterations == b.opt_G.iterations == 1 and a.opt_D.iterations == b.opt_D.iterations == 1
    assert torch.equal(a.opt_G.hyper, b.opt_G.hyper) and torch.equal(a.opt_D.hyper, b.opt_D.hyper)
    assert rel(a.D.flat_grads, b.D.flat_grads) < 1e-5
    assert rel(a.G.flat_grads, b.G.flat_grads) < 5e-2        # (ii); part 2 holds this bucket to 1e-5
    assert rel(a.opt_D.v, b.opt_D.v) < 1e-5 and rel(a.opt_G.v, b.opt_G.v) < 1e-1
    assert rel(a.G.sn_group.out, b.G.sn_group.out) < 1e-6 and rel(a.D.sn_group.out, b.D.sn_group.out) < 1e-5
    for ta, tb in zip(_bn_buffers(a), _bn_buffers(b)):
        assert torch.allclose(ta, tb, rtol=1e-5, atol=1e-7)
    for (na, nb, lr) in ((a.G, b.G, cfg["lr_g"]), (a.D, b.D, cfg["lr_d"])):
        diff = (na.flat_params - nb.flat_params).abs()
        assert float(diff.max()) <= 2.001 * lr, float(diff.max())
        if na is a.D:      # D's update follows the (agreeing) D-phase bucket; G's follows the G-phase bucket, see (ii)
            assert float((diff > 1e-6).float().mean()) < 1e-3
    # ---- part 2
    frozen = dict(cfg, lr_g=1e-12, lr_d=1e-12)
    _lib.load().sagan_deterministic_forward(1)
    try:
        a = make_trainer(frozen, mode_name, seed=4)
        b = make_trainer(frozen, mode_name, seed=4)
        b.capture(warmup=3, static_noise=True)
        for step in range(4):
            img, nd, ng = inputs()
            a.train_step(img, noises_d=nd, noise_g=ng)
            b.graph_step(img, noises_d=nd, noise_g=ng)
            la, lb = a.losses(), b.losses()
            assert abs(la["D_loss"] - lb["D_loss"]) < 1e-6 * max(1, abs(la["D_loss"])), (step, la, lb)
            assert abs(la["G_loss"] - lb["G_loss"]) < 1e-6 * max(1, abs(la["G_loss"])), (step, la, lb)
            e_d, e_g = rel(a.D.flat_grads, b.D.flat_grads), rel(a.G.flat_grads, b.G.flat_grads)
            assert e_d < 1e-5 and e_g < 1e-5, (step, e_d, e_g)
            assert a.opt_G.iterations == b.opt_G.iterations == step + 1
            assert rel(a.G.sn_group.out, b.G.sn_group.out) < 1e-6 and rel(a.D.sn_group.out, b.D.sn_group.out) < 1e-6
        assert torch.allclose(a.G.flat_params, b.G.flat_params, rtol=0, atol=1e-9)
        assert torch.allclose(a.D.flat_params, b.D.flat_params, rtol=0, atol=1e-9)
    finally:
        _lib.load().sagan_deterministic_forward(0)
    # ---- the graph draws its own noise when none is injected (bench.py's mode) and stays finite
    c = make_trainer(cfg, mode_name, seed=3)
    c.capture(warmup=3)
    c.graph_step(img)
    lc = c.losses()
    assert np.isfinite(lc["D_loss"]) and np.isfinite(lc["G_loss"])
    with pytest.raises(RuntimeError, match="static noise"):
        c.graph_step(img, noise_g=ng)


# ---------------------------------------------------------------------------------------- residual topologies
RES_CFG = dict(model="resnet", z_dim=16, gf_dim=8, df_dim=8, img_size=32, num_classes=6, use_label=True, use_attention=True,
               attn_dim_G=[16], batch_size=2, lr_g=2e-4, lr_d=7e-4, decay_rate=0.99, update_ratio=1)


def _res_pair(cfg, mode_name, seed=0):
    """Oracle parameters / spectral-norm state and GPU residual networks holding the same values."""
    from oracle import resnets as ores
    from sagan_b200 import nets as gnets
    from sagan_b200 import nn as snn
    gs, ds = ores.res_generator_spec(cfg), ores.res_discriminator_spec(cfg)
    pg = onets.init_params(gs, seed, torch.float64, attn_sigma=0.3, bias_scale=0.05)
    pd = onets.init_params(ds, seed + 100, torch.float64, attn_sigma=0.3, bias_scale=0.05)
    sg, sd = ores.init_res_sn_state(gs, seed + 1), ores.init_res_sn_state(ds, seed + 101)
    snn.set_default_math_mode(_math(mode_name))
    try:
        G, D = gnets.get_res_generator(cfg), gnets.get_res_discriminator(cfg)
        B = cfg["batch_size"]
        with torch.no_grad():       # build pass
            lab = torch.zeros(B, dtype=torch.int64, device="cuda")
            D([G([torch.zeros(B, cfg["z_dim"], device="cuda"), lab]), lab])
    finally:
        snn.set_default_math_mode(_math("fp32_strict"))
    G.load_keras_weights({k: v.numpy() for k, v in pg.items()}, {k: v.numpy() for k, v in sg.items()})
    D.load_keras_weights({k: v.numpy() for k, v in pd.items()}, {k: v.numpy() for k, v in sd.items()})
    return (pg, sg, pd, sd), (G, D)


def test_resnet_inventory_and_reference_shapes():
    """The reference's own tests for these models assert output SHAPES: [B,128,128,3] (test/test_generator.py:26) and
    [B,1] (test/test_discriminator.py:28).  Same here at the reference's size, plus the parameter inventory against the
    oracle's restatement of models/generator.py:23-43 / models/discriminator.py:40-57."""
    from oracle import resnets as ores
    from sagan_b200 import nets as gnets
    cfg = dict(RES_CFG, img_size=128, gf_dim=16, df_dim=16, z_dim=128, num_classes=10, attn_dim_G=[32])
    G, D = gnets.get_res_generator(cfg), gnets.get_res_discriminator(cfg)
    B = 2
    lab = torch.tensor([3, 7], device="cuda")
    with torch.no_grad():
        img = G([torch.randn(B, 128, device="cuda"), lab])
        out = D([img, lab])
    assert tuple(img.shape) == (B, 128, 128, 3) and tuple(out.shape) == (B, 1)
    assert float(img.abs().max()) <= 1.0
    assert {k: tuple(v.shape) for k, v in G.named_parameters_by_oracle_name()} == dict(ores.res_generator_spec(cfg))
    assert {k: tuple(v.shape) for k, v in D.named_parameters_by_oracle_name()} == dict(ores.res_discriminator_spec(cfg))
    assert [k for k, _ in G.sn_by_oracle_name()] == list(ores.res_sn_keys(ores.res_generator_spec(cfg)))
    assert [k for k, _ in D.sn_by_oracle_name()] == list(ores.res_sn_keys(ores.res_discriminator_spec(cfg)))
    assert len(G.attn) == 1 and len(D.attn) == 1          # attention at 32x32 only (models/generator.py:34, discriminator.py:42)


@pytest.mark.parametrize("mode_name", MODES)
def test_resnet_forward_and_gradients_vs_oracle(mode_name):
    """Residual G / D (3x3 SN convs, Conv2DTranspose(3,2) with bias, BN + ReLU, residual adds, attention, projection head
    with SN Embedding) against the fp64 oracle: D(real) logits and all D gradients, D(G(z)) and all G gradients."""
    from oracle import resnets as ores
    cfg = dict(RES_CFG)
    (pg, sg, pd, sd), (G, D) = _res_pair(cfg, mode_name)
    B = cfg["batch_size"]
    rng = np.random.Generator(np.random.PCG64(81))
    img = rng.uniform(-1, 1, (B, 32, 32, 3))
    z = rng.standard_normal((B, cfg["z_dim"]))
    lab, flab = np.array([1, 4]), np.array([5, 0])
    cot = rng.standard_normal((B, 1))
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    # ---- oracle
    for v in pd.values():
        v.requires_grad_(True)
    d_real_ref = ores.res_discriminator_forward(pd, sd, t64(img), torch.tensor(lab), cfg, True)
    dg_ref = dict(zip(pd.keys(), torch.autograd.grad((d_real_ref * t64(cot)).sum(), list(pd.values()))))
    for v in pd.values():
        v.requires_grad_(False)
    for v in pg.values():
        v.requires_grad_(True)
    fake_ref = ores.res_generator_forward(pg, sg, t64(z), torch.tensor(flab), cfg, True, None)
    d_fake_ref = ores.res_discriminator_forward(pd, sd, fake_ref, torch.tensor(flab), cfg, True)
    gg_ref = dict(zip(pg.keys(), torch.autograd.grad((d_fake_ref * t64(cot)).sum(), list(pg.values()))))
    # ---- GPU
    D.zero_grad_flat()
    d_real = D([cu(img), torch.tensor(lab).cuda()], training=True)
    d_real.backward(cu(cot))
    torch.cuda.synchronize()
    strict = mode_name == "fp32_strict"
    # measured worst 1.0e-6 / 7.6e-4, median 5e-7 / 6.5e-5.  In the tensor-core mode ONE ReLU pre-activation of D's last
    # 4x4 map (2 048 elements here) landing on the other side of zero costs 0.9 / sqrt(2048) = 2e-2 on everything upstream
    # (DESIGN.md section 2): the worst-case bound leaves room for that rare event, the median is held tight
    tol_f, tol_g = (1e-5, 2e-4) if strict else (2e-4, 5e-2)
    assert rel_l2(d_real.detach().cpu().numpy(), d_real_ref.detach().numpy()) < tol_f
    def grad_errs(tag, net, ref):
        """Biases that feed a BatchNormalization (models/generator.py:11-13: the batch mean removes them) and the attention
        key bias have an exactly-zero gradient: the computed value is rounding noise and is only bounded."""
        out = {}
        for k, p in net.named_parameters_by_oracle_name():
            r, g = ref[k].numpy(), p.grad.cpu().numpy()
            if np.linalg.norm(r) < 1e-9 * max(1.0, np.sqrt(r.size)):
                assert np.linalg.norm(g) < 1e-4, (k, np.linalg.norm(g))
                continue
            out[tag + k] = rel_l2(g, r)
        return out

    errs = grad_errs("D.", D, dg_ref)
    G.zero_grad_flat()
    for p in D.parameters():
        p.requires_grad_(False)
    fake = G([cu(z), torch.tensor(flab).cuda()], training=True)
    d_fake = D([fake, torch.tensor(flab).cuda()], training=True)
    d_fake.backward(cu(cot))
    torch.cuda.synchronize()
    assert rel_l2(fake.detach().cpu().numpy(), fake_ref.detach().numpy()) < tol_f
    assert rel_l2(d_fake.detach().cpu().numpy(), d_fake_ref.detach().numpy()) < tol_f
    errs.update(grad_errs("G.", G, gg_ref))
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("resnet", mode_name, "worst gradient rel-L2 %.2e at %s, median %.2e" % (worst[1], worst[0], float(np.median(list(errs.values())))))
    assert worst[1] < tol_g, worst
    assert float(np.median(list(errs.values()))) < (1e-5 if strict else 2e-2)
    for k, m in G.sn_by_oracle_name() + D.sn_by_oracle_name():
        ref = (sg if k in sg and m in [w for _, w in G.sn_by_oracle_name()] else sd)[k]
        assert rel_l2(m.u.cpu().numpy(), ref.numpy()) < 1e-5, k


def test_resnet_train_step_runs_through_the_trainer():
    """`model: 'resnet'` (the branch sagan/main.py:104-107 leaves disabled) through Trainer: eager and captured steps."""
    tr = make_trainer(dict(RES_CFG), "bf16_tc", seed=1)
    B = RES_CFG["batch_size"]
    img = torch.rand(B, 32, 32, 3, device="cuda") * 2 - 1
    lab = torch.tensor([2, 5], device="cuda")
    tr.train_step(img, lab)
    l0 = tr.losses()
    tr.capture()
    tr.graph_step(img, lab)
    l1 = tr.losses()
    assert all(np.isfinite(v) for v in list(l0.values()) + list(l1.values()))
    assert tr.opt_G.iterations == 2 and tr.opt_D.iterations == 2


def test_uint8_record_input_contract():
    """SURVEY.md §8f row 4 / sagan/dataset.py:27-40: a step fed the raw uint8 records equals the step fed the decoded
    float images (eager and captured graph), with sagan_deterministic_forward so that nothing but the input path differs."""
    from sagan_b200 import _lib
    from oracle import weightnorm as own
    cfg = dict(mg.TEST_CFG, lr_g=1e-12, lr_d=1e-12)
    B = cfg["batch_size"]
    rng = np.random.Generator(np.random.PCG64(71))
    raw = rng.integers(0, 256, (B, 64, 64, 3), dtype=np.uint8)
    nd, ng = [cu(rng.standard_normal((B, 128)))], cu(rng.standard_normal((B, 128)))
    _lib.load().sagan_deterministic_forward(1)
    try:
        a, b, c = (make_trainer(cfg, "bf16_tc", seed=6) for _ in range(3))
        a.train_step(cu(own.decode_records(raw)), noises_d=nd, noise_g=ng)
        b.train_step(torch.tensor(raw).cuda(), noises_d=nd, noise_g=ng)
        c.capture(static_noise=True, uint8_input=True)
        c.graph_step(torch.tensor(raw).pin_memory(), noises_d=nd, noise_g=ng)
        la, lb, lc = a.losses(), b.losses(), c.losses()
        for k in la:
            assert abs(la[k] - lb[k]) < 1e-6 and abs(la[k] - lc[k]) < 1e-6, (k, la, lb, lc)
        for t in (b, c):
            assert float((a.D.flat_grads - t.D.flat_grads).norm() / a.D.flat_grads.norm()) < 1e-5
        with pytest.raises(RuntimeError, match="uint8"):
            c.graph_step(cu(own.decode_records(raw)))
    finally:
        _lib.load().sagan_deterministic_forward(0)


@pytest.mark.parametrize("mode_name", MODES)
def test_overlapped_step_equals_single_stream_step(mode_name):
    """The step whose generator forwards run as a side-stream branch (Trainer overlap_streams=True, the default and
    what bench.py times) computes what the single-stream step computes: same injected noise, the losses of three
    steps and the weights after them.  The first step's losses see identical weights (agreement to rounding); after
    that the two runs differ by the order of the atomic partial sums in the weight-gradient kernels, which Adam with
    beta_1 = 0 turns into +-lr steps on parameters whose gradient is rounding noise (loss-trajectory
    parity), hence the looser bounds.  Repeated to give a stream race more than one chance to show."""
    cfg = dict(mg.TEST_CFG)
    B = cfg["batch_size"]
    for rep in range(3):
        a = make_trainer(cfg, mode_name, seed=5, overlap_streams=True)
        b = make_trainer(cfg, mode_name, seed=5, overlap_streams=False)
        b.G.flat_params.copy_(a.G.flat_params); b.D.flat_params.copy_(a.D.flat_params)
        b.G.sn_group.out.copy_(a.G.sn_group.out); b.D.sn_group.out.copy_(a.D.sn_group.out)
        g = torch.Generator(device="cuda").manual_seed(11 + rep)
        for step in range(3):
            img = torch.rand(B, cfg["img_size"], cfg["img_size"], 3, device="cuda", generator=g) * 2 - 1
            nd = [torch.randn(B, cfg["z_dim"], device="cuda", generator=g)]
            ng = torch.randn(B, cfg["z_dim"], device="cuda", generator=g)
            a.train_step(img, noises_d=nd, noise_g=ng)
            b.train_step(img, noises_d=nd, noise_g=ng)
            la, lb = a.losses(), b.losses()
            tol = 1e-5 if step == 0 else (2e-3 if mode_name == "fp32_strict" else 1e-2)
            assert abs(la["D_loss"] - lb["D_loss"]) < tol and abs(la["G_loss"] - lb["G_loss"]) < tol, (rep, step, la, lb)
        for pa, pb in ((a.G.flat_params, b.G.flat_params), (a.D.flat_params, b.D.flat_params)):
            rel = float((pa - pb).norm() / pb.norm())
            assert rel < (2e-3 if mode_name == "fp32_strict" else 5e-3), (rep, rel)
