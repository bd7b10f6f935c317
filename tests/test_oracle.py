"""CPU tests of the oracle itself (no GPU): golden fixtures, analytic-vs-autograd gradients, finite
differences, fp32-vs-fp64 agreement.  The reference holds no golden vectors for this path: the forward path is
pinned to the reference's own code in tests/test_reference_vectors.py, the gradients and the optimiser here, by
self-consistency (see oracle/__init__.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402

from oracle import attention, nets, sn, train  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def test_sn_golden_and_identities():
    gold = np.load(os.path.join(GOLD, "sn.npz"))
    for i, (R, K) in enumerate(mg.SN_SHAPES):
        W, u, dW = mg.sn_inputs(R, K, 100 + i)
        W64, u64 = W.astype(np.float64), u.astype(np.float64)
        for Ip, factor in ((1, None), (2, 1.5)):
            u2, v2, sig, Wb = sn.power_iteration(W64, u64, Ip, factor)
            tag = f"{R}x{K}_Ip{Ip}"
            assert rel(u2, gold[tag + "_u"]) < 1e-13 and rel(v2, gold[tag + "_v"]) < 1e-13
            assert abs(sig - gold[tag + "_sigma"]) <= 1e-13 * abs(sig)
            assert rel(mg.summarize(Wb), gold[tag + "_Wbar_sum"]) < 1e-13
            # sigma shortcut used by the CUDA kernel: ||t||^2 / (||t|| + eps), t = v W^T
            t = v2 @ sn.matricize(W64).T
            nt = np.linalg.norm(t)
            f = factor or 1.0
            assert abs(sig - nt * nt / (nt + 1e-12) / f) < 1e-12 * abs(sig)
            assert abs(np.linalg.norm(u2) - 1) < 1e-9 and abs(np.linalg.norm(v2) - 1) < 1e-9


def test_sn_raw_reshape_is_not_a_transpose():
    """layers.py:56: reshape(W, [W.shape[-1], -1]) reinterprets memory; its sigma differs from the true matricisation's."""
    rng = np.random.Generator(np.random.PCG64(0))
    W = rng.standard_normal((4, 4, 64, 128))
    raw = np.linalg.svd(sn.matricize(W), compute_uv=False)[0]
    true = np.linalg.svd(W.reshape(-1, 128).T, compute_uv=False)[0]
    assert abs(raw - true) > 1e-3


def test_sn_backward_finite_difference():
    rng = np.random.Generator(np.random.PCG64(1))
    W = rng.standard_normal((3, 3, 5, 7))
    u, _ = sn.make_param(W, rng)
    u2, v2, sig, Wb = sn.power_iteration(W, u, 1, 2.0)
    dWb = rng.standard_normal(W.shape)
    g = sn.backward(dWb, Wb, u2, v2, sig, 2.0)
    L = lambda Wx: np.sum(dWb * (Wx / (np.sum((u2 @ sn.matricize(Wx)) * v2) / 2.0)))
    for idx in [(0, 0, 0, 0), (2, 1, 4, 6), (1, 2, 3, 5)]:
        Wp, Wn = W.copy(), W.copy()
        Wp[idx] += 1e-6
        Wn[idx] -= 1e-6
        assert abs((L(Wp) - L(Wn)) / 2e-6 - g[idx]) < 1e-6 * max(1, abs(g[idx]))


def test_sn_ip_validation():
    import pytest
    with pytest.raises(ValueError):
        sn.power_iteration(np.ones((2, 3)), np.ones((1, 3)), Ip=0)


def test_attention_golden_autograd_and_fp32():
    gold = np.load(os.path.join(GOLD, "attention.npz"))
    for i, (B, N, C) in enumerate(mg.ATTN_CASES):
        X, dY, w = mg.attn_inputs(B, N, C, 200 + i)
        w64 = {k: np.asarray(v, np.float64) for k, v in w.items()}
        X64, dY64 = X.astype(np.float64), dY.astype(np.float64)
        Y = attention.forward(X64, **w64)
        g = attention.backward(dY64, X64, **w64)
        tag = f"B{B}_N{N}_C{C}"
        assert rel(Y, gold[tag + "_Y"]) < 1e-6 and rel(g["dX"], gold[tag + "_dX"]) < 1e-6
        # analytic backward == autograd of the literal forward
        tw = {k: torch.tensor(v, requires_grad=True) for k, v in w64.items()}
        Xt = torch.tensor(X64, requires_grad=True)
        phi, th, gg = Xt @ tw["Wphi"] + tw["bphi"], Xt @ tw["Wtheta"] + tw["btheta"], Xt @ tw["Wg"] + tw["bg"]
        P = torch.softmax(th @ phi.transpose(1, 2), -1)
        Yt = Xt + tw["gamma"] * ((P @ gg) @ tw["Wo"] + tw["bo"])
        Yt.backward(torch.tensor(dY64))
        assert rel(Yt.detach().numpy(), Y) < 1e-13 and rel(Xt.grad.numpy(), g["dX"]) < 1e-12
        for k in attention.WEIGHT_NAMES:
            if k == "bphi":
                # the key bias shifts every logit of a row by the same amount: softmax-invariant, gradient == 0
                assert np.abs(g["dbphi"]).max() < 1e-12 * np.abs(g["dbtheta"]).max()
                continue
            assert rel(tw[k].grad.numpy(), g["d" + k]) < 1e-11, k
            assert rel(g["d" + k], gold[tag + "_d" + k]) < 1e-12
        # fp32 evaluation of the same oracle agrees to fp32 accuracy
        Y32 = attention.forward(X, **{k: np.asarray(v, np.float32) for k, v in w.items()})
        assert rel(Y32, Y) < 1e-5


def test_attention_golden_n1024():
    """The in-model token count (N = 1024) fixture added for the tensor-core backward parity tests."""
    gold = np.load(os.path.join(GOLD, "attention_n1024.npz"))
    i, (B, N, C) = 0, mg.ATTN_CASES_N1024[0]
    X, dY, w = mg.attn_inputs(B, N, C, 300 + i)
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    tag = f"B{B}_N{N}_C{C}"
    assert rel(attention.forward(X.astype(np.float64), **w64), gold[tag + "_Y"]) < 1e-6      # stored as fp32
    g = attention.backward(dY.astype(np.float64), X.astype(np.float64), **w64)
    assert rel(g["dX"], gold[tag + "_dX"]) < 1e-6
    for k in attention.WEIGHT_NAMES:
        if k != "bphi":
            assert rel(g["d" + k], gold[tag + "_d" + k]) < 1e-12, k
    assert {f"B{b}_N{n}_C{c}_Y" for b, n, c in mg.ATTN_CASES_N1024} <= set(gold.files)


def test_same_padding_matches_tf_convention():
    assert nets.same_pad(64, 4, 2) == (1, 1, 32)
    assert nets.same_pad(64, 4, 1) == (1, 2, 64)     # k=4, s=1: 1 before / 2 after
    assert nets.same_pad(9, 3, 2) == (1, 1, 5)
    x = torch.zeros(1, 6, 6, 1, dtype=torch.float64)
    x[0, 0, 0, 0] = 1.0
    w = torch.arange(16, dtype=torch.float64).reshape(4, 4, 1, 1)
    y = nets.conv2d_same(x, w, None, 1)
    assert float(y[0, 0, 0, 0]) == float(w[1, 1, 0, 0])   # y[0,0] picks w[1,1]


def test_conv_transpose_is_adjoint_of_conv():
    """<conv(x), y> == <x, conv_transpose(y)> for the Keras 'same' geometry (k=4, s=2)."""
    rng = np.random.Generator(np.random.PCG64(2))
    big = torch.tensor(rng.standard_normal((2, 8, 8, 3)))
    small = torch.tensor(rng.standard_normal((2, 4, 4, 5)))
    w = torch.tensor(rng.standard_normal((4, 4, 3, 5)))
    lhs = (nets.conv2d_same(big, w, None, 2) * small).sum()
    rhs = (big * nets.conv2d_transpose_same(small, w, 2)).sum()
    assert abs(float(lhs - rhs)) < 1e-10 * abs(float(lhs))


def test_nets_golden_shapes_and_counts():
    cfg = mg.TEST_CFG
    gs, ds = nets.generator_spec(cfg), nets.discriminator_spec(cfg)
    assert sum(int(np.prod(s)) for _, s in gs) == 1227638
    assert sum(int(np.prod(s)) for _, s in ds) == 175438
    assert list(nets.sn_keys(gs).values())[:4] == [(4096, 128), (256, 2048), (128, 1024), (64, 512)]
    gold = np.load(os.path.join(GOLD, "nets.npz"))
    tr = train.OracleTrainer(cfg, torch.float64, seed=0, attn_sigma=0.37, bias_scale=0.05, global_batch_size=4)
    img, nd, ng = (torch.tensor(a, dtype=torch.float64) for a in mg.step_inputs(cfg, 0))
    dgr, dl = tr.d_grads(img, nd)
    assert dl.shape == (4, 4, 4, 1)                              # patch logits (discriminator.py:35)
    assert rel(dl.numpy(), gold["D_loss_elems"]) < 1e-12
    for k, v in dgr.items():
        assert rel(mg.summarize(v.numpy()), gold["Dgrad." + k]) < 1e-10, k


def test_keras_adam_and_schedule():
    p = {"w": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    opt = train.KerasAdam(p, 1e-2, decay_steps=2, decay_rate=0.5)
    g = {"w": torch.tensor([0.5, -0.25], dtype=torch.float64)}
    ref, v = np.array([1.0, -2.0]), np.zeros(2)
    for t in range(1, 6):
        lr = 1e-2 * 0.5 ** ((t - 1) // 2)
        v = 0.999 * v + 0.001 * g["w"].numpy() ** 2
        ref = ref - lr * np.sqrt(1 - 0.999 ** t) * g["w"].numpy() / (np.sqrt(v) + 1e-7)
        opt.apply_gradients(p, g)
    assert rel(p["w"].numpy(), ref) < 1e-14


def test_hinge_losses():
    a, b = torch.tensor([0.5, 2.0, -1.0]), torch.tensor([-2.0, 0.0, 3.0])
    assert torch.equal(train.hinge_loss_d(a, b), torch.tensor([0.5, 1.0, 6.0]))
    assert torch.equal(train.hinge_loss_g(b), -b)


def test_pooled_attention_oracle_matches_torch_autograd():
    """Down-sampled keys / values (oracle only so far): the numpy forward / analytic backward
    against torch autograd of the same graph built from max_pool2d, in fp64."""
    import torch
    import torch.nn.functional as TF
    B, H, W, C = 2, 8, 6, 16
    X, dY, w = attention.make_inputs(B, H * W, C, seed=5, gamma=0.6, dtype=np.float64)
    Y = attention.forward_pooled(X, **w, hw=(H, W))
    g = attention.backward_pooled(dY, X, **w, hw=(H, W))
    t = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in w.items()}
    x = torch.tensor(X, requires_grad=True)

    def pool(z):       # [B, N, c] -> [B, N/4, c]
        c = z.shape[-1]
        z4 = z.reshape(B, H, W, c).permute(0, 3, 1, 2)
        return TF.max_pool2d(z4, 2, 2).permute(0, 2, 3, 1).reshape(B, -1, c)

    phi = pool(x @ t["Wphi"] + t["bphi"])
    theta = x @ t["Wtheta"] + t["btheta"]
    gg = pool(x @ t["Wg"] + t["bg"])
    P = torch.softmax(theta @ phi.transpose(1, 2), dim=-1)
    y = x + t["gamma"] * ((P @ gg) @ t["Wo"] + t["bo"])
    assert P.shape == (B, H * W, H * W // 4)
    np.testing.assert_allclose(Y, y.detach().numpy(), rtol=1e-12, atol=1e-12)
    y.backward(torch.tensor(dY))
    np.testing.assert_allclose(g["dX"], x.grad.numpy(), rtol=1e-10, atol=1e-12)
    for k in attention.WEIGHT_NAMES:
        np.testing.assert_allclose(g["d" + k], t[k].grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=k)
    # with one key per 2x2 window the block must differ from the un-pooled one (it is not a no-op) ...
    assert np.abs(Y - attention.forward(X, **w)).max() > 1e-3
    # ... and reduce to it when every window holds four identical tokens
    Xr = np.repeat(np.repeat(X.reshape(B, H, W, C)[:, ::2, ::2], 2, axis=1), 2, axis=2).reshape(B, H * W, C)
    np.testing.assert_allclose(attention.forward_pooled(Xr, **w, hw=(H, W)), attention.forward(Xr, **w), rtol=1e-10, atol=1e-12)


def test_weightnorm_oracle_gradients_and_record_decode():
    """oracle.weightnorm (sagan/layers.py:124,152-194; sagan/dataset.py:31-34) against torch autograd / exact values."""
    from oracle import weightnorm as own
    rng = np.random.Generator(np.random.PCG64(5))
    v, g, dw = rng.standard_normal((3, 3, 5, 7)), rng.uniform(0.5, 2, 7), rng.standard_normal((3, 3, 5, 7))
    tv, tg = torch.tensor(v, requires_grad=True), torch.tensor(g, requires_grad=True)
    w = tv / tv.pow(2).sum(dim=(0, 1, 2), keepdim=True).clamp_min(1e-12).sqrt() * tg
    w.backward(torch.tensor(dw))
    dv, dg = own.backward(dw, v, g)
    assert rel(own.kernel_from_vg(v, g), w.detach().numpy()) < 1e-14
    assert rel(dv, tv.grad.numpy()) < 1e-13 and rel(dg, tg.grad.numpy()) < 1e-13
    assert rel(own.kernel_from_vg(v, own.init_norm(v)), v) < 1e-14            # g = ||v|| reproduces v
    x = rng.standard_normal((4, 6, 6, 7)) * 3 + 2
    g2, b2 = own.data_dep_init(x, np.ones(7), np.zeros(7))
    z = x * g2 + b2
    assert np.abs(z.mean((0, 1, 2))).max() < 1e-12 and np.abs(z.std((0, 1, 2)) - 1).max() < 1e-8
    raw = np.arange(256, dtype=np.uint8)
    dec = own.decode_records(raw)
    assert dec.dtype == np.float32 and dec[0] == -1.0 and abs(dec[255] - 1.0) < 1e-6 and np.all(np.diff(dec) > 0)
