"""GPU parity tests: every kernel of libsagan_b200.so (called through the C ABI via the ctypes
binding) against the CPU oracle on the same seeded inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): sigma / u / v <= 1e-5 relative; attention and model
forward / gradients <= 1e-5 relative L2 in FP32_STRICT mode, <= 2e-3 in BF16_TC mode.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402

from oracle import attention as oattn  # noqa: E402
from oracle import nets as onets  # noqa: E402
from oracle import sn as osn  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STRICT_TOL = 1e-5
TC_TOL = 2e-3


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.fixture(scope="module")
def F():
    import sagan_b200.functional as F
    from sagan_b200 import _lib
    _lib.load()
    return F


# ---------------------------------------------------------------------------------------- spectral norm
@pytest.mark.parametrize("Ip,factor", [(1, None), (2, 1.5)])
def test_sn_power_iteration_all_shapes_one_launch(F, Ip, factor):
    """All in-model SN matrix shapes in ONE multi-tensor launch, vs the fp64 oracle and the golden file."""
    gold = np.load(os.path.join(GOLD, "sn.npz"))
    Ws, us, dWs = [], [], []
    for i, (R, K) in enumerate(mg.SN_SHAPES):
        W, u, dW = mg.sn_inputs(R, K, 100 + i)
        Ws.append(W), us.append(u), dWs.append(dW)
    tW = [cu(W).requires_grad_(True) for W in Ws]
    group = F.SpectralNormGroup(tW, [cu(u) for u in us], Ip, [factor] * len(Ws))
    wbars = group.normalized(update=True)
    loss = sum((wb * cu(dW)).sum() for wb, dW in zip(wbars, dWs))
    loss.backward()
    torch.cuda.synchronize()
    for i, (R, K) in enumerate(mg.SN_SHAPES):
        u2, v2, sig, Wb = osn.power_iteration(Ws[i].astype(np.float64), us[i].astype(np.float64), Ip, factor)
        tag = f"{R}x{K}_Ip{Ip}"
        # oracle == golden
        assert rel_l2(u2, gold[tag + "_u"]) < 1e-12 and abs(sig - gold[tag + "_sigma"]) < 1e-12 * abs(sig)
        assert rel_l2(group.u(i).cpu().numpy(), u2) < STRICT_TOL, tag
        assert rel_l2(group.v(i).cpu().numpy(), v2) < STRICT_TOL, tag
        assert abs(float(group.sigma(i).cpu()) - sig) < STRICT_TOL * abs(sig), tag
        assert rel_l2(wbars[i].detach().cpu().numpy(), Wb) < STRICT_TOL, tag
        g = osn.backward(dWs[i].astype(np.float64), Wb, u2, v2, sig, factor)
        assert rel_l2(tW[i].grad.cpu().numpy(), g) < 2e-5, tag
        assert rel_l2(mg.summarize(tW[i].grad.cpu().numpy()), gold[tag + "_dW_sum"]) < 2e-5, tag


def test_sn_persistence_and_idempotence(F):
    """u persists across calls: k single-iteration calls == one call with Ip = k (size-independent property);
    at the fixed point sigma equals the top singular value."""
    R, K = 96, 200
    W, u, _ = mg.sn_inputs(R, K, 7)
    g1 = F.SpectralNormGroup([cu(W)], [cu(u)], 1)
    for _ in range(3):
        g1.run()
    g3 = F.SpectralNormGroup([cu(W)], [cu(u)], 3)
    g3.run()
    torch.cuda.synchronize()
    assert rel_l2(g1.u(0).cpu().numpy(), g3.u(0).cpu().numpy()) < 1e-6
    assert abs(float(g1.sigma(0).cpu()) - float(g3.sigma(0).cpu())) < 1e-6 * float(g3.sigma(0).cpu())
    for _ in range(300):
        g1.run()
    torch.cuda.synchronize()
    top = np.linalg.svd(osn.matricize(W.astype(np.float64)), compute_uv=False)[0]
    assert abs(float(g1.sigma(0).cpu()) - top) < 1e-4 * top


def test_sn_large_matrix_property(F):
    """Roofline-sized matrix (4096 x 4096): W_bar * sigma == W and ||u|| == ||v|| == 1."""
    R = K = 4096
    gen = torch.Generator(device="cuda").manual_seed(3)
    W = torch.randn(K, R, device="cuda", generator=gen) * 0.02
    u = torch.randn(1, R, device="cuda", generator=gen)
    g = F.SpectralNormGroup([W], [u / u.norm()], 1)
    g.run()
    torch.cuda.synchronize()
    sig = g.sigma(0)
    assert torch.allclose(g.w_bar(0) * sig, W, rtol=1e-6, atol=1e-9)
    assert abs(float(g.u(0).norm()) - 1) < 1e-5 and abs(float(g.v(0).norm()) - 1) < 1e-5
    # against torch fp64 on the device data
    Wm = W.reshape(R, K).double()
    v = (u.double() / u.double().norm()) @ Wm
    v = v / (v.norm() + 1e-12)
    t = v @ Wm.t()
    assert abs(float(sig) - float(t.norm())) < 1e-5 * float(t.norm())


def test_sn_refresh_without_iteration(F):
    """sagan_sn_plan_refresh -- the wrapped layer called with training=False: sigma = sum((u W_mat) * v) [/ factor] from
    the STORED u, v and the CURRENT kernel, W_bar = W / sigma; u and v untouched (layers.py:62-68 without :58-60)."""
    shapes = [(4096, 128), (32, 256), (4, 32), (7, 13), (16, 48)]
    rng = np.random.Generator(np.random.PCG64(23))
    Ws, us = zip(*[mg.sn_inputs(R, K, 400 + i)[:2] for i, (R, K) in enumerate(shapes)])
    factors = [None, 2.0, None, 0.5, None]
    tW = [cu(W) for W in Ws]
    g = F.SpectralNormGroup(tW, [cu(u) for u in us], 1, factors)
    g.run()                                                   # a training call: u, v of the ORIGINAL kernels
    torch.cuda.synchronize()
    u1 = [g.u(i).cpu().numpy().astype(np.float64) for i in range(len(shapes))]
    v1 = [g.v(i).cpu().numpy().astype(np.float64) for i in range(len(shapes))]
    W2 = [W.astype(np.float64) + rng.standard_normal(W.shape) * 0.004 for W in Ws]      # an optimiser step later
    for t, w in zip(tW, W2):
        t.copy_(cu(w))
    g.refresh()
    torch.cuda.synchronize()
    for i, (R, K) in enumerate(shapes):
        w32 = tW[i].cpu().numpy().astype(np.float64)
        sig = np.sum((u1[i].reshape(1, R) @ osn.matricize(w32)) * v1[i].reshape(1, K))
        if factors[i]:
            sig = sig / factors[i]
        assert abs(float(g.sigma(i).cpu()) - sig) < STRICT_TOL * abs(sig), (R, K)
        assert rel_l2(g.w_bar(i).cpu().numpy(), w32 / sig) < STRICT_TOL, (R, K)
        assert np.array_equal(g.u(i).cpu().numpy().astype(np.float64), u1[i])          # not advanced
        assert np.array_equal(g.v(i).cpu().numpy().astype(np.float64), v1[i])


def test_sn_rejects_bad_ip(F):
    W, u, _ = mg.sn_inputs(8, 16, 1)
    with pytest.raises(ValueError, match="positive integer"):
        F.SpectralNormGroup([cu(W)], [cu(u)], 0)


# ---------------------------------------------------------------------------------------- conv family
CONV_CASES = [  # B, H, W, Cin, Cout, k, stride   (the D / G layer shapes at small batch + odd ones)
    (2, 64, 64, 3, 16, 4, 2), (2, 32, 32, 16, 32, 4, 2), (2, 8, 8, 64, 128, 4, 2),
    (2, 64, 64, 16, 3, 4, 1), (2, 4, 4, 128, 1, 4, 1), (3, 9, 7, 5, 6, 3, 2), (2, 16, 16, 16, 2, 1, 1),
    (1, 5, 5, 8, 20, 3, 1),
    # the generator's output conv at other widths (shared-memory-tiled direct kernel: 16 / 8 / 4 output rows per CTA) and
    # at a width the tiled form does not take (48: 512 % 48 != 0 -> thread-per-four-pixels form)
    (2, 32, 32, 16, 3, 4, 1), (1, 128, 128, 16, 3, 4, 1), (1, 64, 48, 16, 3, 4, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_dgrad_wgrad(F, case):
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(11))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cin, Cout)) * 0.1
    b = rng.standard_normal(Cout) * 0.1
    tx, tw, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, b))
    ref = torch.nn.functional.leaky_relu(onets.conv2d_same(tx, tw, tb, s), 0.1)
    dy = rng.standard_normal(tuple(ref.shape))
    ref.backward(torch.tensor(dy))
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    y = F.conv2d(gx, gw, gb, s, "same", F.ACT_LRELU, 0.1)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()) < STRICT_TOL
    assert rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()) < STRICT_TOL
    assert rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()) < STRICT_TOL
    assert rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()) < STRICT_TOL


def test_conv2d_tanh_head(F):
    rng = np.random.Generator(np.random.PCG64(12))
    x = rng.standard_normal((2, 16, 16, 16))
    w = rng.standard_normal((4, 4, 16, 3)) * 0.1
    tx, tw = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w))
    ref = torch.tanh(onets.conv2d_same(tx, tw, None, 1))
    dy = rng.standard_normal(tuple(ref.shape))
    ref.backward(torch.tensor(dy))
    gx, gw = (cu(a).requires_grad_(True) for a in (x, w))
    y = F.conv2d(gx, gw, None, 1, "same", F.ACT_TANH)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()) < STRICT_TOL
    assert rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()) < STRICT_TOL
    assert rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()) < STRICT_TOL


@pytest.mark.parametrize("case", [(2, 4, 4, 256, 128, 4, 2), (2, 16, 16, 64, 32, 4, 2), (2, 32, 32, 32, 16, 4, 2),
                                  (2, 5, 6, 8, 12, 3, 2), (1, 4, 4, 8, 8, 4, 2)])
def test_conv2d_transpose(F, case):
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(13))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cout, Cin)) * 0.1
    tx, tw = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w))
    ref = onets.conv2d_transpose_same(tx, tw, s)
    dy = rng.standard_normal(tuple(ref.shape))
    ref.backward(torch.tensor(dy))
    gx, gw = (cu(a).requires_grad_(True) for a in (x, w))
    y = F.conv2d_transpose(gx, gw, s, "same")
    assert tuple(y.shape) == tuple(ref.shape)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()) < STRICT_TOL
    assert rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()) < STRICT_TOL
    assert rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()) < STRICT_TOL


def test_dense(F):
    rng = np.random.Generator(np.random.PCG64(14))
    x, w, b = rng.standard_normal((8, 128)), rng.standard_normal((128, 4096)) * 0.05, rng.standard_normal(4096)
    tx, tw, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, b))
    ref = tx @ tw + tb
    dy = rng.standard_normal((8, 4096))
    ref.backward(torch.tensor(dy))
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    y = F.dense(gx, gw, gb)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()) < STRICT_TOL
    assert rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()) < STRICT_TOL
    assert rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()) < STRICT_TOL
    assert rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()) < STRICT_TOL


# (the last two: channel counts that are not a multiple of 4 take the scalar apply loops)
@pytest.mark.parametrize("shape", [(4, 8, 8, 128), (2, 64, 64, 16), (3, 5, 7, 12), (2, 9, 5, 2), (3, 4, 5, 7)])
def test_batchnorm_lrelu(F, shape):
    rng = np.random.Generator(np.random.PCG64(15))
    x = rng.standard_normal(shape) * 1.7 + 0.3
    C = shape[-1]
    gam, bet = rng.standard_normal(C) * 0.3 + 1, rng.standard_normal(C) * 0.2
    tx, tg, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, gam, bet))
    stats = {"bn.moving_mean": torch.zeros(C, dtype=torch.float64), "bn.moving_var": torch.ones(C, dtype=torch.float64)}
    ref = torch.nn.functional.leaky_relu(onets.batchnorm_train(tx, tg, tb, stats, "bn"), 0.1)
    dy = rng.standard_normal(shape)
    ref.backward(torch.tensor(dy))
    gx, gg, gb = (cu(a).requires_grad_(True) for a in (x, gam, bet))
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    y = F.batchnorm_lrelu(gx, gg, gb, mm, mv, 1e-3, 0.99, 0.1)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()) < STRICT_TOL
    assert rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()) < 2e-5
    assert rel_l2(gg.grad.cpu().numpy(), tg.grad.numpy()) < 2e-5
    assert rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()) < 2e-5
    assert rel_l2(mm.cpu().numpy(), stats["bn.moving_mean"].numpy()) < STRICT_TOL
    assert rel_l2(mv.cpu().numpy(), stats["bn.moving_var"].numpy()) < STRICT_TOL


@pytest.mark.parametrize("shape", [(4, 8, 8, 128), (3, 5, 7, 12), (2, 9, 5, 2), (3, 4, 5, 7)])
@pytest.mark.parametrize("slope", [0.1, 0.0, 1.0])
def test_batchnorm_lrelu_inference_mode(F, shape, slope):
    """sagan_bn_lrelu_infer: Keras BatchNormalization(training=False) on the moving statistics + LeakyReLU / ReLU / none."""
    rng = np.random.Generator(np.random.PCG64(16))
    C = shape[-1]
    x = rng.standard_normal(shape) * 1.7 + 0.3
    gam, bet = rng.standard_normal(C) * 0.3 + 1, rng.standard_normal(C) * 0.2
    mm, mv = rng.standard_normal(C) * 0.5, rng.uniform(0.3, 2.0, C)
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    stats = {"bn.moving_mean": t(mm), "bn.moving_var": t(mv)}
    ref = torch.nn.functional.leaky_relu(onets.batchnorm_infer(t(x), t(gam), t(bet), stats, "bn"), slope)
    gmm, gmv = cu(mm), cu(mv)
    y = F.batchnorm_lrelu_infer(cu(x), cu(gam), cu(bet), gmm, gmv, 1e-3, slope)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu().numpy(), ref.numpy()) < STRICT_TOL
    assert np.array_equal(gmm.cpu().numpy(), mm.astype(np.float32))          # the moving statistics are read-only here


# ---------------------------------------------------------------------------------------- attention
def _run_attn(F, X, dY, w, mode):
    t = {k: cu(np.asarray(v)).requires_grad_(True) for k, v in w.items()}
    tx = cu(X).requires_grad_(True)
    y = F.attention(tx, t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"], mode)
    y.backward(cu(dY))
    torch.cuda.synchronize()
    return y.detach().cpu().numpy(), tx.grad.cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in t.items()}


@pytest.mark.parametrize("case", list(enumerate(mg.ATTN_CASES)))
def test_attention_strict_vs_golden(F, case):
    i, (B, N, C) = case
    gold = np.load(os.path.join(GOLD, "attention.npz"))
    X, dY, w = mg.attn_inputs(B, N, C, 200 + i)
    y, dx, gw = _run_attn(F, X, dY, w, F.MATH_FP32_STRICT)
    tag = f"B{B}_N{N}_C{C}"
    assert rel_l2(y, gold[tag + "_Y"]) < STRICT_TOL
    assert rel_l2(dx, gold[tag + "_dX"]) < STRICT_TOL
    for k in oattn.WEIGHT_NAMES:
        if k == "bphi":   # mathematically zero (softmax is invariant to a per-row shift of the logits)
            assert np.abs(gw[k]).max() < 1e-4 * np.abs(gw["btheta"]).max()
            continue
        assert rel_l2(gw[k], gold[tag + "_d" + k]) < 2e-5, k


def test_attention_strict_vs_oracle_ragged_and_large_logits(F):
    """N not a multiple of the tile (200) and un-scaled logits of magnitude ~30 (no 1/sqrt(d), layers.py:108)."""
    B, N, C = 2, 200, 32
    X, dY, w = oattn.make_inputs(B, N, C, seed=5, gamma=0.8, dtype=np.float32)
    w["Wtheta"] = w["Wtheta"] * 6
    w["Wphi"] = w["Wphi"] * 6
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    Y = oattn.forward(X.astype(np.float64), **w64)
    g = oattn.backward(dY.astype(np.float64), X.astype(np.float64), **w64)
    y, dx, gw = _run_attn(F, X, dY, w, F.MATH_FP32_STRICT)
    assert rel_l2(y, Y) < STRICT_TOL
    assert rel_l2(dx, g["dX"]) < 5e-5
    for k in oattn.WEIGHT_NAMES:
        if k == "bphi":
            assert np.abs(gw[k]).max() < 1e-4 * np.abs(gw["btheta"]).max()
            continue
        assert rel_l2(gw[k], g["d" + k]) < 5e-5, k


def test_attention_gamma_zero_is_identity(F):
    """gamma is zero-initialised (layers.py:76-79): the block is the identity and only dgamma is non-zero."""
    X, dY, w = oattn.make_inputs(2, 128, 16, seed=9, gamma=0.0, dtype=np.float32)
    y, dx, gw = _run_attn(F, X, dY, w, F.MATH_FP32_STRICT)
    assert np.array_equal(y, X)
    assert np.array_equal(dx, dY)
    assert abs(gw["gamma"]).max() > 0
    assert all(np.abs(gw[k]).max() == 0 for k in ("Wphi", "Wtheta", "Wg", "Wo", "bo"))


def test_attention_full_size_properties(F):
    """church64_attn G attention at 64x64 (N = 4096, C = 16), B = 4: permutation equivariance over tokens
    and row-stochasticity (constant values => A == that constant), which do not need the O(N^2) oracle."""
    B, N, C = 4, 4096, 16
    X, dY, w = oattn.make_inputs(B, N, C, seed=21, gamma=0.5, dtype=np.float32)
    t = {k: cu(np.asarray(v)) for k, v in w.items()}
    args = (t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"])
    x = cu(X)
    y = F.attention(x, *args)
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    yp = F.attention(x[:, perm].contiguous(), *args)
    torch.cuda.synchronize()
    assert rel_l2(yp.cpu().numpy(), y[:, perm].cpu().numpy()) < 1e-5
    # values independent of the token (Wg = 0): O = bg Wo + bo for every row
    zg = torch.zeros_like(t["Wg"])
    y2 = F.attention(x, t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], zg, t["bg"], t["Wo"], t["bo"], t["gamma"])
    const = t["bg"] @ t["Wo"] + t["bo"]
    torch.cuda.synchronize()
    assert rel_l2(y2.cpu().numpy(), (x + t["gamma"] * const).cpu().numpy()) < 1e-6


# ---------------------------------------------------------------------------------------- losses / optimiser
def test_hinge_and_adam(F):
    rng = np.random.Generator(np.random.PCG64(17))
    dr, df = rng.standard_normal((4, 4, 4, 1)) * 2, rng.standard_normal((4, 4, 4, 1)) * 2
    loss = torch.zeros(2, device="cuda")
    gr, gf = F.hinge_d_grads(cu(dr), cu(df), 8, loss[0:1])
    gg = F.hinge_g_grads(cu(df), 8, loss[1:2])
    torch.cuda.synchronize()
    n = dr.size
    assert abs(float(loss[0]) - (np.maximum(1 - dr, 0) + np.maximum(1 + df, 0)).sum()) < 1e-4
    assert abs(float(loss[1]) - (-df).sum()) < 1e-4
    assert np.allclose(gr.cpu().numpy(), np.where(1 - dr > 0, -1.0, 0.0) / (n * 8))
    assert np.allclose(gf.cpu().numpy(), np.where(1 + df > 0, 1.0, 0.0) / (n * 8))
    assert np.allclose(gg.cpu().numpy(), -1.0 / (n * 8))
    # Keras Adam, beta_1 = 0, three steps
    p0, g = rng.standard_normal(1000), rng.standard_normal((3, 1000))
    p, v = cu(p0), torch.zeros(1000, device="cuda")
    pr, vr = p0.copy(), np.zeros(1000)
    for t in range(1, 4):
        lr_t = 7e-4 * np.sqrt(1 - 0.999 ** t)
        hyper = cu(np.array([lr_t, 0.0, 0.999, 1e-7]))
        F.adam_step(p, cu(g[t - 1]), v, hyper)
        vr = 0.999 * vr + 0.001 * g[t - 1] ** 2
        pr = pr - lr_t * g[t - 1] / (np.sqrt(vr) + 1e-7)
    torch.cuda.synchronize()
    assert rel_l2(p.cpu().numpy(), pr) < 1e-6


# ---------------------------------------------------------------------------------------- attention, BF16_TC (tcgen05)
@pytest.mark.parametrize("case", list(enumerate(mg.ATTN_CASES)))
def test_attention_tc_vs_golden(F, case):
    """tcgen05 / TMEM / TMA forward (bf16 operands, fp32 accumulate) against the fp64 golden: <= 2e-3 rel-L2."""
    i, (B, N, C) = case
    gold = np.load(os.path.join(GOLD, "attention.npz"))
    X, dY, w = mg.attn_inputs(B, N, C, 200 + i)
    y, dx, gw = _run_attn(F, X, dY, w, F.MATH_BF16_TC)
    tag = f"B{B}_N{N}_C{C}"
    errs = {k: rel_l2(gw[k], gold[tag + "_d" + k]) for k in oattn.WEIGHT_NAMES if k != "bphi"}
    e_y, e_att, e_dx = rel_l2(y, gold[tag + "_Y"]), rel_l2(y - X, gold[tag + "_Y"] - X), rel_l2(dx, gold[tag + "_dX"])
    print(tag, "Y %.2e Y-X %.2e dX %.2e dX-dY %.2e" % (e_y, e_att, e_dx, rel_l2(dx - dY, gold[tag + "_dX"] - dY)),
          {k: "%.1e" % v for k, v in errs.items()})
    assert e_y < TC_TOL
    # the attention contribution alone (Y - X) must also be accurate, not just hidden behind the residual
    assert e_att < 1e-2
    assert e_dx < TC_TOL
    # parameter gradients: the cancellation-prone operands of the backward GEMMs (V, dA in dP = dA g^T; dS in the
    # theta / phi gradients) are carried as split bf16 pairs, so only the bf16 rounding of P itself is left for
    # C <= 32.  C = 64 (dv = 32): the 3-term split of dP does not fit the 64-column operand row, V stays rounded.
    tol_w = TC_TOL if C <= 32 else 4e-3
    for k, e in errs.items():
        # d gamma = sum(dY * O) is ONE scalar summed over zero-mean terms (this test's dY is independent of O), so its
        # relative error is a ratio of two random-walk sums: loose bound
        assert e < (1e-2 if k == "gamma" else tol_w), (k, e)


def _check_attn_grads(tag, mode_name, y, dx, gw, X, dY, Y_ref, g_ref, tol, tol_w, tol_att):
    errs = {k: rel_l2(gw[k], g_ref["d" + k]) for k in oattn.WEIGHT_NAMES if k != "bphi"}
    e_y, e_att, e_dx = rel_l2(y, Y_ref), rel_l2(y - X, Y_ref - X), rel_l2(dx, g_ref["dX"])
    e_dxa = rel_l2(dx - dY, g_ref["dX"] - dY)
    print(tag, mode_name, "Y %.2e Y-X %.2e dX %.2e dX-dY %.2e" % (e_y, e_att, e_dx, e_dxa),
          {k: "%.1e" % v for k, v in errs.items()})
    assert e_y < tol and e_dx < tol, (tag, mode_name, e_y, e_dx)
    assert e_att < tol_att and e_dxa < tol_att, (tag, mode_name, e_att, e_dxa)
    for k, e in errs.items():
        # d gamma is ONE scalar summed over zero-mean terms (dY independent of O): ratio of two random-walk sums
        assert e < (5 * tol_w if k == "gamma" else tol_w), (tag, mode_name, k, e)
    assert np.abs(gw["bphi"]).max() < 1e-2 * np.abs(gw["btheta"]).max() + 1e-7     # exactly zero in exact arithmetic


@pytest.mark.parametrize("case", list(enumerate(mg.ATTN_CASES_N1024)))
@pytest.mark.parametrize("mode_name", ["strict", "tc"])
def test_attention_fwd_bwd_vs_golden_n1024(F, case, mode_name):
    """In-model token count N = 1024 (the 32x32 maps), C in {16, 32, 64}: forward AND all gradients of
    the fused kernels against the committed fp64 golden vectors, in both math modes (BASELINE.json tolerances:
    1e-5 strict, 2e-3 tensor cores; the strict gradients sum ~1e3 fp32 terms per output: 2e-5)."""
    i, (B, N, C) = case
    gold = np.load(os.path.join(GOLD, "attention_n1024.npz"))
    X, dY, w = mg.attn_inputs(B, N, C, 300 + i)
    tag = f"B{B}_N{N}_C{C}"
    g_ref = {"dX": gold[tag + "_dX"], **{"d" + k: gold[tag + "_d" + k] for k in oattn.WEIGHT_NAMES}}
    mode = F.MATH_FP32_STRICT if mode_name == "strict" else F.MATH_BF16_TC
    y, dx, gw = _run_attn(F, X, dY, w, mode)
    if mode_name == "strict":
        _check_attn_grads(tag, mode_name, y, dx, gw, X, dY, gold[tag + "_Y"], g_ref, STRICT_TOL, 3e-5, 5e-5)
    else:
        # C = 64 (dv = 32): the 3-term split of the dP contraction does not fit the 64-column operand row, V stays
        # bf16-rounded there -> parameter-gradient tier 4e-3 (measured 2.2e-3 on the kernels, 3.7e-3 on btheta)
        _check_attn_grads(tag, mode_name, y, dx, gw, X, dY, gold[tag + "_Y"], g_ref, TC_TOL, TC_TOL if C <= 32 else 4e-3, 1e-2)


@pytest.mark.parametrize("mode_name", ["strict", "tc"])
def test_attention_fwd_bwd_vs_oracle_n4096(F, mode_name):
    """church64_attn G attention at the 64x64 map (N = 4096, C = 16 -- the step's dominant kernel) against the fp64
    oracle evaluated here (one sample: the materialised [N, N] maps are 128 MiB each in fp64)."""
    B, N, C = 1, 4096, 16
    X, dY, w = oattn.make_inputs(B, N, C, seed=404, gamma=0.37, dtype=np.float32)
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    Y_ref = oattn.forward(X.astype(np.float64), **w64)
    g_ref = oattn.backward(dY.astype(np.float64), X.astype(np.float64), **w64)
    mode = F.MATH_FP32_STRICT if mode_name == "strict" else F.MATH_BF16_TC
    y, dx, gw = _run_attn(F, X, dY, w, mode)
    if mode_name == "strict":
        _check_attn_grads("B1_N4096_C16", mode_name, y, dx, gw, X, dY, Y_ref, g_ref, STRICT_TOL, 5e-5, 1e-4)
    else:
        _check_attn_grads("B1_N4096_C16", mode_name, y, dx, gw, X, dY, Y_ref, g_ref, TC_TOL, TC_TOL, 1e-2)


@pytest.mark.parametrize("shape", [(3, 512, 16), (2, 1024, 32)])
@pytest.mark.parametrize("log2_scale", [-40, -24, 14])
def test_attention_tc_backward_gradient_range(F, shape, log2_scale):
    """The C <= 32 tensor-core backward carries dS = P' (dP - D) as ONE fp16 term (5 exponent bits), normalised per
    sample by a power of two taken from max |dY|.  Gradients of the size a GAN step really sees (|dY| ~ 1e-7 ... 1e-12)
    and large ones must come out as accurately as O(1) ones: with dY scaled by an exact power of two every gradient
    must scale by exactly that factor (one sample is scaled a further 2^-9 so the samples' scales differ)."""
    B, N, C = shape
    X, dY, w = oattn.make_inputs(B, N, C, seed=91 + C, gamma=0.37, dtype=np.float32)
    per_sample = np.ones((B, 1, 1), np.float32)
    per_sample[-1] = 2.0 ** -9
    _, dx0, gw0 = _run_attn(F, X, dY * per_sample, w, F.MATH_BF16_TC)
    c = np.float32(2.0 ** log2_scale)
    _, dx1, gw1 = _run_attn(F, X, dY * per_sample * c, w, F.MATH_BF16_TC)
    # the residual path adds dY itself: compare the attention part of dX
    att0, att1 = dx0 - dY * per_sample, (dx1 - dY * per_sample * c) / c
    assert rel_l2(att1, att0) < 1e-5, rel_l2(att1, att0)
    for k in oattn.WEIGHT_NAMES:
        if k == "bphi":
            continue
        assert rel_l2(gw1[k] / c, gw0[k]) < 1e-4, (k, rel_l2(gw1[k] / c, gw0[k]))      # atomics: summation order differs


@pytest.mark.parametrize("shape", [(45, 512, 16), (23, 900, 32), (160, 100, 16)])
def test_attention_tc_persistent_work_items(F, shape):
    """The C <= 32 tensor-core forward is a persistent kernel (one CTA per SM walks the (sample, 128-query tile) work
    items, barrier phases and K / V rings running on across items): more items than the 148 SMs, so CTAs take 2 (and 3)
    items, with 1 / 4 / 8 key tiles per item, full and ragged last tiles, both S-buffer depths (C = 16: 3, C = 32: 2).
    Forward and every gradient against the fp64 oracle."""
    B, N, C = shape
    X, dY, w = oattn.make_inputs(B, N, C, seed=77 + N, gamma=0.37, dtype=np.float32)
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    Y_ref = oattn.forward(X.astype(np.float64), **w64)
    g_ref = oattn.backward(dY.astype(np.float64), X.astype(np.float64), **w64)
    y, dx, gw = _run_attn(F, X, dY, w, F.MATH_BF16_TC)
    _check_attn_grads(f"B{B}_N{N}_C{C}", "tc", y, dx, gw, X, dY, Y_ref, g_ref, TC_TOL, TC_TOL, 1e-2)


@pytest.mark.parametrize("shape", [(4, 4096, 16), (4, 1024, 32), (2, 1024, 64), (3, 1000, 16)])
def test_attention_tc_vs_strict_full_size(F, shape):
    """In-model sizes (N = 4096 / 1024): the tensor-core forward against the fp32 CUDA-core forward."""
    B, N, C = shape
    X, dY, w = oattn.make_inputs(B, N, C, seed=31, gamma=0.7, dtype=np.float32)
    w["Wtheta"] = w["Wtheta"] * 3      # un-scaled logits of a few units, like a trained model
    t = {k: cu(np.asarray(v)) for k, v in w.items()}
    args = (t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"])
    x = cu(X)
    ys = F.attention(x, *args, F.MATH_FP32_STRICT)
    yt = F.attention(x, *args, F.MATH_BF16_TC)
    torch.cuda.synchronize()
    assert rel_l2(yt.cpu().numpy(), ys.cpu().numpy()) < TC_TOL
    assert rel_l2((yt - x).cpu().numpy(), (ys - x).cpu().numpy()) < 1e-2


# ---------------------------------------------------------------------------------------- attention, down-sampled keys / values
POOL_CASES = [(2, 16, 16, 16), (2, 16, 16, 32), (1, 32, 32, 64), (2, 8, 8, 16), (2, 8, 16, 32), (3, 12, 10, 16), (1, 64, 64, 16)]


@pytest.mark.parametrize("case", POOL_CASES)
@pytest.mark.parametrize("mode_name", ["strict", "tc"])
def test_attention_pooled_vs_oracle(F, case, mode_name):
    """phi and g max-pooled 2x2 / stride 2 over the [H, W] token grid (what layers.py:96,100,113
    reaches for; N / 4 keys), forward and every gradient against oracle.attention.forward_pooled / backward_pooled
    (fp64), in both math modes.  Covers a ragged key tile (8x8 -> 16 keys), non-square and non-power-of-two grids and the
    in-model 64x64 map (4096 queries, 1024 keys)."""
    B, H, W, C = case
    N = H * W
    X, dY, w = oattn.make_inputs(B, N, C, seed=500 + C + H, gamma=0.37, dtype=np.float32)
    w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    Y_ref = oattn.forward_pooled(X.astype(np.float64), **w64, hw=(H, W))
    g_ref = oattn.backward_pooled(dY.astype(np.float64), X.astype(np.float64), **w64, hw=(H, W))
    mode = F.MATH_FP32_STRICT if mode_name == "strict" else F.MATH_BF16_TC
    t = {k: cu(np.asarray(v)).requires_grad_(True) for k, v in w.items()}
    tx = cu(X).requires_grad_(True)
    y = F.attention(tx, t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"], mode,
                    pool_grid=(H, W))
    y.backward(cu(dY))
    torch.cuda.synchronize()
    gw = {k: v.grad.cpu().numpy() for k, v in t.items()}
    tag = f"pooled B{B}_{H}x{W}_C{C}"
    yk, dx = y.detach().cpu().numpy(), tx.grad.cpu().numpy()
    # not the un-pooled block by accident
    assert rel_l2(yk - X, oattn.forward(X.astype(np.float64), **w64) - X) > 1e-2
    errs = {k: rel_l2(gw[k], g_ref["d" + k]) for k in oattn.WEIGHT_NAMES}
    e_y, e_att, e_dx = rel_l2(yk, Y_ref), rel_l2(yk - X, Y_ref - X), rel_l2(dx, g_ref["dX"])
    print(tag, mode_name, "Y %.2e Y-X %.2e dX %.2e" % (e_y, e_att, e_dx), {k: "%.1e" % v for k, v in errs.items()})
    # d(phi bias) is exactly zero in exact arithmetic here too (max-pooling commutes with a per-channel shift of all keys,
    # and softmax is invariant to it): the computed value is rounding noise and is only bounded
    assert np.abs(gw["bphi"]).max() < 1e-2 * np.abs(gw["btheta"]).max() + 1e-6
    if mode_name == "strict":
        assert e_y < STRICT_TOL and e_dx < STRICT_TOL and e_att < 5e-5
        for k, e in errs.items():
            if k != "bphi":
                assert e < 1e-4, (k, e)     # fp32 sums over up to 4096 tokens (measured <= 5e-5, d(theta bias) at 64x64)
    else:
        tol_w = TC_TOL if C <= 32 else 4e-3
        assert e_y < TC_TOL and e_dx < TC_TOL and e_att < 1e-2
        for k, e in errs.items():
            if k == "bphi":
                continue
            # d gamma and d(theta bias) are sums of zero-mean terms over ALL tokens (d theta_i = sum_j dS_ij phi_j with
            # sum_j dS_ij = 0): their relative error is a ratio of random-walk sums (measured up to 4.2e-3)
            assert e < (5 * tol_w if k in ("gamma", "btheta") else tol_w), (k, e)


def test_attention_layer_pool_option(F):
    """Layer surface: AttentionLayer(pool='2x2s2') routes through sagan_attn_pool_* and differs from the un-pooled
    layer; spectral-norm wrapped kernels, NHWC input."""
    from sagan_b200 import nn as snn
    torch.manual_seed(0)
    x = torch.randn(2, 16, 16, 32, device="cuda")
    a, b = snn.AttentionLayer(pool="2x2s2"), snn.AttentionLayer()
    ya = a(x)
    b(x)
    with torch.no_grad():
        for pa, pb in zip(a.parameters(), b.parameters()):
            pb.copy_(pa)
        a.sigma.fill_(0.5); b.sigma.fill_(0.5)
        for sa, sb in zip(a.SN_conv, b.SN_conv):
            sb._group.u(0).copy_(sa._group.u(0))
    ya, yb = a(x, training=False), b(x, training=False)
    assert ya.shape == x.shape and torch.isfinite(ya).all()
    assert float((ya - yb).norm() / (yb - x).norm()) > 1e-2
    ya.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in a.parameters())


# ---------------------------------------------------------------------------------------- attention, large C (sweep regime)
@pytest.mark.parametrize("shape", [(2, 256, 128), (2, 384, 256), (1, 512, 512), (1, 128, 512), (2, 512, 256), (1, 1024, 512), (1, 256, 512),
                                   (2, 2048, 256)])
def test_attention_tc_large_c_forward(F, shape):
    """BASELINE.json configs[4] regime (C = 128..512: d = 16..64, dv = 64..256): projection GEMM -> flash forward ->
    output-conv GEMM with the gamma residual, all on tcgen05, against the fp64 oracle.  Logits here are plain bf16
    products (no room for the split-bf16 trick at d >= 16), so the inputs keep them at a few units (std 1.5), the
    regime a spectrally-normalised trained block lives in; the error grows linearly with the logit scale."""
    B, N, C = shape
    d = C // 8
    X, dY, w = oattn.make_inputs(B, N, C, seed=77 + C, gamma=0.8, dtype=np.float64)
    w["Wtheta"] = w["Wtheta"] * (1.5 ** 0.5) / d ** 0.25
    w["Wphi"] = w["Wphi"] * (1.5 ** 0.5) / d ** 0.25
    ref, cache = oattn.forward(X, **w, return_cache=True)
    t = {k: cu(np.asarray(v)) for k, v in w.items()}
    x = cu(X)
    y = F.attention(x, t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"],
                    F.MATH_BF16_TC)
    torch.cuda.synchronize()
    yk = y.cpu().numpy()
    e_y, e_att = rel_l2(yk, ref), rel_l2(yk - X, ref - X)
    print(shape, "Y %.2e Y-X %.2e" % (e_y, e_att))
    assert e_y < TC_TOL
    assert e_att < 1e-2


@pytest.mark.parametrize("shape", [(2, 256, 128), (1, 384, 256), (1, 256, 512)])
def test_attention_tc_large_c_backward(F, shape):
    """Large-C backward (composed of the library's tensor-core GEMMs and two row kernels, csrc/attn_big_bwd.cu)
    against the fp64 oracle gradients.  The accuracy tier of this regime is set by the plain bf16 / tf32 logits (no
    room for the split-bf16 trick at d >= 16): the forward's attention term Y - X is good to 4-5e-3 (bound 1e-2 in
    test_attention_tc_large_c_forward), and the gradients inherit exactly that -- dWo = gamma A^T dY uses the saved A.
    Bounds: dX 3e-3 (dominated by the exact dY term), its attention part and the parameter gradients 1e-2, the
    scalar d gamma loosely."""
    B, N, C = shape
    d = C // 8
    X, dY, w = oattn.make_inputs(B, N, C, seed=91 + C, gamma=0.8, dtype=np.float64)
    w["Wtheta"] = w["Wtheta"] * (1.5 ** 0.5) / d ** 0.25
    w["Wphi"] = w["Wphi"] * (1.5 ** 0.5) / d ** 0.25
    gref = oattn.backward(dY, X, **w)
    t = {k: cu(np.asarray(v)).requires_grad_(True) for k, v in w.items()}
    x = cu(X).requires_grad_(True)
    y = F.attention(x, t["Wtheta"], t["btheta"], t["Wphi"], t["bphi"], t["Wg"], t["bg"], t["Wo"], t["bo"], t["gamma"],
                    F.MATH_BF16_TC)
    y.backward(cu(dY))
    torch.cuda.synchronize()
    e_dx = rel_l2(x.grad.cpu().numpy(), gref["dX"])
    e_att = rel_l2(x.grad.cpu().numpy() - dY, gref["dX"] - dY)
    errs = {k: rel_l2(t[k].grad.cpu().numpy(), gref["d" + k]) for k in oattn.WEIGHT_NAMES if k != "bphi"}
    print(shape, "dX %.2e dX-dY %.2e" % (e_dx, e_att), {k: "%.1e" % v for k, v in errs.items()})
    assert e_dx < 3e-3
    assert e_att < 1e-2
    for k, e in errs.items():
        assert e < (5e-2 if k == "gamma" else 1e-2), (k, e)
    # the key bias gradient is exactly zero in exact arithmetic (softmax is shift-invariant along the keys)
    assert np.abs(t["bphi"].grad.cpu().numpy()).max() < 1e-2 * np.abs(t["btheta"].grad.cpu().numpy()).max() + 1e-6


# ---------------------------------------------------------------------------------------- conv family, BF16_TC (tcgen05)
TC_CONV_CASES = [  # B, H, W, Cin, Cout, k, stride
    (2, 32, 32, 16, 32, 4, 2), (2, 16, 16, 32, 64, 4, 2), (2, 8, 8, 64, 128, 4, 2), (4, 64, 64, 16, 3, 4, 1),
    (2, 4, 4, 128, 8, 4, 1), (3, 9, 7, 8, 24, 3, 2), (2, 16, 16, 16, 16, 1, 1), (1, 5, 5, 8, 200, 3, 1),
]


def bf16r(a):
    """Round to bf16 (round-to-nearest-even), back to float64: the operand values the tensor cores see."""
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(torch.bfloat16).to(torch.float64)


def tf32r(a, mode):
    """fp32 -> tf32 (10 mantissa bits) -> float64; mode 'trunc' drops the low 13 bits, 'rn' rounds to nearest."""
    t = torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous()
    bits = t.view(torch.int32)
    if mode == "rn":
        bits = bits + 0x1000
    return (bits & ~0x1FFF).view(torch.float32).to(torch.float64)


@pytest.fixture(params=["split_bf16", "tf32"])
def conv_prec(request):
    """Operand precision of the tensor-core conv kernels (sagan_conv_tc_precision): split-bf16 is the default of
    BF16_TC mode, tf32 (+ plain bf16 backward-filter) the round-1 arithmetic kept for comparison."""
    from sagan_b200 import _lib
    lib = _lib.load()
    lib.sagan_conv_tc_precision(_lib.CONV_TC_SPLIT_BF16 if request.param == "split_bf16" else _lib.CONV_TC_TF32)
    yield request.param
    lib.sagan_conv_tc_precision(_lib.CONV_TC_SPLIT_BF16)


SPLIT_TOL = 5e-5     # split-bf16 products carry 16 mantissa bits (dropped lo*lo term 2^-16): fp32-grade results


@pytest.mark.parametrize("case", TC_CONV_CASES)
def test_conv2d_tc_split_vs_fp64(F, case):
    """Default BF16_TC conv arithmetic (split-bf16, three MMAs per K step): forward, backward-data, backward-filter and
    bias gradient of Conv2D + LeakyReLU against UN-ROUNDED fp64 (the kernel's own activation mask is used for the
    reference gradient: a pre-activation within 1e-5 of zero may legitimately land on either side)."""
    from sagan_b200 import _lib
    assert _lib.load().sagan_conv_tc_precision(-1) == _lib.CONV_TC_SPLIT_BF16
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(41))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cin, Cout)) * 0.1
    b = rng.standard_normal(Cout) * 0.1
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    y = F.conv2d(gx, gw, gb, s, "same", F.ACT_LRELU, 0.1, F.MATH_BF16_TC)
    dy = rng.standard_normal(tuple(y.shape))
    y.backward(cu(dy))
    torch.cuda.synchronize()
    yk = y.detach().cpu().double()
    dz = torch.tensor(dy) * torch.where(yk > 0, 1.0, 0.1)
    tx, tw, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, b))
    z = onets.conv2d_same(tx, tw, tb, s)
    z.backward(dz)
    errs = (rel_l2(yk.numpy(), torch.nn.functional.leaky_relu(z, 0.1).detach().numpy()),
            rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()), rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()),
            rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()))
    print(case, "split-bf16 vs fp64: y %.2e dx %.2e dw %.2e db %.2e" % errs)
    assert max(errs) < SPLIT_TOL


@pytest.mark.parametrize("case", TC_CONV_CASES)
def test_conv2d_tc(F, case):
    """Round-1 arithmetic (sagan_conv_tc_precision = TF32), kept selectable.
    tcgen05 implicit-GEMM conv: forward / backward-data run kind::tf32 on the fp32 operands, backward-filter
    kind::f16 on bf16-converted operands, fp32 accumulation in TMEM.

    (1) Exactness of the kernels: against fp64 torch evaluated on the SAME rounded operands the only difference is
    fp32 accumulation order -> 2e-5 (tf32: whichever of truncation / round-to-nearest the tensor core applies).
    (2) Against the un-rounded fp64 reference the forward is inside the tf32 tier (the tensor core TRUNCATES fp32 to
    tf32: relative operand error < 2^-10, measured 8e-4 on y; bf16 operands gave 2.3e-3).  Gradients are not compared
    un-rounded through LeakyReLU: a rounding-sized perturbation of y flips the sign of the few outputs nearest zero
    and each flip changes dz by 0.9 dy -- a property of reduced-precision operands, not of the kernel."""
    from sagan_b200 import _lib
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(41))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cin, Cout)) * 0.1
    b = rng.standard_normal(Cout) * 0.1
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    _lib.load().sagan_conv_tc_precision(_lib.CONV_TC_TF32)
    try:
        y = F.conv2d(gx, gw, gb, s, "same", F.ACT_LRELU, 0.1, F.MATH_BF16_TC)
        dy = rng.standard_normal(tuple(y.shape))
        y.backward(cu(dy))
        torch.cuda.synchronize()
    finally:
        _lib.load().sagan_conv_tc_precision(_lib.CONV_TC_SPLIT_BF16)
    yk = y.detach().cpu().double()
    tb = torch.tensor(b, dtype=torch.float64)
    dz = torch.tensor(dy) * torch.where(yk > 0, 1.0, 0.1)        # the kernel's own activation mask
    tc_dgrad = Cout % 8 == 0                                     # ragged channel counts run the fp32 CUDA-core dgrad
    direct_fwd = (Cin, Cout, k, s) == (16, 3, 4, 1)              # the image-side layers run direct fp32 kernels (conv_small.cu)
    e_y = e_dx = 1.0
    for mode in ("trunc", "rn"):
        ref_r = torch.nn.functional.leaky_relu(onets.conv2d_same(torch.tensor(x) if direct_fwd else tf32r(x, mode),
                                                                 torch.tensor(w) if direct_fwd else tf32r(w, mode), tb, s), 0.1)
        e_y = min(e_y, rel_l2(yk.numpy(), ref_r.numpy()))
        ux = torch.tensor(x, requires_grad=True)
        onets.conv2d_same(ux, tf32r(w, mode) if tc_dgrad else torch.tensor(w), None, s).backward(
            tf32r(dz, mode) if tc_dgrad else dz)
        e_dx = min(e_dx, rel_l2(gx.grad.cpu().numpy(), ux.grad.numpy()))
    # backward-filter: bf16 operands
    tx, tw = bf16r(x).requires_grad_(True), bf16r(w).requires_grad_(True)
    onets.conv2d_same(tx, tw, None, s).backward(bf16r(dz))
    e_dw = rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy())
    e_db = rel_l2(gb.grad.cpu().numpy(), bf16r(dz).sum(dim=(0, 1, 2)).numpy())
    # (2) un-rounded reference, forward
    ref = torch.nn.functional.leaky_relu(onets.conv2d_same(torch.tensor(x), torch.tensor(w), tb, s), 0.1)
    e_true = rel_l2(yk.numpy(), ref.numpy())
    print(case, "same-operand: y %.2e dx %.2e dw %.2e db %.2e | true y %.2e" % (e_y, e_dx, e_dw, e_db, e_true))
    assert max(e_y, e_dx, e_dw, e_db) < 2e-5
    assert e_true < 1.2e-3


@pytest.mark.parametrize("case", TC_CONV_CASES[:4])
def test_conv2d_tc_linear_vs_fp64(F, case, conv_prec):
    """No activation (so no mask flips): fwd / dgrad / wgrad / dbias against un-rounded fp64 inside the BF16_TC tier."""
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(42))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cin, Cout)) * 0.1
    b = rng.standard_normal(Cout) * 0.1
    tx, tw, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, b))
    ref = onets.conv2d_same(tx, tw, tb, s)
    dy = rng.standard_normal(tuple(ref.shape))
    ref.backward(torch.tensor(dy))
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    y = F.conv2d(gx, gw, gb, s, "same", F.ACT_NONE, 0.0, F.MATH_BF16_TC)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    errs = (rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()), rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()),
            rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()), rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()))
    print(case, conv_prec, "y %.2e dx %.2e dw %.2e db %.2e" % errs)
    assert max(errs) < (SPLIT_TOL if conv_prec == "split_bf16" else 2 * TC_TOL)


@pytest.mark.parametrize("case", [(2, 4, 4, 256, 128, 4, 2), (2, 16, 16, 64, 32, 4, 2), (2, 32, 32, 32, 16, 4, 2),
                                  (2, 5, 6, 8, 16, 3, 2)])
def test_conv2d_transpose_tc(F, case, conv_prec):
    B, H, W, Cin, Cout, k, s = case
    rng = np.random.Generator(np.random.PCG64(43))
    x = rng.standard_normal((B, H, W, Cin))
    w = rng.standard_normal((k, k, Cout, Cin)) * 0.1
    tx, tw = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w))
    ref = onets.conv2d_transpose_same(tx, tw, s)
    dy = rng.standard_normal(tuple(ref.shape))
    ref.backward(torch.tensor(dy))
    gx, gw = (cu(a).requires_grad_(True) for a in (x, w))
    y = F.conv2d_transpose(gx, gw, s, "same", F.MATH_BF16_TC)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    errs = (rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()), rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()),
            rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()))
    print(case, conv_prec, "y %.2e dx %.2e dw %.2e" % errs)
    assert max(errs) < (SPLIT_TOL if conv_prec == "split_bf16" else 5e-3)


def test_dense_tc(F, conv_prec):
    rng = np.random.Generator(np.random.PCG64(44))
    x, w, b = rng.standard_normal((64, 128)), rng.standard_normal((128, 4096)) * 0.05, rng.standard_normal(4096)
    tx, tw, tb = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, w, b))
    ref = tx @ tw + tb
    dy = rng.standard_normal((64, 4096))
    ref.backward(torch.tensor(dy))
    gx, gw, gb = (cu(a).requires_grad_(True) for a in (x, w, b))
    y = F.dense(gx, gw, gb, F.MATH_BF16_TC)
    y.backward(cu(dy))
    torch.cuda.synchronize()
    errs = (rel_l2(y.detach().cpu().numpy(), ref.detach().numpy()), rel_l2(gx.grad.cpu().numpy(), tx.grad.numpy()),
            rel_l2(gw.grad.cpu().numpy(), tw.grad.numpy()), rel_l2(gb.grad.cpu().numpy(), tb.grad.numpy()))
    print("dense", conv_prec, "y %.2e dx %.2e dw %.2e db %.2e" % errs)
    assert max(errs) < (SPLIT_TOL if conv_prec == "split_bf16" else 5e-3)


# ---------------------------------------------------------------------------------------- weight norm / record decode
@pytest.mark.parametrize("shape", [(4, 4, 256, 128), (4, 4, 3, 16), (128, 4096), (1, 1, 32, 4), (3, 3, 24, 40)])
def test_weight_norm_kernel_vs_oracle(F, shape):
    """sagan/layers.py:124: kernel = l2_normalize(v, all axes but the last) * g, forward and both gradients."""
    from oracle import weightnorm as own
    rng = np.random.Generator(np.random.PCG64(61))
    v = rng.standard_normal(shape) * 0.05
    g = rng.uniform(0.5, 2.0, shape[-1])
    dw = rng.standard_normal(shape)
    tv, tg = cu(v).requires_grad_(True), cu(g).requires_grad_(True)
    w = F.weight_norm(tv, tg)
    w.backward(cu(dw))
    torch.cuda.synchronize()
    dv_ref, dg_ref = own.backward(dw, v, g)
    assert rel_l2(w.detach().cpu().numpy(), own.kernel_from_vg(v, g)) < 1e-6
    assert rel_l2(tv.grad.cpu().numpy(), dv_ref) < 1e-5 and rel_l2(tg.grad.cpu().numpy(), dg_ref) < 1e-5
    # per-filter norms of the result equal g (size-independent property)
    assert np.allclose(np.sqrt((w.detach().cpu().numpy().reshape(-1, shape[-1]) ** 2).sum(0)), g, rtol=1e-5)


@pytest.mark.parametrize("data_init", [True, False])
def test_weight_normalization_layer(F, data_init):
    """The wrapper of sagan/layers.py:6-211 around Conv2D: data-dependent initialisation (first batch gets zero-mean,
    unit-variance pre-activations per filter, sagan/layers.py:159-194) or g = ||v|| (:152-157), then
    layer(x) with kernel v / ||v|| * g, against the oracle."""
    from oracle import weightnorm as own
    from sagan_b200 import nn as snn
    rng = np.random.Generator(np.random.PCG64(62))
    x = rng.standard_normal((4, 16, 16, 8))
    layer = snn.WeightNormalization(snn.Conv2D(24, 3, 1, padding="same"), data_init=data_init)
    tx = cu(x)
    y = layer(tx)
    torch.cuda.synchronize()
    v = layer.v.detach().cpu().double().numpy()
    conv = lambda k, b: onets.conv2d_same(torch.tensor(x), torch.tensor(k), torch.tensor(b), 1).numpy()
    if data_init:
        g_ref, b_ref = own.data_dep_init(conv(v, np.zeros(24)), np.ones(24), np.zeros(24))
        # (the literal reference takes the moments of the layer with the RAW kernel v, sagan/layers.py:103,169, so the
        # first batch comes out with zero mean / unit variance only up to the factor ||v||: kept as is, oracle == kernel)
    else:
        g_ref, b_ref = own.init_norm(v), np.zeros(24)
    assert rel_l2(layer.g.detach().cpu().numpy(), g_ref) < 1e-5
    assert np.abs(layer.layer.bias.detach().cpu().numpy() - b_ref).max() < 1e-5
    y_ref = conv(own.kernel_from_vg(v, g_ref), b_ref)
    assert rel_l2(y.detach().cpu().numpy(), y_ref) < 2e-5
    if not data_init:      # g = ||v||: the first call reproduces the un-normalised layer
        assert rel_l2(y.detach().cpu().numpy(), conv(v, b_ref)) < 2e-5
    # second call: no re-initialisation, gradients flow to v and g
    y2 = layer(tx)
    y2.square().sum().backward()
    assert layer.g.grad is not None and layer.v.grad is not None and torch.isfinite(layer.v.grad).all()
    assert rel_l2(layer.g.detach().cpu().numpy(), g_ref) < 1e-5


@pytest.mark.parametrize("n", [64 * 64 * 3 * 4, 1003])
def test_record_decode_is_bit_exact(F, n):
    """sagan/dataset.py:31-34: `cast(uint8, float32) * (2. / 255) - 1.` -- byte work: bit-exact against numpy float32."""
    from oracle import weightnorm as own
    rng = np.random.Generator(np.random.PCG64(63))
    raw = rng.integers(0, 256, n, dtype=np.uint8)
    if n > 256:
        raw[:256] = np.arange(256, dtype=np.uint8)          # every byte value
    got = F.decode_records(torch.tensor(raw).cuda()).cpu().numpy()
    ref = own.decode_records(raw)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert got.min() >= -1.0 and got.max() <= 1.0
