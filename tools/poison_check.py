#!/usr/bin/env python
"""Finds kernels whose result depends on uninitialised workspace / output memory: the caching allocator's free blocks are
filled with NaN (or a huge finite value) before each op, so torch.empty() hands out poisoned memory (diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import sagan_b200.functional as F  # noqa: E402
from sagan_b200 import MATH_BF16_TC, MATH_FP32_STRICT  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)


def poison(val):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((64 * 1024 * 1024,), val, device="cuda") for _ in range(6)]     # 1.5 GB
    small = [torch.full((n,), val, device="cuda") for n in (256, 1024, 4096, 16384, 65536, 262144) for _ in range(32)]
    torch.cuda.synchronize()
    del blocks, small


def attn(B, N, C, mode, need_w):
    d, dv = C // 8, C // 2
    x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
    mk = lambda *sh: (torch.randn(*sh, device="cuda", generator=g) / sh[0] ** 0.5).requires_grad_(need_w)
    w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C), torch.tensor(0.5, device="cuda", requires_grad=need_w)]
    dy = torch.randn(B, N, C, device="cuda", generator=g)
    res = []
    for val in (0.0, float("nan"), 3e4):
        poison(val)
        y = F.attention(x, *w, mode)
        gr = torch.autograd.grad(y, [x] + (w if need_w else []), dy)
        torch.cuda.synchronize()
        res.append([y.detach().clone()] + [t.clone() for t in gr])
    for i, val in ((1, "nan"), (2, "3e4")):
        errs = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(res[i], res[0])]
        print(f"attn B{B} N{N} C{C} mode{mode} need_w={need_w} poison={val}: max rel diff vs zero-filled %.2e" % max(
            e if e == e else float("inf") for e in errs), ["%.1e" % e for e in errs])


def conv(B, H, Cin, Cout, k, s, mode, transpose=False, bias=True, act=1):
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(k, k, Cout, Cin, device="cuda", generator=g) * 0.1).requires_grad_(True) if transpose else \
        (torch.randn(k, k, Cin, Cout, device="cuda", generator=g) * 0.1).requires_grad_(True)
    b = torch.randn(Cout, device="cuda", generator=g).requires_grad_(True) if (bias and not transpose) else None
    res = []
    dy = None
    for val in (0.0, float("nan"), 3e4):
        poison(val)
        y = F.conv2d_transpose(x, w, s, "same", mode) if transpose else F.conv2d(x, w, b, s, "same", act, 0.1, mode)
        if dy is None:
            dy = torch.randn(*y.shape, device="cuda", generator=g)
        gr = torch.autograd.grad(y, [x, w] + ([b] if b is not None else []), dy)
        torch.cuda.synchronize()
        res.append([y.detach().clone()] + [t.clone() for t in gr])
    for i, val in ((1, "nan"), (2, "3e4")):
        errs = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(res[i], res[0])]
        print(f"conv B{B} H{H} {Cin}->{Cout} k{k} s{s} T={transpose} mode{mode} poison={val}: %.2e" % max(
            e if e == e else float("inf") for e in errs), ["%.1e" % e for e in errs])


def bn(shape):
    x = torch.randn(*shape, device="cuda", generator=g, requires_grad=True)
    gm = torch.ones(shape[-1], device="cuda", requires_grad=True)
    bt = torch.zeros(shape[-1], device="cuda", requires_grad=True)
    dy = torch.randn(*shape, device="cuda", generator=g)
    res = []
    for val in (0.0, float("nan"), 3e4):
        poison(val)
        y = F.batchnorm_lrelu(x, gm, bt)
        gr = torch.autograd.grad(y, [x, gm, bt], dy)
        torch.cuda.synchronize()
        res.append([y.detach().clone()] + [t.clone() for t in gr])
    for i, val in ((1, "nan"), (2, "3e4")):
        errs = [float((a - b).norm() / (b.norm() + 1e-30)) for a, b in zip(res[i], res[0])]
        print(f"bn {shape} poison={val}: %.2e" % max(e if e == e else float("inf") for e in errs))


TC, ST = MATH_BF16_TC, MATH_FP32_STRICT
for mode in (TC, ST):
    for (B, N, C) in ((4, 1024, 16), (4, 1024, 32), (4, 4096, 16), (2, 1024, 64), (3, 1000, 16)):
        for need_w in (True, False):
            attn(B, N, C, mode, need_w)
attn(1, 512, 256, TC, True)
for mode in (TC, ST):
    conv(4, 64, 3, 16, 4, 2, mode); conv(4, 32, 16, 32, 4, 2, mode); conv(4, 16, 32, 64, 4, 2, mode); conv(4, 8, 64, 128, 4, 2, mode)
    conv(4, 4, 128, 1, 4, 1, mode, act=0); conv(4, 64, 16, 3, 4, 1, mode, bias=False, act=2)
    conv(4, 4, 256, 128, 4, 2, mode, True); conv(4, 8, 128, 64, 4, 2, mode, True); conv(4, 16, 64, 32, 4, 2, mode, True)
    conv(4, 32, 32, 16, 4, 2, mode, True)
    conv(4, 1, 128, 4096, 1, 1, mode, act=0)
for shape in ((4, 8, 8, 128), (4, 64, 64, 16)):
    bn(shape)
