#!/usr/bin/env python
"""Run-to-run / eager-vs-graph agreement of one training step, per parameter (diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import make_golden as mg  # noqa: E402
from sagan_b200 import MATH_BF16_TC, MATH_FP32_STRICT, nn as snn  # noqa: E402
from sagan_b200.trainer import Trainer  # noqa: E402


def mk(mode, **kw):
    snn.set_default_math_mode(mode)
    t = Trainer(dict(mg.TEST_CFG), seed=3, steps_per_epoch=2, **kw)
    snn.set_default_math_mode(MATH_FP32_STRICT)
    return t


def report(tag, a, b):
    for name, na, nb in (("G", a.G, b.G), ("D", a.D, b.D)):
        worst = []
        base = na.flat_params.data_ptr()
        for k, p in na.named_parameters_by_oracle_name():
            off = (p.data_ptr() - base) // 4
            ga, gb = na.flat_grads[off:off + p.numel()], nb.flat_grads[off:off + p.numel()]
            worst.append((float((ga - gb).norm() / (gb.norm() + 1e-30)), k))
        worst.sort(reverse=True)
        tot = float((na.flat_grads - nb.flat_grads).norm() / nb.flat_grads.norm())
        print(tag, name, "flat_grads rel %.2e" % tot, [(k, "%.1e" % e) for e, k in worst[:5]])


mode = MATH_BF16_TC if (len(sys.argv) < 2 or sys.argv[1] == "tc") else MATH_FP32_STRICT
cfg = mg.TEST_CFG
B = cfg["batch_size"]
g = torch.Generator(device="cuda").manual_seed(99)
img = torch.rand(B, 64, 64, 3, device="cuda", generator=g) * 2 - 1
nd = [torch.randn(B, 128, device="cuda", generator=g)]
ng = torch.randn(B, 128, device="cuda", generator=g)
for overlap in (False, True):
    a, b = mk(mode, overlap_streams=overlap), mk(mode, overlap_streams=overlap)
    a.train_step(img, noises_d=nd, noise_g=ng); b.train_step(img, noises_d=nd, noise_g=ng)
    torch.cuda.synchronize()
    report(f"eager/eager overlap={overlap}", a, b)
    c = mk(mode, overlap_streams=overlap)
    c.capture(static_noise=True)
    c.graph_step(img, noises_d=nd, noise_g=ng)
    torch.cuda.synchronize()
    report(f"eager/graph overlap={overlap}", a, c)
