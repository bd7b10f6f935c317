#!/usr/bin/env python
"""Kernel-level breakdown of one attention forward + backward call (torch profiler, warm)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import sagan_b200.functional as F  # noqa: E402
from sagan_b200 import MATH_BF16_TC  # noqa: E402

B, N, C = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 4096, 512)
pool = (int(N ** 0.5), int(N ** 0.5)) if len(sys.argv) > 4 and sys.argv[4] == "pool" else None
d, dv = C // 8, C // 2
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
mk = lambda *sh: (torch.randn(*sh, device="cuda", generator=g) / np.sqrt(sh[0])).requires_grad_(True)
w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C), torch.tensor(0.5, device="cuda", requires_grad=True)]
with torch.no_grad():
    w[0] *= 1.5 ** 0.5 / d ** 0.25
    w[2] *= 1.5 ** 0.5 / d ** 0.25
dy = torch.randn(B, N, C, device="cuda", generator=g)
for _ in range(3):
    y = F.attention(x, *w, MATH_BF16_TC, pool)
    torch.autograd.grad(y, [x] + w, dy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        y = F.attention(x, *w, MATH_BF16_TC, pool)
        torch.autograd.grad(y, [x] + w, dy)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows) / 3
print(f"B={B} N={N} C={C}: {tot:.1f} us of kernel time per forward + backward")
for e in rows[:25]:
    print("%9.1f us  %5.1fx  %5.1f%%  %s" % (e.device_time_total / 3, e.count / 3, 100 * e.device_time_total / 3 / tot, e.key[:90]))
