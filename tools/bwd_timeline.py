#!/usr/bin/env python
"""Diagnostics: per-phase cycle counts of attn_bwd_tc_kernel (library built with EXTRA=-DSAGAN_TIMELINE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-attention-gan_b200"))
import numpy as np, torch
import sagan_b200.functional as F
from sagan_b200 import _lib
B, N, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d, dv = C // 8, C // 2
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
mk = lambda *s: (torch.randn(*s, device="cuda", generator=g) / np.sqrt(s[0])).requires_grad_(True)
w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C), torch.tensor(0.5, device="cuda", requires_grad=True)]
dy = torch.randn(B, N, C, device="cuda", generator=g)
y = F.attention(x, *w, F.MATH_BF16_TC)
for _ in range(3):
    torch.autograd.grad(y, [x] + w, dy, retain_graph=True)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 48)()
_lib.load().sagan_debug_bwd_timeline(out)
nt = (N + 127) // 128
names = ["misc", "waitS", "ldS", "exp", "waitDP", "ldDP", "dSmath", "waitG", "store+arrive", "flush_dq"]
for base, who in ((0, "compute w0"),):
    print(who, {n: int(out[base + i]) // nt for i, n in enumerate(names)}, "per tile")
#