#!/usr/bin/env python
"""Diagnostics: per-phase cycle counts of attn_fwd_tc_kernel (library built with EXTRA=-DSAGAN_TIMELINE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-attention-gan_b200"))
import numpy as np, torch
import sagan_b200.functional as F
from sagan_b200 import _lib
B, N, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d, dv = C // 8, C // 2
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, N, C, device="cuda", generator=g)
mk = lambda *s: torch.randn(*s, device="cuda", generator=g) / np.sqrt(s[0])
w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C), torch.tensor(0.5, device="cuda")]
with torch.no_grad():
    for _ in range(3):
        F.attention(x, *w, F.MATH_BF16_TC)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 48)()
_lib.load().sagan_debug_fwd_timeline(out)
nt = (N + 127) // 128
names = ["misc", "waitS", "ld", "max+check", "waitPbuf", "exp+store", "fence+arrive"]
print("softmax w0", {n: int(out[i]) // nt for i, n in enumerate(names)}, "per tile")
